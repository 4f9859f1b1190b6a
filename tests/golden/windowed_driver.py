"""Executed through scripts/run_reference_script.py (so `windowed_inference` is the reference's own, unchanged file and
`models.*` is whichever side the launcher selected): tags a recording with EATagger and writes the result as JSON with
full-precision probabilities (the script's own __main__ prints two decimals of the top five only).

    run_reference_script.py --side reference --export-ensemble-in-mn-model /abs/tests/golden/windowed_driver.py -- \
        OUT.json WAV WINDOW_S HOP_S cpu|cuda NAME [NAME ...]          (more than one NAME = an ensemble)
"""
import json
import sys

from windowed_inference import EATagger

out, wav, win, hop, device = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4]), sys.argv[5]
names = sys.argv[6:]
tagger = EATagger(model_name=names[0], device=device) if len(names) == 1 else EATagger(ensemble=names, device=device)
tags = tagger.tag_audio_window(wav, window_size=win, hop_length=hop)
with open(out, "w") as f:
    json.dump([{"start": float(w["start"]), "end": float(w["end"]),
                "tags": [[t["tag"], float(t["probability"])] for t in w["tags"]]} for w in tags], f)
