"""Recipe for `baseline/_ref/`: a byte-for-byte mirror of the reference checkout (fschmid56/EfficientAT) that can
travel to the GPU box, where /root/reference does not exist.

    python baseline/make_ref.py [--src /root/reference]

`baseline/_ref/` is git-ignored (never part of the history or of the product) but NOT gpurun-ignored.  It is used
by exactly three things, all of them measurement / test infrastructure:
  * `bench.py` `gpu_baseline` / `--impl reference-gpu`: the UNMODIFIED reference modules (cuFFT / cuDNN / cuBLAS)
    timed on the same B200, the bar BASELINE.md section 3 names;
  * `tests/test_gpu_refscripts.py`: the reference's own `ex_audioset.py` / `inference.py` run unchanged against
    this package;
  * `tests/golden/make_golden.py`, which may equally read /root/reference directly.
Nothing under efficientat_b200/ imports it.  Large non-code assets (images/, the wandb/ run directory) are skipped.
"""
import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SKIP_DIRS = {".git", "images", "wandb", "__pycache__"}


def make(src="/root/reference", dst=DST):
    if not os.path.isdir(src):
        return None
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        rel = os.path.relpath(root, src)
        out = os.path.join(dst, rel) if rel != "." else dst
        os.makedirs(out, exist_ok=True)
        for f in files:
            if f.endswith((".pyc", ".pt")):
                continue
            s, d = os.path.join(root, f), os.path.join(out, f)
            if not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s) or os.path.getsize(d) != os.path.getsize(s):
                shutil.copy2(s, d)
                n += 1
    return dst, n


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    r = make(a.src)
    print("reference checkout not found, nothing mirrored" if r is None else f"{r[0]}: {r[1]} files updated")
    sys.exit(0)
