"""Oracle-side measurement (test infrastructure): the reference algorithm as stock PyTorch ops on the GPU
(cuFFT / cuDNN / cuBLAS through torch), i.e. the library path the reference itself takes on a CUDA device
(BASELINE.md section 3 "Reference GPU path").  Prints one JSON line per configuration.

    python oracle/bench_gpu_port.py --batch 128 --steps 10 [--bf16]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import net_oracle  # noqa: E402
from efficientat_b200.models.mn.model import get_model  # noqa: E402
from efficientat_b200.synth import synth_labels, synth_state_, synth_waveform  # noqa: E402


def torch_mel(x, window, fb):
    """reference models/preprocess.py:40-67 with torch.stft (cuFFT) exactly as the reference calls it."""
    x = torch.nn.functional.conv1d(x.unsqueeze(1), torch.tensor([[[-0.97, 1.0]]], device=x.device)).squeeze(1)
    x = torch.stft(x, 1024, hop_length=320, win_length=800, center=True, normalized=False, window=window,
                   return_complex=True)
    p = x.real ** 2 + x.imag ** 2
    with torch.autocast("cuda", enabled=False):
        mel = torch.matmul(fb, p.float())
    return ((mel + 1e-5).log() + 4.5) / 5.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--mode", default="train", choices=["train", "eval"])
    a = ap.parse_args()
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True
    from oracle.mel_oracle import kaldi_mel_banks
    fb = kaldi_mel_banks(128, 1024, 32000, 0.0, 15000.0).to(dev)
    window = torch.hann_window(800, periodic=False, device=dev)
    model = synth_state_(get_model(width_mult=1.0, verbose=False), seed=7)
    sd = {k: v.to(dev) for k, v in model.state_dict().items()}
    names = [k for k, _ in model.named_parameters()]
    for k in names:
        sd[k].requires_grad_(a.mode == "train")
    opt = torch.optim.Adam([sd[k] for k in names], lr=8e-4)
    wave = synth_waveform(a.batch, 320000, seed=1).to(dev)
    y = synth_labels(a.batch, 527, seed=2).to(dev)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.bf16):
            spec = torch_mel(wave, window, fb).unsqueeze(1)
            if a.mode == "eval":
                with torch.no_grad():
                    net_oracle.mn_forward(sd, spec, training=False)
                return
            logits, _ = net_oracle.mn_forward(sd, spec, training=True)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.float(), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"impl": "oracle port on GPU (stock torch: cuFFT/cuDNN/cuBLAS)", "mode": a.mode,
                      "precision": "bf16 autocast" if a.bf16 else "fp32 (cudnn.allow_tf32 default)",
                      "batch": a.batch, "ms_per_step": ms, "clips_per_s": a.batch / ms * 1e3,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
