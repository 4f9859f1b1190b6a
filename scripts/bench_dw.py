"""Micro-benchmark of the depthwise kernels (fwd training-mode, dgrad, wgrad) at mn10 shapes."""
import argparse, os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientat_b200._lib import lib
# (F, T, C, k, s) of the depthwise input
LAYERS = [(64, 500, 16, 3, 1), (64, 500, 64, 3, 2), (32, 250, 72, 3, 1), (32, 250, 72, 5, 2), (16, 125, 120, 5, 1),
          (16, 125, 240, 3, 2), (8, 63, 200, 3, 1), (8, 63, 184, 3, 1), (8, 63, 480, 3, 1), (8, 63, 672, 3, 1),
          (8, 63, 672, 5, 2), (4, 32, 960, 5, 1)]
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--dtype", default="fp32")
ap.add_argument("--only", type=int, default=-1)
ap.add_argument("--eval", action="store_true")
a = ap.parse_args()
L = lib(); st = torch.cuda.current_stream().cuda_stream
td = torch.float32 if a.dtype == "fp32" else torch.bfloat16
code = 0 if a.dtype == "fp32" else 1; es = 4 if code == 0 else 2
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timeit(fn, n=5):
    fn(); ts = []
    for _ in range(n):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
tot = [0.0, 0.0, 0.0]
for li, (F, T, C, k, s) in enumerate(LAYERS):
    if a.only >= 0 and li != a.only: continue
    B = a.batch; pad = (k - 1) // 2
    Fo, To = (F + 2 * pad - k) // s + 1, (T + 2 * pad - k) // s + 1
    x = torch.randn(B, F, T, C, device="cuda").to(td); w = torch.randn(C, 1, k, k, device="cuda") * 0.2
    wt = torch.empty(k * k, C, device="cuda"); L.dw_repack(w.data_ptr(), wt.data_ptr(), C, k, st)
    out = torch.empty(B, Fo, To, C, device="cuda", dtype=td); sc = torch.rand(2, C, device="cuda")
    stats = torch.zeros(2, C, device="cuda", dtype=torch.float64)
    dz = torch.randn(B, Fo, To, C, device="cuda").to(td); din = torch.empty_like(x); dw = torch.zeros_like(w)
    pool = torch.zeros(B, C, device='cuda')
    f_fwd = (lambda: L.dw_conv_fwd(x.data_ptr(), wt.data_ptr(), out.data_ptr(), code, B, F, T, C, k, s, 0, 0, 0, sc[0].data_ptr(), sc[1].data_ptr(), 2, pool.data_ptr(), 0, 0, st)) if a.eval else (lambda: L.dw_conv_fwd(x.data_ptr(), wt.data_ptr(), out.data_ptr(), code, B, F, T, C, k, s, sc[0].data_ptr(), sc[1].data_ptr(), 2, 0, 0, 0, 0, stats[0].data_ptr(), stats[1].data_ptr(), st))
    f_dg = lambda: L.dw_conv_dgrad(dz.data_ptr(), wt.data_ptr(), 0, 0, din.data_ptr(), code, B, F, T, C, k, s, st)
    f_wg = lambda: L.dw_conv_wgrad(dz.data_ptr(), x.data_ptr(), sc[0].data_ptr(), sc[1].data_ptr(), 2, dw.data_ptr(), 0, code, B, F, T, C, k, s, st)
    nb = B * C * es * (F * T + Fo * To)
    t = [timeit(f) for f in (f_fwd, f_dg, f_wg)]
    for i in range(3): tot[i] += t[i]
    print(f"F={F:3d} T={T:4d} C={C:4d} k={k} s={s}  fwd {t[0]*1e3:7.1f}us {nb/t[0]/1e6:7.0f} GB/s | dgrad {t[1]*1e3:7.1f}us {nb/t[1]/1e6:7.0f} GB/s | wgrad {t[2]*1e3:7.1f}us {nb/t[2]/1e6:7.0f} GB/s", flush=True)
print(json.dumps({"dtype": a.dtype, "batch": a.batch, "fwd_ms": tot[0], "dgrad_ms": tot[1], "wgrad_ms": tot[2]}))
