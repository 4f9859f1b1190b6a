"""Micro-benchmark of the tcgen05 weight-gradient kernel at mn10 layer shapes: dW[N,K] += G[M,N]^T . xf(A)[M,K]."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientat_b200._lib import lib
LAYERS = [(32000, 16, 16, 0), (32000, 16, 64, 0), (8000, 64, 24, 1), (8000, 24, 72, 0), (8000, 72, 24, 1),
          (8000, 24, 72, 0), (2000, 72, 40, 1), (2000, 40, 120, 0), (2000, 120, 40, 1), (2000, 40, 240, 0),
          (504, 240, 80, 1), (504, 80, 200, 0), (504, 200, 80, 1), (504, 80, 480, 0), (504, 480, 112, 1),
          (504, 112, 672, 0), (504, 672, 112, 1), (128, 672, 160, 1), (128, 160, 960, 0), (128, 960, 160, 1)]
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=256); a = ap.parse_args()
L = lib(); st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
tot = 0.0; totb = 0
for i, (rows, K, N, xf) in enumerate(LAYERS):
    M = rows * a.batch
    A = torch.randn(M, K, device="cuda"); G = torch.randn(M, N, device="cuda"); dW = torch.zeros(N, K, device="cuda")
    isc = torch.rand(2, K, device="cuda")
    args = (G.data_ptr(), 0, A.data_ptr(), 0, dW.data_ptr(), 0, M, N, K, isc[0].data_ptr() if xf else 0, isc[1].data_ptr() if xf else 0, 2 if xf else 0, 0, rows, st)
    L.pw_tc_wgrad(*args); ts = []
    for _ in range(6):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.pw_tc_wgrad(*args); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]; nb = M * (K + N) * 4; tot += ms; totb += nb
    print(f"{i:2d} M={M:8d} K={K:4d} N={N:4d} xf={xf}  {ms*1e3:8.1f} us  {nb/ms/1e6:8.1f} GB/s", flush=True)
print(json.dumps({"impl": "pw_tc_wgrad", "batch": a.batch, "total_ms": tot, "total_GBps": totb / tot / 1e6}))
