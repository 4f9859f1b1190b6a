"""Micro-benchmark of the pointwise-conv GEMM kernels at mn10 layer shapes (B clips): time + achieved GB/s.
    python scripts/bench_gemm.py [--batch 32] [--impl pw_tc_fwd] [--dtype fp32] [--only IDX]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientat_b200._lib import lib

# (rows per clip, K, N, residual, in-transform+gate)  -- mn10 forward pointwise layers
LAYERS = [(32000, 16, 16, 1, 0), (32000, 16, 64, 0, 0), (8000, 64, 24, 0, 0), (8000, 24, 72, 0, 0), (8000, 72, 24, 1, 0),
          (8000, 24, 72, 0, 0), (2000, 72, 40, 0, 1), (2000, 40, 120, 0, 0), (2000, 120, 40, 1, 1), (2000, 40, 240, 0, 0),
          (504, 240, 80, 0, 0), (504, 80, 200, 0, 0), (504, 200, 80, 1, 0), (504, 80, 480, 0, 0), (504, 480, 112, 0, 1),
          (504, 112, 672, 0, 0), (504, 672, 112, 1, 1), (128, 672, 160, 0, 1), (128, 160, 960, 0, 0), (128, 960, 160, 1, 1)]

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--impl", default="pw_tc_fwd")
ap.add_argument("--dtype", default="fp32")
ap.add_argument("--only", type=int, default=-1)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--raw", action="store_true", help="no input transform, no epilogue, no statistics (data-gradient shape)")
ap.add_argument("--train", action="store_true", help="training-mode variant: raw output + statistics, BN+act on load")
a = ap.parse_args()
L = lib()
fn = getattr(L, a.impl)
td = torch.float32 if a.dtype == "fp32" else torch.bfloat16
code = 0 if a.dtype == "fp32" else 1
es = 4 if code == 0 else 2
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
tot_t = tot_b = 0.0
for i, (rows, K, N, res, xf) in enumerate(LAYERS):
    if a.only >= 0 and i != a.only:
        continue
    M = rows * a.batch
    A = torch.randn(M, K, device="cuda").to(td)
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    C = torch.empty(M, N, device="cuda", dtype=td)
    R = torch.randn(M, N, device="cuda").to(td) if res else None
    sc = torch.rand(2, N, device="cuda")
    isc = torch.rand(2, K, device="cuda")
    gate = torch.rand(a.batch, K, device="cuda") if xf else None
    stats = torch.zeros(2, N, device="cuda", dtype=torch.float64)
    p = lambda t: 0 if t is None else t.data_ptr()
    if a.raw:
        args = (A.data_ptr(), code, W.data_ptr(), 0, C.data_ptr(), code, M, N, K, 0, 0, 0, 0, rows, 0, 0, 0, 0, 0, 0, st)
    elif a.train:
        args = (A.data_ptr(), code, W.data_ptr(), 0, C.data_ptr(), code, M, N, K, isc[0].data_ptr(), isc[1].data_ptr(), 2,
                p(gate), rows, 0, 0, 0, 0, stats[0].data_ptr(), stats[1].data_ptr(), st)
    else:
        args = (A.data_ptr(), code, W.data_ptr(), 0, C.data_ptr(), code, M, N, K, 0, 0, 0, p(gate), rows,
                sc[0].data_ptr(), sc[1].data_ptr(), 2, p(R), 0, 0, st)
    for _ in range(2):
        fn(*args)
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(*args); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = M * K * es + M * N * es * (2 if (res and not a.train and not a.raw) else 1) + N * K * 4
    tot_t += ms; tot_b += nbytes
    print(f"{i:2d} M={M:8d} K={K:4d} N={N:4d} res={res} xf={xf}  {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s  "
          f"{2*M*N*K/ms/1e9:8.2f} TFLOP/s", flush=True)
print(json.dumps({"impl": a.impl, "dtype": a.dtype, "batch": a.batch, "train": a.train, "total_ms": tot_t,
                  "total_GBps": tot_b / tot_t / 1e6}))
