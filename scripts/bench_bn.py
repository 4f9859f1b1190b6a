"""Micro-benchmark of the BatchNorm-backward passes (reduce, apply) and bn_apply at mn10 tensor shapes: time + GB/s.
    python scripts/bench_bn.py [--batch 256]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientat_b200._lib import lib
SHAPES = [(32000, 64, 2, 0), (32000, 16, 0, 0), (8000, 64, 1, 0), (8000, 72, 1, 0), (8000, 24, 0, 0), (2000, 72, 1, 1), (2000, 120, 2, 1),
          (2000, 40, 0, 0), (504, 240, 2, 0), (504, 480, 2, 1), (504, 112, 0, 0), (128, 672, 2, 1), (128, 960, 2, 1), (128, 160, 0, 0)]
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=256); a = ap.parse_args()
L = lib(); st = torch.cuda.current_stream().cuda_stream; B = a.batch
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timeit(fn):
    fn(); ts = []
    for _ in range(6):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
tot = {"reduce": [0.0, 0], "apply": [0.0, 0]}
for (P, C, act, gated) in SHAPES:
    z = torch.randn(B, P, C, device="cuda"); g = torch.randn(B, P, C, device="cuda"); dz = torch.empty_like(z)
    sc = torch.rand(4, C, device="cuda") + 0.5; c12 = torch.rand(2, C, device="cuda") * 0.01
    s12 = torch.zeros(2, C, device="cuda", dtype=torch.float64)
    gate = torch.rand(B, C, device="cuda") if gated else None; dpool = torch.randn(B, C, device="cuda") * 0.01 if gated else None
    p = lambda t: 0 if t is None else t.data_ptr()
    red = lambda: L.bn_bwd_reduce(g.data_ptr(), p(gate), p(dpool), z.data_ptr(), sc[0].data_ptr(), sc[1].data_ptr(), sc[2].data_ptr(), sc[3].data_ptr(), act, 0, B, P, C, s12[0].data_ptr(), s12[1].data_ptr(), st)
    app = lambda: L.bn_bwd_apply(g.data_ptr(), p(gate), p(dpool), z.data_ptr(), sc[0].data_ptr(), sc[1].data_ptr(), sc[2].data_ptr(), sc[3].data_ptr(), act, c12[0].data_ptr(), c12[1].data_ptr(), dz.data_ptr(), 0, B, P, C, st)
    t1, t2 = timeit(red), timeit(app)
    nb = B * P * C * 4
    tot["reduce"][0] += t1; tot["reduce"][1] += 2 * nb; tot["apply"][0] += t2; tot["apply"][1] += 3 * nb
    print(f"P={P:6d} C={C:4d} act={act} gated={gated}  reduce {t1*1e3:8.1f} us {2*nb/t1/1e6:7.0f} GB/s   apply {t2*1e3:8.1f} us {3*nb/t2/1e6:7.0f} GB/s", flush=True)
print(json.dumps({k: {"total_ms": v[0], "GBps": v[1] / v[0] / 1e6} for k, v in tot.items()}))
