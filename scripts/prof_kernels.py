"""One launch each of the training step's hot kernels at mn10 shapes (for `ncu --set full`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientat_b200._lib import lib
B = int(os.environ.get("PB", "64"))
L = lib(); st = torch.cuda.current_stream().cuda_stream
dev = "cuda"
def dw(F, T, C, k, s):
    pad = (k - 1) // 2
    Fo, To = (F + 2 * pad - k) // s + 1, (T + 2 * pad - k) // s + 1
    x = torch.randn(B, F, T, C, device=dev); w = torch.randn(C, 1, k, k, device=dev) * 0.2
    wt = torch.empty(k * k, C, device=dev); L.dw_repack(w.data_ptr(), wt.data_ptr(), C, k, st)
    out = torch.empty(B, Fo, To, C, device=dev); sc = torch.rand(2, C, device=dev)
    stats = torch.zeros(2, C, device=dev, dtype=torch.float64)
    dz = torch.randn(B, Fo, To, C, device=dev); din = torch.empty_like(x); dwg = torch.zeros_like(w)
    L.dw_conv_fwd(x.data_ptr(), wt.data_ptr(), out.data_ptr(), 0, B, F, T, C, k, s, sc[0].data_ptr(), sc[1].data_ptr(), 2, 0, 0, 0, 0, stats[0].data_ptr(), stats[1].data_ptr(), st)
    L.dw_conv_dgrad(dz.data_ptr(), wt.data_ptr(), 0, 0, din.data_ptr(), 0, B, F, T, C, k, s, st)
    L.dw_conv_wgrad(dz.data_ptr(), x.data_ptr(), sc[0].data_ptr(), sc[1].data_ptr(), 2, dwg.data_ptr(), 0, 0, B, F, T, C, k, s, st)
    torch.cuda.synchronize()
def pw(rows, K, N):
    M = rows * B
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5
    C = torch.empty(M, N, device=dev); isc = torch.rand(2, K, device=dev)
    stats = torch.zeros(2, N, device=dev, dtype=torch.float64)
    # training forward: BN+act on load, raw output + statistics
    L.pw_tc_fwd(A.data_ptr(), 0, W.data_ptr(), 0, C.data_ptr(), 0, M, N, K, isc[0].data_ptr(), isc[1].data_ptr(), 2, 0, rows, 0, 0, 0, 0, stats[0].data_ptr(), stats[1].data_ptr(), st)
    # data gradient: dA[M,K] = dZ[M,N] . W  (W^T staged as [K,N])
    Wt = W.t().contiguous(); dA = torch.empty(M, K, device=dev)
    L.pw_tc_fwd(C.data_ptr(), 0, Wt.data_ptr(), 0, dA.data_ptr(), 0, M, K, N, 0, 0, 0, 0, rows, 0, 0, 0, 0, 0, 0, st)
    dW = torch.zeros(N, K, device=dev)
    L.pw_tc_wgrad(C.data_ptr(), 0, A.data_ptr(), 0, dW.data_ptr(), 0, M, N, K, isc[0].data_ptr(), isc[1].data_ptr(), 2, 0, rows, st)
    # BatchNorm backward passes on the [M, N] tensor
    sc = torch.rand(4, N, device=dev); s12 = torch.zeros(2, N, device=dev, dtype=torch.float64); c12 = torch.rand(2, N, device=dev)
    gA = torch.randn(M, N, device=dev); dz = torch.empty(M, N, device=dev)
    L.bn_bwd_reduce(gA.data_ptr(), 0, 0, C.data_ptr(), sc[0].data_ptr(), sc[1].data_ptr(), sc[2].data_ptr(), sc[3].data_ptr(), 2, 0, B, rows, N, s12[0].data_ptr(), s12[1].data_ptr(), st)
    L.bn_bwd_apply(gA.data_ptr(), 0, 0, C.data_ptr(), sc[0].data_ptr(), sc[1].data_ptr(), sc[2].data_ptr(), sc[3].data_ptr(), 2, c12[0].data_ptr(), c12[1].data_ptr(), dz.data_ptr(), 0, B, rows, N, st)
    torch.cuda.synchronize()
which = os.environ.get("PK", "all")
if which in ("all", "dw"):
    dw(64, 500, 64, 3, 2); dw(32, 250, 72, 3, 1); dw(32, 250, 72, 5, 2); dw(16, 125, 120, 5, 1)
if which in ("all", "pw"):
    pw(32000, 16, 64); pw(8000, 72, 24); pw(504, 112, 672)
print("done")
