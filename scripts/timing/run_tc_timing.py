"""Per-role cycle accounting of pw_tc_kernel (library built with -DEAT_TC_TIMING, see scripts/timing/build.sh).
Prints, per layer shape, the average cycles per tile each role spends in each phase (mean over the 148 CTAs)."""
import ctypes, os, sys
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "libeat_tc_timing.so"))
vp, i32, i64, dp = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
lib.eat_pw_tc_fwd.restype = i32
lib.eat_pw_tc_fwd.argtypes = [vp, i32, vp, i32, vp, i32, i64, i32, i32, vp, vp, i32, vp, i32, vp, vp, i32, vp, vp, vp, vp]
lib.eat_debug_tc_timing.restype = i32
lib.eat_debug_tc_timing.argtypes = [vp]
B = int(os.environ.get("PB", "256"))
st = torch.cuda.current_stream().cuda_stream
buf = torch.zeros(148 * 4 * 8, dtype=torch.int64, device="cuda")
assert lib.eat_debug_tc_timing(buf.data_ptr()) == 0
ROLES = ["producer group 0", "producer group 1", "MMA thread", "epilogue warp 0"]
PH = [["walk/set-up", "wait empty", "load+convert+store", "fence+arrive"],
      ["walk/set-up", "wait empty", "load+convert+store", "fence+arrive"],
      ["loop", "wait tempty", "wait full", "issue+commit"],
      ["loop", "wait tfull", "tmem->regs->smem->global", "fence+arrive"]]
for (rows, K, N, mode) in [(32000, 16, 16, "raw"), (32000, 16, 64, "raw"), (32000, 16, 64, "train"), (8000, 64, 24, "raw"),
                           (8000, 24, 72, "raw"), (2000, 40, 240, "raw"), (504, 112, 672, "raw")]:
    M = rows * B
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
    C = torch.empty(M, N, device="cuda"); isc = torch.rand(2, K, device="cuda"); stats = torch.zeros(2, N, device="cuda", dtype=torch.float64)
    def call():
        if mode == "train":
            return lib.eat_pw_tc_fwd(A.data_ptr(), 0, W.data_ptr(), 0, C.data_ptr(), 0, M, N, K, isc[0].data_ptr(), isc[1].data_ptr(), 2,
                                     None, rows, None, None, 0, None, stats[0].data_ptr(), stats[1].data_ptr(), st)
        return lib.eat_pw_tc_fwd(A.data_ptr(), 0, W.data_ptr(), 0, C.data_ptr(), 0, M, N, K, None, None, 0, None, rows, None, None, 0,
                                 None, None, None, st)
    assert call() == 0
    torch.cuda.synchronize(); buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); assert call() == 0; e1.record(); torch.cuda.synchronize()
    d = buf.view(148, 4, 8).double().cpu()
    tiles_per_cta = (M + 127) // 128 * ((N + 127) // 128) / 148
    print(f"\nM={M} K={K} N={N} {mode}: {e0.elapsed_time(e1)*1e3:.1f} us, {tiles_per_cta:.0f} tiles/CTA, "
          f"{e0.elapsed_time(e1)*1e3*1965/tiles_per_cta:.0f} cycles/tile at 1965 MHz")
    for r in range(4):
        cnt = d[:, r, 4].clamp(min=1)
        per = (d[:, r, :4] / cnt[:, None]).mean(0)
        print(f"  {ROLES[r]:18s} " + "  ".join(f"{PH[r][i]} {per[i]:7.0f}" for i in range(4)) + f"   (sum {per.sum():.0f} per own tile, {cnt.mean():.0f} tiles)")
