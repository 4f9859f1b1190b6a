"""Per-role cycle accounting of pw_tma_kernel (library built with -DEAT_TMA_TIMING, see scripts/timing/build_tma.sh).
Prints, per layer shape, the average cycles per tile each role spends in each phase (mean over the CTAs)."""
import ctypes, os, sys
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "libeat_tma_timing.so"))
vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
lib.eat_pw_tma_fwd.restype = i32
lib.eat_pw_tma_fwd.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp, i32, vp, i32, vp, vp, i32, vp, vp, vp, vp]
lib.eat_debug_tma_timing.restype = i32
lib.eat_debug_tma_timing.argtypes = [vp]
lib.eat_last_error.restype = ctypes.c_char_p
B = int(os.environ.get("PB", "256"))
st = torch.cuda.current_stream().cuda_stream
buf = torch.zeros(148 * 4 * 8, dtype=torch.int64, device="cuda")
assert lib.eat_debug_tma_timing(buf.data_ptr()) == 0
ROLES = ["TMA thread", "MMA thread", "fix-up warp 0", "epilogue warp 0"]
PH = [["loop", "wait empty", "issue", "-"],
      ["loop", "wait tempty", "wait ready", "issue+commit"],
      ["loop/W", "wait full", "fix-up", "fence+arrive"],
      ["loop", "wait tfull", "wait_read+tmem ld", "math+sts+fence+store"]]
SHAPES = [(32000, 16, 16, "raw"), (32000, 16, 64, "raw"), (32000, 16, 64, "train"), (8000, 64, 24, "raw"), (8000, 24, 72, "raw"),
          (8000, 72, 24, "res"), (2000, 40, 240, "raw"), (504, 112, 672, "raw"), (504, 672, 112, "train"), (128, 960, 160, "raw")]
for (rows, K, N, mode) in SHAPES:
    M = rows * B
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
    C = torch.empty(M, N, device="cuda"); isc = torch.rand(2, K, device="cuda"); stats = torch.zeros(2, N, device="cuda", dtype=torch.float64)
    R = torch.randn(M, N, device="cuda") if mode == "res" else None
    def call():
        if mode == "train":
            return lib.eat_pw_tma_fwd(A.data_ptr(), W.data_ptr(), C.data_ptr(), M, N, K, isc[0].data_ptr(), isc[1].data_ptr(), 2, None, rows,
                                      None, None, 0, None, stats[0].data_ptr(), stats[1].data_ptr(), st)
        return lib.eat_pw_tma_fwd(A.data_ptr(), W.data_ptr(), C.data_ptr(), M, N, K, None, None, 0, None, rows, None, None, 0,
                                  R.data_ptr() if R is not None else None, None, None, st)
    assert call() == 0, lib.eat_last_error()
    torch.cuda.synchronize(); buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); assert call() == 0; e1.record(); torch.cuda.synchronize()
    d = buf.view(148, 4, 8).double().cpu()
    bn = N if N <= 128 else 128
    tiles_per_cta = (M + 127) // 128 * ((N + bn - 1) // bn) / 148
    nbytes = 4 * (M * K + M * N * (2 if R is not None else 1))
    print(f"\nM={M} K={K} N={N} {mode}: {e0.elapsed_time(e1)*1e3:.1f} us ({nbytes/e0.elapsed_time(e1)/1e6:.0f} GB/s), {tiles_per_cta:.0f} tiles/CTA, "
          f"{e0.elapsed_time(e1)*1e3*1965/tiles_per_cta:.0f} cycles/tile at 1965 MHz")
    for r in range(4):
        cnt = d[:, r, 4].clamp(min=1)
        per = (d[:, r, :4] / cnt[:, None]).mean(0)
        print(f"  {ROLES[r]:16s} " + "  ".join(f"{PH[r][i]} {per[i]:7.0f}" for i in range(4)) + f"   (sum {per.sum():.0f} per tile, {cnt.mean():.0f} tiles)")
