// Weight gradient of NARROW 1x1 convolutions on CUDA cores (exact fp32 FMAs):
//     dW[N, K] += G[M, N]^T . xf(A)[M, K]            N * K <= 512, M in the millions
// This is the first pointwise layer of the reference network (models/mn/model.py:246: 16->16 channels at 64x500).
// Its reduction dimension is enormous and its output tiny, so a 128-wide tensor-core tile is mostly padding and the
// per-stage hand-shake of the tcgen05 pipeline (wgrad_tcgen05.cu) dominates: measured 1.3 TB/s there, 1.85 TB/s here
// (B=256).  For the next layers (16->64, 64->24, 24->72: N*K >= 1024) the tensor-core kernel is as fast or faster
// (profiles/r01_wgrad_microbench_b256.txt), so the launcher takes only N*K <= 512.
// A CTA streams slabs of rows through shared memory
// (coalesced 16-byte loads, producing layer's BatchNorm + activation applied once per element), every lane keeps a
// 4 x 4 block of dW in registers, and the 8 warps split into row groups that each take every RG-th row of the slab.
// One shared-memory + global atomic flush per CTA.  HBM bound: algorithmic bytes M * (N + K) * 4.
#include "common.cuh"

namespace {

constexpr int kNT = 256;          // threads per CTA
constexpr int kSlab = 64;         // rows per shared-memory slab

template <int XACT>
__device__ __forceinline__ float xact_n(float v) {
  if (XACT == EAT_ACT_RELU) return fmaxf(v, 0.f);
  if (XACT == EAT_ACT_HSWISH) return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
  return v;
}

// XACT: -1 no input transform, 0 affine, 1 affine + ReLU, 2 affine + Hardswish
template <int XACT>
__global__ void __launch_bounds__(kNT, 4) wgrad_narrow_kernel(const float* __restrict__ G, const float* __restrict__ A,
                                                             float* __restrict__ dW, long long M, int N, int K,
                                                             const float* __restrict__ xscale,
                                                             const float* __restrict__ xshift, int wpg /*warps per row group*/) {
  extern __shared__ __align__(16) float smem[];
  const int Np = (N + 3) & ~3, Kp = (K + 3) & ~3;
  float* sG = smem;                       // [kSlab][Np]
  float* sA = sG + kSlab * Np;            // [kSlab][Kp]
  float* sW = sA + kSlab * Kp;            // [Np][Kp] partial dW of this CTA
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < Np * Kp; i += kNT) sW[i] = 0.f;
  const int nt4 = Np >> 2, kt4 = Kp >> 2, tiles = nt4 * kt4;
  const int rg = warp / wpg, RG = (kNT / 32) / wpg;        // this warp's row group / number of row groups
  const int tile = (warp - rg * wpg) * 32 + lane;          // 4x4 output block of this lane
  const bool tact = tile < tiles;
  const int tn = tact ? (tile / kt4) * 4 : 0, tk = tact ? (tile - (tile / kt4) * kt4) * 4 : 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int gv = N >> 2, av = K >> 2;                       // 16-byte vectors per row (N, K are multiples of 4)
  const long long slabs = (M + kSlab - 1) / kSlab;
  for (long long s = blockIdx.x; s < slabs; s += gridDim.x) {
    const long long m0 = s * kSlab;
    const int rows = (int)(M - m0 < kSlab ? M - m0 : kSlab);
    __syncthreads();                                       // previous slab fully consumed (also orders the sW zeroing)
    for (int i = tid; i < rows * gv; i += kNT) {
      const int r = i / gv, c = i - r * gv;
      *reinterpret_cast<float4*>(sG + r * Np + 4 * c) = __ldg(reinterpret_cast<const float4*>(G + (m0 + r) * N) + c);
    }
    for (int i = tid; i < rows * av; i += kNT) {
      const int r = i / av, c = i - r * av;
      float4 v = __ldg(reinterpret_cast<const float4*>(A + (m0 + r) * K) + c);
      if (XACT >= 0) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(xscale) + c), sh = __ldg(reinterpret_cast<const float4*>(xshift) + c);
        v.x = xact_n<XACT>(fmaf(v.x, sc.x, sh.x)); v.y = xact_n<XACT>(fmaf(v.y, sc.y, sh.y));
        v.z = xact_n<XACT>(fmaf(v.z, sc.z, sh.z)); v.w = xact_n<XACT>(fmaf(v.w, sc.w, sh.w));
      }
      *reinterpret_cast<float4*>(sA + r * Kp + 4 * c) = v;
    }
    __syncthreads();
    if (tact) {
#pragma unroll 4
      for (int r = rg; r < rows; r += RG) {
        const float4 g4 = *reinterpret_cast<const float4*>(sG + r * Np + tn);
        const float4 a4 = *reinterpret_cast<const float4*>(sA + r * Kp + tk);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w}, a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(g[i], a[j], acc[i][j]);
      }
    }
  }
  __syncthreads();
  if (tact) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&sW[(tn + i) * Kp + tk + j], acc[i][j]);
  }
  __syncthreads();
  for (int i = tid; i < N * K; i += kNT) {
    const int n = i / K, k = i - n * K;
    atomicAdd(dW + i, sW[n * Kp + k]);
  }
}

}  // namespace

// returns EAT_ERR_UNSUPPORTED (without setting an error) when the shape is outside this kernel's range
int wgrad_narrow_launch(const float* G, const float* A, float* dW, long long M, int N, int K, const float* in_scale,
                        const float* in_shift, int in_act, cudaStream_t st) {
  if (N % 4 != 0 || K % 4 != 0 || N * K > 512 || M < 65536) return EAT_ERR_UNSUPPORTED;
  const int tiles = (N / 4) * (K / 4);
  int wpg = (tiles + 31) / 32;                 // warps per row group: 1, 2, 4 or 8
  if (wpg > 8) return EAT_ERR_UNSUPPORTED;
  if (wpg == 3) wpg = 4;
  if (wpg > 4 && wpg < 8) wpg = 8;
  const size_t smem = ((size_t)kSlab * (N + K) + (size_t)N * K) * sizeof(float);
  const long long slabs = (M + kSlab - 1) / kSlab;
  int grid = (int)(slabs < 148 * 4 ? slabs : 148 * 4);
  const int xact = in_scale != nullptr ? in_act : -1;
  switch (xact) {
    case -1: wgrad_narrow_kernel<-1><<<grid, kNT, smem, st>>>(G, A, dW, M, N, K, in_scale, in_shift, wpg); break;
    case EAT_ACT_NONE: wgrad_narrow_kernel<EAT_ACT_NONE><<<grid, kNT, smem, st>>>(G, A, dW, M, N, K, in_scale, in_shift, wpg); break;
    case EAT_ACT_RELU: wgrad_narrow_kernel<EAT_ACT_RELU><<<grid, kNT, smem, st>>>(G, A, dW, M, N, K, in_scale, in_shift, wpg); break;
    case EAT_ACT_HSWISH: wgrad_narrow_kernel<EAT_ACT_HSWISH><<<grid, kNT, smem, st>>>(G, A, dW, M, N, K, in_scale, in_shift, wpg); break;
    default: return EAT_ERR_UNSUPPORTED;
  }
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}
