// CUDA-core (FFMA, exact fp32) GEMMs for the pointwise-conv / Linear family.  This is the
// numerically exact path used (a) for the tiny head matmuls, (b) as the on-device cross-check of
// the tcgen05 kernel in pw_tcgen05.cu, and (c) for the backward weight-gradient reductions.
//   fwd :  C[M,N] = epi( xf(A)[M,K] . W[N,K]^T )          (1x1 conv on NHWC rows; nn.Linear)
//   wgrad: dW[N,K] += G[M,N]^T . xf(A)[M,K] ;  db[N] += colsum(G)
// Reference call sites: 1x1 ConvNormActivation in models/mn/block_types.py:140-147,167-171,
// classifier Linear layers models/mn/model.py:187-194.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, kThreads = 256;

template <typename TA>
__device__ __forceinline__ void load8(const TA* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  Vec<__nv_bfloat16>::load(p, v);
}

template <typename TA, typename TC>
__global__ void __launch_bounds__(kThreads) gemm_nt_kernel(
    const TA* __restrict__ A, const float* __restrict__ W, TC* __restrict__ C, int M, int N, int K,
    InXform xf, const float* __restrict__ scale, const float* __restrict__ shift, int act,
    const TC* __restrict__ residual, double* __restrict__ stat_sum, double* __restrict__ stat_sq) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ float s_sum[BN], s_sq[BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  if (stat_sum != nullptr && tid < BN) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }

  const int a_row = tid >> 1, a_kc = (tid & 1) * 8;
  const int b_row = tid >> 2, b_kc = (tid & 3) * 4;
  const long long a_m = m0 + a_row;
  const bool a_ok = a_m < M;
  const float* gate_row = nullptr;
  if (xf.gate != nullptr && a_ok) gate_row = xf.gate + (a_m / xf.rows_per_sample) * K;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    float av[8];
    const int ka = k0 + a_kc;
    if (a_ok && ka < K) {
      load8<TA>(A + a_m * K + ka, av);
      if (xf.scale != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = act_fwd(fmaf(av[i], __ldg(xf.scale + ka + i), __ldg(xf.shift + ka + i)), xf.act);
      }
      if (gate_row != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] *= __ldg(gate_row + ka + i);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = 0.f;
    }
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    const int kb = k0 + b_kc;
    if (n0 + b_row < N && kb < K) {
      float4 t = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + b_row) * K + kb));
      bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) As[a_kc + i][a_row] = av[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[b_kc + i][b_row] = bv[i];
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
  }

  // ---- epilogue
  float csum[4] = {0.f, 0.f, 0.f, 0.f}, csq[4] = {0.f, 0.f, 0.f, 0.f};
  const int nb = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = nb + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (stat_sum != nullptr) { csum[j] += v; csq[j] = fmaf(v, v, csq[j]); }
      if (scale != nullptr) v *= __ldg(scale + n);
      if (shift != nullptr) v += __ldg(shift + n);
      v = act_fwd(v, act);
      if (residual != nullptr) v += to_f32<TC>(residual[m * N + n]);
      C[m * N + n] = from_f32<TC>(v);
    }
  }
  if (stat_sum != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(&s_sum[tx * 4 + j], csum[j]); atomicAdd(&s_sq[tx * 4 + j], csq[j]); }
    __syncthreads();
    if (tid < BN && n0 + tid < N) {
      atomicAdd(stat_sum + n0 + tid, (double)s_sum[tid]);
      atomicAdd(stat_sq + n0 + tid, (double)s_sq[tid]);
    }
  }
}

template <typename TA, typename TC>
int launch_gemm(const void* A, const float* W, void* C, long long M, int N, int K, InXform xf, const float* scale,
                const float* shift, int act, const void* residual, double* ssum, double* ssq, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div_ll(M, BM), (unsigned)ceil_div(N, BN));
  gemm_nt_kernel<TA, TC><<<grid, kThreads, 0, st>>>((const TA*)A, W, (TC*)C, (int)M, N, K, xf, scale, shift, act,
                                                    (const TC*)residual, ssum, ssq);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

}  // namespace

extern "C" int eat_gemm_simt_fwd(const void* A, int a_dtype, const float* W, void* C, int c_dtype, long long M, int N,
                                 int K, const float* in_scale, const float* in_shift, int in_act,
                                 const float* gate, int rows_per_sample, const float* scale, const float* shift,
                                 int act, const void* residual, double* stat_sum, double* stat_sq,
                                 cudaStream_t st) {
  if (M == 0) return EAT_OK;
  if (K % 8 != 0) { eat_set_error("gemm: K must be a multiple of 8"); return EAT_ERR_ARG; }
  if (M >= (1ll << 31)) { eat_set_error("gemm: M too large"); return EAT_ERR_ARG; }
  InXform xf{in_scale, in_shift, gate, in_act, rows_per_sample > 0 ? rows_per_sample : 1};
  if (a_dtype == EAT_F32 && c_dtype == EAT_F32)
    return launch_gemm<float, float>(A, W, C, M, N, K, xf, scale, shift, act, residual, stat_sum, stat_sq, st);
  if (a_dtype == EAT_BF16 && c_dtype == EAT_BF16)
    return launch_gemm<__nv_bfloat16, __nv_bfloat16>(A, W, C, M, N, K, xf, scale, shift, act, residual, stat_sum, stat_sq, st);
  if (a_dtype == EAT_BF16 && c_dtype == EAT_F32)
    return launch_gemm<__nv_bfloat16, float>(A, W, C, M, N, K, xf, scale, shift, act, residual, stat_sum, stat_sq, st);
  if (a_dtype == EAT_F32 && c_dtype == EAT_BF16)
    return launch_gemm<float, __nv_bfloat16>(A, W, C, M, N, K, xf, scale, shift, act, residual, stat_sum, stat_sq, st);
  eat_set_error("gemm: unsupported dtype combination");
  return EAT_ERR_UNSUPPORTED;
}
