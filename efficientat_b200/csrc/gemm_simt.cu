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

// element-wise guarded load for rows whose length is not a multiple of 8 (classifier: 527 classes)
template <typename TA>
__device__ __forceinline__ void load8_guard(const TA* row, int k, int K, float (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (k + i < K) ? to_f32<TA>(row[k + i]) : 0.f;
}

// BT == false: W is [N, K] (forward / nn.Linear layout).  BT == true: W is [K, N] (the same weight
// tensor used for the data gradient: dA[M, Cin] = G[M, Cout] . W[Cout, Cin]).
template <typename TA, typename TC, bool BT>
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

  float av[8], bv[4];
  // global -> registers for the k-block at k0 (input transform applied here); issued one block ahead of the math
  auto fetch = [&](int k0) {
    const int ka = k0 + a_kc;
    if (a_ok && ka < K) {
      if ((K & 7) == 0) load8<TA>(A + a_m * K + ka, av);
      else load8_guard<TA>(A + a_m * K, ka, K, av);
      if (xf.scale != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (ka + i < K) av[i] = act_fwd(fmaf(av[i], __ldg(xf.scale + ka + i), __ldg(xf.shift + ka + i)), xf.act);
      }
      if (gate_row != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (ka + i < K) av[i] *= __ldg(gate_row + ka + i);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = 0.f;
    }
    bv[0] = bv[1] = bv[2] = bv[3] = 0.f;
    if (!BT) {
      const int kb = k0 + b_kc;
      if (n0 + b_row < N && kb < K) {
        const float* wr = W + (size_t)(n0 + b_row) * K;
        if ((K & 3) == 0) {
          float4 t = __ldg(reinterpret_cast<const float4*>(wr + kb));
          bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) bv[i] = kb + i < K ? __ldg(wr + kb + i) : 0.f;
        }
      }
    } else {
      const int kk = tid >> 4, nn = (tid & 15) * 4;      // 16 k-rows x 64 columns, float4 along n
      if (k0 + kk < K) {
        const float* wr = W + (size_t)(k0 + kk) * N + n0 + nn;
        if ((N & 3) == 0 && n0 + nn + 3 < N) {
          float4 t = __ldg(reinterpret_cast<const float4*>(wr));
          bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) bv[i] = n0 + nn + i < N ? __ldg(wr + i) : 0.f;
        }
      }
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) As[a_kc + i][a_row] = av[i];
    if (!BT) {
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[b_kc + i][b_row] = bv[i];
    } else {
      *reinterpret_cast<float4*>(&Bs[tid >> 4][(tid & 15) * 4]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    }
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);          // the next block's loads fly while this block is multiplied
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
int launch_gemm(const void* A, const float* W, int w_trans, void* C, long long M, int N, int K, InXform xf,
                const float* scale, const float* shift, int act, const void* residual, double* ssum, double* ssq,
                cudaStream_t st) {
  dim3 grid((unsigned)ceil_div_ll(M, BM), (unsigned)ceil_div(N, BN));
  if (w_trans)
    gemm_nt_kernel<TA, TC, true><<<grid, kThreads, 0, st>>>((const TA*)A, W, (TC*)C, (int)M, N, K, xf, scale, shift,
                                                            act, (const TC*)residual, ssum, ssq);
  else
    gemm_nt_kernel<TA, TC, false><<<grid, kThreads, 0, st>>>((const TA*)A, W, (TC*)C, (int)M, N, K, xf, scale, shift,
                                                             act, (const TC*)residual, ssum, ssq);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

// ------------------------------------------------------------------------------------------
// Small-M variant (M <= 1024 rows: the squeeze-excitation MLPs and the classifier, M = batch).  With 128 x 64 tiles those
// GEMMs ran on 2-40 CTAs and were pure latency (53 us per launch on average, 1 ms per training step); 32 x 32 tiles and
// a 32-deep k-block give 8-16x more CTAs.  Same contract as gemm_nt_kernel (fp32 in / out, no statistics).
template <bool BT>
__global__ void __launch_bounds__(256) gemm_small_kernel(
    const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ C, int M, int N, int K, InXform xf,
    const float* __restrict__ scale, const float* __restrict__ shift, int act, const float* __restrict__ residual,
    float alpha) {
  __shared__ float As[32][33];       // [row][k]
  __shared__ float Bs[32][36];       // [k][n], rows 16-byte aligned
  const int tid = threadIdx.x;
  const int lr = tid >> 3, lc = (tid & 7) * 4;
  const int m0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int m = m0 + lr;
  const float* gate_row = (xf.gate != nullptr && m < M) ? xf.gate + (size_t)(m / xf.rows_per_sample) * K : nullptr;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
    float av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + lc + i;
      float v = (m < M && k < K) ? __ldg(A + (size_t)m * K + k) : 0.f;
      if (m < M && k < K) {
        if (xf.scale != nullptr) v = act_fwd(fmaf(v, __ldg(xf.scale + k), __ldg(xf.shift + k)), xf.act);
        if (gate_row != nullptr) v *= __ldg(gate_row + k);
      }
      av[i] = v;
      if (!BT) { const int n = n0 + lr; bv[i] = (n < N && k < K) ? __ldg(W + (size_t)n * K + k) : 0.f; }      // W [N, K]
      else { const int kk = k0 + lr, n = n0 + lc + i; bv[i] = (kk < K && n < N) ? __ldg(W + (size_t)kk * N + n) : 0.f; }   // W [K, N]
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lr][lc + i] = av[i];
      if (!BT) Bs[lc + i][lr] = bv[i]; else Bs[lr][lc + i] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float a = As[lr][kk];
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][lc]);
      acc[0] = fmaf(a, b.x, acc[0]); acc[1] = fmaf(a, b.y, acc[1]); acc[2] = fmaf(a, b.z, acc[2]); acc[3] = fmaf(a, b.w, acc[3]);
    }
  }
  if (m >= M) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + lc + j;
    if (n >= N) continue;
    float v = acc[j] * alpha;
    if (scale != nullptr) v *= __ldg(scale + n);
    if (shift != nullptr) v += __ldg(shift + n);
    v = act_fwd(v, act);
    if (residual != nullptr) v += residual[(size_t)m * N + n];
    C[(size_t)m * N + n] = v;
  }
}

// ------------------------------------------------------------------------------------------
// Weight gradient: dW[N,K] += G[M,N]^T . xf(A)[M,K], db[N] += colsum(G).  64x64 output tile per CTA,
// the M reduction is split over gridDim.z chunks and combined with fp32 atomics (dW is zeroed by the caller).
constexpr int WN = 64, WK = 64, WM = 16;

template <typename TG, typename TA>
__global__ void __launch_bounds__(kThreads) gemm_wgrad_kernel(
    const TG* __restrict__ G, const TA* __restrict__ A, float* __restrict__ dW, float* __restrict__ db, int M, int N,
    int K, InXform xf, int rows_per_cta) {
  __shared__ __align__(16) float Gs[WM][WN];
  __shared__ __align__(16) float As[WM][WK];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * WN, k0 = blockIdx.y * WK;
  const long long m_begin = (long long)blockIdx.z * rows_per_cta;
  long long m_end = m_begin + rows_per_cta;
  if (m_end > M) m_end = M;
  const int lr = tid >> 4, lc = (tid & 15) * 4;     // loader: row lr of the 16-row slab, 4 columns from lc
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bias = db != nullptr && blockIdx.y == 0 && tx == 0;
  for (long long m0 = m_begin; m0 < m_end; m0 += WM) {
    const long long m = m0 + lr;
    float gv[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < m_end) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = n0 + lc + i, k = k0 + lc + i;
        if (n < N) gv[i] = to_f32<TG>(G[m * N + n]);
        if (k < K) {
          float v = to_f32<TA>(A[m * K + k]);
          if (xf.scale != nullptr) v = act_fwd(fmaf(v, __ldg(xf.scale + k), __ldg(xf.shift + k)), xf.act);
          if (xf.gate != nullptr) v *= __ldg(xf.gate + (m / xf.rows_per_sample) * K + k);
          av[i] = v;
        }
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Gs[lr][lc]) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    *reinterpret_cast<float4*>(&As[lr][lc]) = make_float4(av[0], av[1], av[2], av[3]);
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < WM; ++mm) {
      float4 g4 = *reinterpret_cast<const float4*>(&Gs[mm][ty * 4]);
      float4 a4 = *reinterpret_cast<const float4*>(&As[mm][tx * 4]);
      float g[4] = {g4.x, g4.y, g4.z, g4.w}, a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(g[i], a[j], acc[i][j]);
        if (do_bias) bacc[i] += g[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dW + (size_t)n * K + k, acc[i][j]);
    }
    if (do_bias) atomicAdd(db + n, bacc[i]);
  }
}

template <typename TG, typename TA>
int launch_wgrad(const void* G, const void* A, float* dW, float* db, long long M, int N, int K, InXform xf,
                 cudaStream_t st) {
  const int nt = ceil_div(N, WN), kt = ceil_div(K, WK);
  long long want = (148ll * 4) / ((long long)nt * kt);          // enough CTAs to fill the GPU
  if (want < 1) want = 1;
  long long rows = ceil_div_ll(M, want);
  rows = ceil_div_ll(rows, WM) * WM;
  if (rows < 4 * WM) rows = 4 * WM;
  const int chunks = (int)ceil_div_ll(M, rows);
  dim3 grid(nt, kt, chunks);
  gemm_wgrad_kernel<TG, TA><<<grid, kThreads, 0, st>>>((const TG*)G, (const TA*)A, dW, db, (int)M, N, K, xf, (int)rows);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

}  // namespace

// C[M, N] = alpha * A[M, K] . W[K, N]   (fp32, small M; used by the squeeze-excitation MLP backward, bwd_kernels.cu)
int gemm_small_kn_launch(const float* A, const float* W, float* C, int M, int N, int K, float alpha, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(M, 32), (unsigned)ceil_div(N, 32));
  InXform xf{nullptr, nullptr, nullptr, 0, 1};
  gemm_small_kernel<true><<<grid, 256, 0, st>>>(A, W, C, M, N, K, xf, nullptr, nullptr, 0, nullptr, alpha);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

extern "C" int eat_gemm_simt_fwd(const void* A, int a_dtype, const float* W, int w_trans, void* C, int c_dtype,
                                 long long M, int N, int K, const float* in_scale, const float* in_shift, int in_act,
                                 const float* gate, int rows_per_sample, const float* scale, const float* shift,
                                 int act, const void* residual, double* stat_sum, double* stat_sq,
                                 cudaStream_t st) {
  if (M == 0) return EAT_OK;
  if (K % 8 != 0 && a_dtype != EAT_F32) { eat_set_error("gemm: K must be a multiple of 8 for bf16 operands"); return EAT_ERR_ARG; }
  if (M >= (1ll << 31)) { eat_set_error("gemm: M too large"); return EAT_ERR_ARG; }
  InXform xf{in_scale, in_shift, gate, in_act, rows_per_sample > 0 ? rows_per_sample : 1};
  if (a_dtype == EAT_F32 && c_dtype == EAT_F32 && M <= 1024 && stat_sum == nullptr) {
    dim3 grid((unsigned)ceil_div_ll(M, 32), (unsigned)ceil_div(N, 32));
    if (w_trans) gemm_small_kernel<true><<<grid, 256, 0, st>>>((const float*)A, W, (float*)C, (int)M, N, K, xf, scale, shift, act, (const float*)residual, 1.f);
    else gemm_small_kernel<false><<<grid, 256, 0, st>>>((const float*)A, W, (float*)C, (int)M, N, K, xf, scale, shift, act, (const float*)residual, 1.f);
    EAT_CHECK_LAUNCH();
    return EAT_OK;
  }
  if (a_dtype == EAT_F32 && c_dtype == EAT_F32)
    return launch_gemm<float, float>(A, W, w_trans, C, M, N, K, xf, scale, shift, act, residual, stat_sum, stat_sq, st);
  if (a_dtype == EAT_BF16 && c_dtype == EAT_BF16)
    return launch_gemm<__nv_bfloat16, __nv_bfloat16>(A, W, w_trans, C, M, N, K, xf, scale, shift, act, residual, stat_sum, stat_sq, st);
  if (a_dtype == EAT_BF16 && c_dtype == EAT_F32)
    return launch_gemm<__nv_bfloat16, float>(A, W, w_trans, C, M, N, K, xf, scale, shift, act, residual, stat_sum, stat_sq, st);
  if (a_dtype == EAT_F32 && c_dtype == EAT_BF16)
    return launch_gemm<float, __nv_bfloat16>(A, W, w_trans, C, M, N, K, xf, scale, shift, act, residual, stat_sum, stat_sq, st);
  eat_set_error("gemm: unsupported dtype combination");
  return EAT_ERR_UNSUPPORTED;
}

extern "C" int eat_gemm_simt_wgrad(const void* G, int g_dtype, const void* A, int a_dtype, float* dW, float* db,
                                   long long M, int N, int K, const float* in_scale, const float* in_shift,
                                   int in_act, const float* gate, int rows_per_sample, cudaStream_t st) {
  if (M == 0) return EAT_OK;
  if (M >= (1ll << 31)) { eat_set_error("wgrad: M too large"); return EAT_ERR_ARG; }
  InXform xf{in_scale, in_shift, gate, in_act, rows_per_sample > 0 ? rows_per_sample : 1};
  if (g_dtype == EAT_F32 && a_dtype == EAT_F32) return launch_wgrad<float, float>(G, A, dW, db, M, N, K, xf, st);
  if (g_dtype == EAT_BF16 && a_dtype == EAT_BF16)
    return launch_wgrad<__nv_bfloat16, __nv_bfloat16>(G, A, dW, db, M, N, K, xf, st);
  if (g_dtype == EAT_F32 && a_dtype == EAT_BF16) return launch_wgrad<float, __nv_bfloat16>(G, A, dW, db, M, N, K, xf, st);
  if (g_dtype == EAT_BF16 && a_dtype == EAT_F32) return launch_wgrad<__nv_bfloat16, float>(G, A, dW, db, M, N, K, xf, st);
  eat_set_error("wgrad: unsupported dtype combination");
  return EAT_ERR_UNSUPPORTED;
}
