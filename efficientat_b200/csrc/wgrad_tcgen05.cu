// Weight gradient of a pointwise (1x1) convolution as a tcgen05 GEMM whose reduction runs over pixels:
//     dW[N, K] += G[M, N]^T . xf(A)[M, K]            G: gradient rows (NHWC), A: saved layer input, dW fp32
// (autograd of the 1x1 ConvNormActivation layers, reference models/mn/block_types.py:140-147,167-171,
// reached from ex_audioset.py:197 loss.backward()).
//
// Both operands are "MN-major" for the tensor core: the reduction index m is the slow (row) index in memory.
// Producers therefore copy 16-byte chunks straight into the canonical MN-major SWIZZLE_128B layout
// (8 m-rows x 128 B atoms; LBO = distance between 64-element atoms along N / K, SBO = distance between
// 8-row groups along m) and the instruction descriptor sets a_major = b_major = MN -- no transposition.
// Each CTA owns one (128 x KT) tile of dW and one slice of the M range, accumulates it in TMEM and adds it to
// dW with vector atomics (dW is zeroed by the caller once per step).
// fp32 operands are split hi/lo into bf16 pairs (3 MMAs), as in pw_tcgen05.cu.
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int WM = 128;                // dW rows per CTA (output channels)
constexpr int MB = 64;                 // reduction rows per pipeline stage
constexpr int kGroupThreads = 128;
constexpr int kThreads = 416;          // 8 producer warps + 1 MMA warp + 4 epilogue warps
constexpr int kMmaWarp = 8;
constexpr int kFirstEpiWarp = 9;
constexpr int STG_LD = 36;
constexpr int G_ATOMS = 2;             // 128 output channels = 2 atoms of 64

struct WgParams {
  const void* G;
  const void* A;
  float* dW;
  int M, N, K;
  InXform xf;
  int KT;                 // dW columns per CTA (multiple of 16, <= KT_MAX)
  int n_tiles, k_tiles, rows_per_split;
  int sample_rows;        // > 0: per-sample gradients dW[b] (DynamicConv); splits never straddle samples
  int splits_per_sample;
};

// MN-major SWIZZLE_128B descriptor: LBO between 64-element atoms along MN, SBO between 8-row groups along K
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc_mn(int n) {       // D fp32, A/B bf16, both MN-major, M = 128
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(WM >> 4) << 24);
}
// byte offset of the 16-byte chunk (row r of the stage, chunk c16 along MN) in a tile with `atoms` atoms per row group
__device__ __forceinline__ uint32_t swz_mn(int r, int c16, int atoms) {
  const int rr = r & 7;
  return (uint32_t)(((r >> 3) * atoms + (c16 >> 3)) * 1024 + rr * 128 + (((c16 & 7) ^ rr) << 4));
}

template <typename T, int NP, int STAGES, int KT_MAX>
__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(WgParams p) {
  constexpr int A_ATOMS = KT_MAX / 64;
  constexpr int G_TILE = MB * G_ATOMS * 128;        // 16 KB
  constexpr int A_TILE = MB * A_ATOMS * 128;        // 32 KB at KT_MAX = 256
  constexpr int STAGE_BYTES = NP * (G_TILE + A_TILE);
  extern __shared__ __align__(1024) unsigned char smem[];
  float* s_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);      // 4 warps x 32 x STG_LD
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + 4 * 32 * STG_LD);     // full[S], empty[S], tfull
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES), bar_tfull = smem_u32(bars + 2 * STAGES);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, kGroupThreads / 32); mbar_init(bar_empty + 8 * s, 1); }   // one arrival per warp
    mbar_init(bar_tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(KT_MAX));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int tiles = p.n_tiles * p.k_tiles;
  const int ot = blockIdx.x % tiles, split = blockIdx.x / tiles;
  const int nt = ot / p.k_tiles, kt = ot - nt * p.k_tiles;
  const int n0 = nt * WM, k0 = kt * p.KT;
  long long m_begin, m_end;
  float* __restrict__ dWout = p.dW;
  if (p.sample_rows > 0) {
    const int b = split / p.splits_per_sample, j = split - b * p.splits_per_sample;
    m_begin = (long long)b * p.sample_rows + (long long)j * p.rows_per_split;
    m_end = m_begin + p.rows_per_split;
    if (m_end > (long long)(b + 1) * p.sample_rows) m_end = (long long)(b + 1) * p.sample_rows;
    dWout += (size_t)b * p.N * p.K;
  } else {
    m_begin = (long long)split * p.rows_per_split;
    m_end = m_begin + p.rows_per_split;
    if (m_end > p.M) m_end = p.M;
  }
  const int n_blocks = (int)((m_end - m_begin + MB - 1) / MB);
  const int N = p.N, K = p.K;
  const int ncG = (min(WM, N - n0) + 7) >> 3;            // valid 16-byte chunks per G row
  const int ncA = (min(p.KT, K - k0) + 7) >> 3;          // valid 16-byte chunks per A row

  if (warp < 8) {
    // ================================================================= producers
    const int grp = warp >> 2;
    const int gtid = threadIdx.x & (kGroupThreads - 1);
    const T* __restrict__ G = reinterpret_cast<const T*>(p.G);
    const T* __restrict__ A = reinterpret_cast<const T*>(p.A);
    int lgG = 0; while ((1 << lgG) < ncG) ++lgG;
    int lgA = 0; while ((1 << lgA) < ncA) ++lgA;
    const int cG = gtid & ((1 << lgG) - 1), rG0 = gtid >> lgG, rGs = kGroupThreads >> lgG;
    const int cA = gtid & ((1 << lgA) - 1), rA0 = gtid >> lgA, rAs = kGroupThreads >> lgA;
    const bool gact = cG < ncG, aact = cA < ncA;
    const int kA = k0 + cA * 8;
    float isc[8], ish[8];
    if (p.xf.scale != nullptr && aact) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { isc[j] = __ldg(p.xf.scale + kA + j); ish[j] = __ldg(p.xf.shift + kA + j); }
    }
    const int rps = p.xf.rows_per_sample;
    for (int blk = grp; blk < n_blocks; blk += 2) {
      const int stage = blk % STAGES;
      const uint32_t phase = (uint32_t)(blk / STAGES) & 1u;
      mbar_wait(bar_empty + 8 * stage, phase ^ 1);
      unsigned char* sG_hi = smem + stage * STAGE_BYTES;
      unsigned char* sG_lo = sG_hi + G_TILE;
      unsigned char* sA_hi = sG_hi + NP * G_TILE;
      unsigned char* sA_lo = sA_hi + A_TILE;
      const long long mb = m_begin + (long long)blk * MB;
      if (gtid == 0) {
        // L2 prefetch of the rows this group stages two of its blocks from now (whole rows: contiguous ranges)
        for (int ahead = (blk == grp ? 2 : 4); ahead <= 4; ahead += 2) {
          const long long m2 = mb + (long long)ahead * MB;
          if (m2 < m_end) {
            const long long rows2 = m_end - m2 < MB ? m_end - m2 : MB;
            l2_prefetch(G + m2 * N, (uint32_t)(rows2 * N * (long long)sizeof(T)));
            l2_prefetch(A + m2 * K, (uint32_t)(rows2 * K * (long long)sizeof(T)));
          }
        }
      }
      if (gact) {
        for (int rb = rG0; rb < MB; rb += 4 * rGs) {
          float v[4][8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rb + i * rGs;
            const long long m = mb + r;
            if (r < MB && m < m_end) load_chunk<T>(G + m * N + n0 + cG * 8, v[i]);
            else {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rb + i * rGs;
            if (r < MB) store_chunk<NP>(sG_hi, sG_lo, swz_mn(r, cG, G_ATOMS), v[i]);
          }
        }
      }
      if (aact) {
        int b_blk = 0, off_blk = 0;
        if (p.xf.gate != nullptr) { b_blk = (int)(mb / rps); off_blk = (int)(mb - (long long)b_blk * rps); }
        for (int rb = rA0; rb < MB; rb += 4 * rAs) {
          float v[4][8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rb + i * rAs;
            const long long m = mb + r;
            if (r < MB && m < m_end) load_chunk<T>(A + m * K + kA, v[i]);
            else {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rb + i * rAs;
            const long long m = mb + r;
            if (r >= MB) continue;
            if (m < m_end) {
              if (p.xf.scale != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[i][j] = act_fwd(fmaf(v[i][j], isc[j], ish[j]), p.xf.act);
              }
              if (p.xf.gate != nullptr) {
                const int rel = off_blk + r;
                const int bb = b_blk + (rps >= MB ? (rel >= rps ? 1 : 0) : rel / rps);
                const float4* gp = reinterpret_cast<const float4*>(p.xf.gate + (size_t)bb * K + kA);
                const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
                v[i][0] *= g0.x; v[i][1] *= g0.y; v[i][2] *= g0.z; v[i][3] *= g0.w;
                v[i][4] *= g1.x; v[i][5] *= g1.y; v[i][6] *= g1.z; v[i][7] *= g1.w;
              }
            }
            store_chunk<NP>(sA_hi, sA_lo, swz_mn(r, cA, A_ATOMS), v[i]);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * stage);
    }
  } else if (warp == kMmaWarp) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_mn(p.KT);
      constexpr uint32_t G_SBO = G_ATOMS * 1024, A_SBO = A_ATOMS * 1024;
      for (int blk = 0; blk < n_blocks; ++blk) {
        const int stage = blk % STAGES;
        const uint32_t phase = (uint32_t)(blk / STAGES) & 1u;
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sG_hi = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sG_lo = sG_hi + G_TILE;
        const uint32_t sA_hi = sG_hi + NP * G_TILE;
        const uint32_t sA_lo = sA_hi + A_TILE;
        const long long mb = m_begin + (long long)blk * MB;
        const int rows = (int)min((long long)MB, m_end - mb);
        const int steps = (rows + 15) >> 4;
        for (int s = 0; s < steps; ++s) {
          const uint32_t go = s * 2 * G_SBO, ao = s * 2 * A_SBO;      // 16 reduction rows = two 8-row groups
          const uint32_t first = (blk == 0 && s == 0) ? 0u : 1u;
          tc_mma(tmem_base, umma_desc_mn(sG_hi + go, 1024, G_SBO), umma_desc_mn(sA_hi + ao, 1024, A_SBO), idesc, first);
          if (NP == 2) {
            tc_mma(tmem_base, umma_desc_mn(sG_lo + go, 1024, G_SBO), umma_desc_mn(sA_hi + ao, 1024, A_SBO), idesc, 1u);
            tc_mma(tmem_base, umma_desc_mn(sG_hi + go, 1024, G_SBO), umma_desc_mn(sA_lo + ao, 1024, A_SBO), idesc, 1u);
          }
        }
        tc_commit(bar_empty + 8 * stage);
      }
      tc_commit(bar_tfull);
    }
    __syncwarp();
  } else {
    // ================================================================= epilogue: TMEM -> transposed staging -> vector atomics
    const int ew = warp - kFirstEpiWarp;
    const int q = warp & 3;
    float* stg = s_stage + ew * 32 * STG_LD;
    const int col4 = (lane & 7) * 4, rg = lane >> 3;
    if (n_blocks > 0) {
      mbar_wait(bar_tfull, 0);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c = 0; c * 32 < p.KT; ++c) {
        uint32_t raw[32];
        tc_ld32(trow + c * 32, raw);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stg + lane * STG_LD + 4 * j) =
              make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                          __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
        __syncwarp();
        const int k = k0 + c * 32 + col4;
        if (c * 32 + col4 < p.KT && k < K) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + rg;
            const int n = n0 + q * 32 + row;
            if (n < N) {
              const float4 v = *reinterpret_cast<const float4*>(stg + row * STG_LD + col4);
              atomicAdd(reinterpret_cast<float4*>(dWout + (size_t)n * K + k), v);
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(KT_MAX));
  }
}

template <typename T, int NP, int STAGES, int KT_MAX>
int launch_wg(const WgParams& p0, cudaStream_t st) {
  WgParams p = p0;
  p.n_tiles = ceil_div(p.N, WM);
  p.k_tiles = ceil_div(p.K, KT_MAX);
  p.KT = ceil_div(ceil_div(p.K, p.k_tiles), 16) * 16;
  p.k_tiles = ceil_div(p.K, p.KT);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = p.n_tiles * p.k_tiles;
  int splits;
  if (p.sample_rows > 0) {
    const int B = p.M / p.sample_rows;
    int sps = max(1, (2 * sms) / max(1, tiles * B));
    long long rows = ceil_div_ll(p.sample_rows, sps);
    rows = ceil_div_ll(rows, MB) * MB;
    sps = (int)ceil_div_ll(p.sample_rows, rows);
    p.rows_per_split = (int)rows;
    p.splits_per_sample = sps;
    splits = sps * B;
  } else {
    splits = max(1, (2 * sms) / tiles);
    long long rows = ceil_div_ll(p.M, splits);
    rows = ceil_div_ll(rows, MB) * MB;
    if (rows < 4 * MB) rows = 4 * MB;
    splits = (int)ceil_div_ll(p.M, rows);
    p.rows_per_split = (int)rows;
    p.splits_per_sample = 0;
  }
  constexpr size_t smem = (size_t)STAGES * NP * (MB * G_ATOMS * 128 + MB * (KT_MAX / 64) * 128) +
                          4 * 32 * STG_LD * sizeof(float) + (2 * STAGES + 1) * sizeof(uint64_t) + 16;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static unsigned long long attr_mask = 0;
  if (int rc = eat_opt_in_smem(wgrad_tc_kernel<T, NP, STAGES, KT_MAX>, smem, attr_mask)) return rc;
  wgrad_tc_kernel<T, NP, STAGES, KT_MAX><<<tiles * splits, kThreads, smem, st>>>(p);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

}  // namespace

static bool use_tma_wgrad() {       // fp32 storage: the TMA-fed kernel (wgrad_tma.cu) unless EAT_WG_IMPL=tc
  static const bool v = [] { const char* e = getenv("EAT_WG_IMPL"); return e == nullptr || strcmp(e, "tc") != 0; }();
  return v;
}

extern "C" int eat_pw_tc_wgrad(const void* G, int g_dtype, const void* A, int a_dtype, float* dW, float* db, long long M,
                               int N, int K, const float* in_scale, const float* in_shift, int in_act,
                               const float* gate, int rows_per_sample, cudaStream_t st) {
  if (M == 0) return EAT_OK;
  if (db != nullptr) { eat_set_error("pw_tc_wgrad: bias gradients are not produced by this kernel"); return EAT_ERR_UNSUPPORTED; }
  if (g_dtype != a_dtype) { eat_set_error("pw_tc_wgrad: G and A must share the storage dtype"); return EAT_ERR_UNSUPPORTED; }
  if (K % 8 != 0 || N % 8 != 0) { eat_set_error("pw_tc_wgrad: K and N must be multiples of 8"); return EAT_ERR_ARG; }
  if (M >= (1ll << 31) - MB) { eat_set_error("pw_tc_wgrad: M too large"); return EAT_ERR_ARG; }
  if ((((uintptr_t)G) | ((uintptr_t)A) | ((uintptr_t)dW)) & 15) { eat_set_error("pw_tc_wgrad: operands must be 16-byte aligned"); return EAT_ERR_ARG; }
  if (K % 4 != 0) { eat_set_error("pw_tc_wgrad: K must be a multiple of 4"); return EAT_ERR_ARG; }
  if (a_dtype == EAT_F32 && use_tma_wgrad() && getenv("EAT_WG_NARROW") == nullptr)
    return eat_pw_tma_wgrad((const float*)G, (const float*)A, dW, M, N, K, in_scale, in_shift, in_act, gate, rows_per_sample, 0, st);
  if (a_dtype == EAT_F32 && gate == nullptr) {
    // narrowest layers (N*K <= 512, huge M): exact-fp32 CUDA-core kernel, see wgrad_narrow.cu for the measurements
    const int rc = wgrad_narrow_launch((const float*)G, (const float*)A, dW, M, N, K, in_scale, in_shift, in_act, st);
    if (rc != EAT_ERR_UNSUPPORTED) return rc;
  }
  if (a_dtype == EAT_F32 && use_tma_wgrad())
    return eat_pw_tma_wgrad((const float*)G, (const float*)A, dW, M, N, K, in_scale, in_shift, in_act, gate, rows_per_sample, 0, st);
  WgParams p;
  p.G = G; p.A = A; p.dW = dW; p.M = (int)M; p.N = N; p.K = K;
  p.xf = InXform{in_scale, in_shift, gate, in_act, rows_per_sample > 0 ? rows_per_sample : 1};
  p.sample_rows = 0;
  if (a_dtype == EAT_BF16) return launch_wg<__nv_bfloat16, 1, 4, 256>(p, st);
  return launch_wg<float, 2, 2, 256>(p, st);
}

// Per-sample weight gradients S[b] = G_b^T . A_b  (S: [B, N, K] fp32, zeroed by the caller); the DynamicConv bank and
// attention gradients follow from S with eat_dyn_wgrad_mix (reference models/dymn/dy_block.py:103-131 backward).
extern "C" int eat_pw_tc_wgrad_persample(const void* G, const void* A, int dtype, float* S, long long M, int N, int K,
                                         int rows_per_sample, cudaStream_t st) {
  if (M == 0) return EAT_OK;
  if (K % 8 != 0 || N % 8 != 0) { eat_set_error("pw_tc_wgrad_persample: K and N must be multiples of 8"); return EAT_ERR_ARG; }
  if (rows_per_sample < 1 || M % rows_per_sample != 0) { eat_set_error("pw_tc_wgrad_persample: M must be B * rows_per_sample"); return EAT_ERR_ARG; }
  if (M >= (1ll << 31) - MB) { eat_set_error("pw_tc_wgrad_persample: M too large"); return EAT_ERR_ARG; }
  if (dtype == EAT_F32 && use_tma_wgrad())
    return eat_pw_tma_wgrad((const float*)G, (const float*)A, S, M, N, K, nullptr, nullptr, 0, nullptr, rows_per_sample, 1, st);
  WgParams p;
  p.G = G; p.A = A; p.dW = S; p.M = (int)M; p.N = N; p.K = K;
  p.xf = InXform{nullptr, nullptr, nullptr, 0, rows_per_sample};
  p.sample_rows = rows_per_sample;
  if (dtype == EAT_BF16) return launch_wg<__nv_bfloat16, 1, 4, 256>(p, st);
  return launch_wg<float, 2, 2, 256>(p, st);
}
