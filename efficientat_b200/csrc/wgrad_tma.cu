// Weight gradient of a pointwise (1x1) convolution for fp32 activations as a TMA-fed tcgen05 GEMM whose reduction runs
// over pixels:     dW[N, K] += G[M, N]^T . xf(X)[M, K]        G: gradient rows (NHWC), X: saved layer input, dW fp32
// (autograd of the 1x1 ConvNormActivation layers, reference models/mn/block_types.py:140-147,167-171, reached from
// ex_audioset.py:197 loss.backward()).  Successor of wgrad_tcgen05.cu for fp32 storage, built like pw_tma.cu:
//
//   warp 0   TMA producer : one thread; per 64-row block of the reduction, cp.async.bulk.tensor lands [64 rows x 32
//                           channels] fp32 boxes of G (2 boxes = 64 output channels) and of X (up to 4 boxes = 128
//                           input channels) in a ring of stages; rows past M are zero-filled by the TMA unit.
//   warps 2-5 fix-up      : the in-place pass of tma_common.cuh: (BatchNorm affine + activation + SE gate on X), then a
//                           row's 32 fp32 values become 32 bf16 hi + 32 bf16 lo values in the same 128 bytes.
//                           Both operands are "MN-major" for the tensor core (the reduction index m is the row index),
//                           and a landed box IS a canonical MN-major SWIZZLE_128B atom column: 8-row groups 1024 B
//                           apart (SBO), boxes 8 KB apart (LBO) -- no transposition, no second buffer.
//   warp 1   MMA issuer   : tcgen05.mma kind::f16, M' = 128, N' = 64 * boxes(X), K = 16 reduction rows per instruction.
//                           Because hi and lo sit side by side along the NON-reduced dimension, the accumulator holds
//                           the four products separately: rows n' = [hi(n) | lo(n)], columns k' = [hi(k) | lo(k)].
//   warps 6-9 epilogue    : once per CTA: dW[n,k] += hi.hi + hi.lo (warps on hi(n) lanes) and += lo.hi (warps on lo(n)
//                           lanes), read from TMEM and added with vector atomics (lo.lo ~ 2^-32 is dropped).
// Each CTA owns a (64 x 128) tile of dW and one slice of the M range (dW is zeroed by the caller once per step); two
// CTAs share an SM (256 TMEM columns and <= 113 KB of shared memory each).
// HBM-bound: algorithmic bytes per launch = 4 * (M*N + M*K + N*K).
#include <cstdlib>

#include "tma_common.cuh"

namespace {
using namespace tc;
using namespace tma;

// MB = reduction rows per pipeline stage (template parameter): 64, or 128 for launches with few boxes per stage, where the
// per-block costs (barrier round trips, TMA issue, fix-up prologue) rather than bytes set the pace.
// BOX = MB * 128 bytes: one landed [MB x 32 fp32] box.
constexpr int GB = 2;                   // G boxes per stage: 64 output channels -> 128 accumulator rows
constexpr int kThreads = 320;
constexpr int kMmaWarp = 1, kFirstFix = 2, kFirstEpi = 6;

struct WgParams {
  float* dW;
  int M, N, K;
  int n_tiles, k_tiles, xb;             // xb: X boxes reserved per stage (<= 4 = 128 channels); a CTA loads only those inside K
  int gbs;                              // G boxes reserved per stage (1 when N <= 32, else 2); a CTA loads only those inside N
  int rows_per_split, sample_rows, splits_per_sample;
  int stages;
  uint32_t stage_bytes, off_f, off_bar;
  const float* in_scale; const float* in_shift; const float* gate; int in_act; int rps;
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
__device__ __forceinline__ uint32_t idesc_mn(int n) {       // D fp32, A/B bf16, both MN-major, M = 128
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int XACT, int MB>
__global__ void __launch_bounds__(kThreads, 2)
wgrad_tma_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapX, const WgParams p) {
  constexpr int BOX = MB * 128;
  extern __shared__ __align__(1024) unsigned char smem[];
  float* s_isc = reinterpret_cast<float*>(smem + p.off_f);          // [128] in-transform scale of this CTA's channels
  float* s_ish = s_isc + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  const int S = p.stages;
  const uint32_t bar_full = smem_u32(bars), bar_ready = bar_full + 8 * S, bar_empty = bar_ready + 8 * S;
  const uint32_t bar_tfull = bar_empty + 8 * S;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 3 * S + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.n_tiles * p.k_tiles;
  const int ot = blockIdx.x % tiles, split = blockIdx.x / tiles;
  const int nt = ot / p.k_tiles, kt = ot - nt * p.k_tiles;
  const int n0 = nt * 64, k0 = kt * 128;
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
  const int n_blocks = m_end > m_begin ? (int)((m_end - m_begin + MB - 1) / MB) : 0;
  // boxes of this CTA's tile that hold real channels.  A box that is not loaded leaves its accumulator rows / columns
  // undefined (the MMA shape stays 128 x 64*xb so that the operand atoms keep their LBO spacing); the epilogue never
  // reads them (n < N, kb < K tests below).
  const int gb = min(p.gbs, (p.N - n0 + KB - 1) / KB);
  const int xb = min(p.xb, (p.K - k0 + KB - 1) / KB);
  const uint32_t x_off = (uint32_t)p.gbs * BOX;
  const uint32_t stage_base = smem_u32(smem);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_ready + 8 * s, 4); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapG)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapX)) : "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (XACT >= 0) {
    for (int i = threadIdx.x; i < 128; i += kThreads) {
      s_isc[i] = k0 + i < p.K ? p.in_scale[k0 + i] : 0.f;          // zero beyond K: act(0) = 0
      s_ish[i] = k0 + i < p.K ? p.in_shift[k0 + i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ================================================================= TMA producer (one thread)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int blk = 0; blk < n_blocks; ++blk) {
        const int m = (int)(m_begin + (long long)blk * MB);
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        const uint32_t dst = stage_base + (uint32_t)s * p.stage_bytes;
        mbar_expect_tx(bar_full + 8 * s, (uint32_t)(gb + xb) * BOX);
        for (int b = 0; b < gb; ++b) tma_load_2d(&mapG, bar_full + 8 * s, dst + b * BOX, n0 + b * KB, m);
        for (int b = 0; b < xb; ++b) tma_load_2d(&mapX, bar_full + 8 * s, dst + x_off + b * BOX, k0 + b * KB, m);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================================================================= MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = idesc_mn(64 * p.xb);
      int s = 0;
      uint32_t ph = 0;
      for (int blk = 0; blk < n_blocks; ++blk) {
        mbar_wait(bar_ready + 8 * s, ph);
        tc_fence_after();
        const uint32_t sg = stage_base + (uint32_t)s * p.stage_bytes, sx = sg + x_off;
        const long long mb = m_begin + (long long)blk * MB;
        const int rows = (int)min((long long)MB, m_end - mb);
        const int steps = (rows + 15) >> 4;
#pragma unroll
        for (int st = 0; st < MB / 16; ++st) {
          if (st < steps)                                           // 16 reduction rows = two 8-row groups = 2048 bytes
            mma_bf16(tmem_base, umma_desc_mn(sg + st * 2048, BOX, 1024), umma_desc_mn(sx + st * 2048, BOX, 1024), idesc,
                     (blk | st) ? 1u : 0u);
        }
        tc_commit(bar_empty + 8 * s);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      tc_commit(bar_tfull);
    }
    __syncwarp();
  } else if (warp < kFirstEpi) {
    // ================================================================= fix-up warps (128 threads)
    const int ft = threadIdx.x - kFirstFix * 32;
    int s = 0;
    uint32_t ph = 0;
    for (int blk = 0; blk < n_blocks; ++blk) {
      const long long mb = m_begin + (long long)blk * MB;
      const int rows_valid = (int)min((long long)MB, m_end - mb);
      int b0 = 0, off0 = 0;
      if (p.gate != nullptr) { b0 = (int)(mb / p.rps); off0 = (int)(mb - (long long)b0 * p.rps); }
      mbar_wait(bar_full + 8 * s, ph);
      unsigned char* st = smem + (size_t)s * p.stage_bytes;
      for (int b = 0; b < gb; ++b)                               // gradient boxes: plain split (rows past the split zeroed)
        fix_a<2, -1, MB>(st + b * BOX, ft, rows_valid, nullptr, nullptr, 0, nullptr, 0, 0, 1, 0);
      for (int b = 0; b < xb; ++b) {
        const int kl = b * KB + (ft & 3) * 8;                      // channel of this thread's chunk pair, local to the CTA
        fix_a<2, XACT, MB>(st + x_off + b * BOX, ft, rows_valid, s_isc - k0, s_ish - k0, k0 + kl, p.gate, off0, b0, p.rps, p.K);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready + 8 * s);
      if (++s == S) { s = 0; ph ^= 1u; }
    }
  } else {
    // ================================================================= epilogue: once per CTA, TMEM -> vector atomics
    const int q = warp & 3;                    // lane quadrant: 0 hi(n 0..31), 1 lo(n 0..31), 2 hi(n 32..63), 3 lo(n 32..63)
    if (n_blocks > 0) {
      mbar_wait(bar_tfull, 0);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
      const int n = n0 + (q >> 1) * 32 + lane;
      const bool hi_lane = (q & 1) == 0;
      for (int b = 0; b < xb; ++b) {
        const int kb = k0 + b * KB;
        if (kb >= p.K) break;
        uint32_t v[32];
        tc_ld32(trow + b * 64, v);                                  // columns hi(k): hi.hi on hi lanes, lo.hi on lo lanes
        if (hi_lane) {
          uint32_t w[32];
          tc_ld32(trow + b * 64 + 32, w);                           // columns lo(k): hi.lo
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        }
        if (n < p.N) {
          float* dst = dWout + (size_t)n * p.K + kb;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (kb + 4 * j < p.K)
              atomicAdd(reinterpret_cast<float4*>(dst + 4 * j),
                        make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                    __uint_as_float(v[4 * j + 3])));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

template <int XACT, int MB>
int launch_wg_mb(const float* G, const float* X, WgParams p, cudaStream_t st) {
  constexpr int BOX = MB * 128;
  p.n_tiles = ceil_div(p.N, 64);
  p.k_tiles = ceil_div(p.K, 128);
  const int kt = p.K < 128 ? p.K : 128;
  p.xb = ceil_div(kt, KB);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = p.n_tiles * p.k_tiles;
  const int slots = 2 * sms;                     // two CTAs per SM
  int splits;
  if (p.sample_rows > 0) {
    const int B = p.M / p.sample_rows;
    int sps = max(1, (2 * slots) / max(1, tiles * B));
    long long rows = ceil_div_ll(p.sample_rows, sps);
    rows = ceil_div_ll(rows, MB) * MB;
    sps = (int)ceil_div_ll(p.sample_rows, rows);
    p.rows_per_split = (int)rows;
    p.splits_per_sample = sps;
    splits = sps * B;
  } else {
    splits = max(1, (2 * slots) / tiles);
    long long rows = ceil_div_ll(p.M, splits);
    rows = ceil_div_ll(rows, MB) * MB;
    if (rows < 4 * MB) rows = 4 * MB;
    splits = (int)ceil_div_ll(p.M, rows);
    p.rows_per_split = (int)rows;
    p.splits_per_sample = 0;
  }
  p.gbs = p.N <= KB ? 1 : GB;
  p.stage_bytes = (uint32_t)((p.gbs + p.xb) * BOX);
  const size_t budget = (227 * 1024 - 2048) / 2;
  const size_t misc = 2 * 128 * 4 + (3 * 8 + 1) * 8 + 16 + 1024;
  p.stages = (int)((budget - misc) / p.stage_bytes);
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) { eat_set_error("wgrad_tma: shared-memory budget exceeded"); return EAT_ERR_UNSUPPORTED; }
  size_t off = (size_t)p.stages * p.stage_bytes;
  p.off_f = (uint32_t)off; off += 2 * 128 * 4;
  p.off_bar = (uint32_t)off; off += (3 * (size_t)p.stages + 1) * 8 + 16;
  const size_t smem = off;
  CUtensorMap mG, mX;
  if (int rc = make_map(&mG, G, p.M, p.N, MB)) return rc;
  if (int rc = make_map(&mX, X, p.M, p.K, MB)) return rc;
  static unsigned long long attr_mask = 0;
  if (int rc = eat_opt_in_smem(wgrad_tma_kernel<XACT, MB>, budget, attr_mask)) return rc;
  wgrad_tma_kernel<XACT, MB><<<tiles * splits, kThreads, smem, st>>>(mG, mX, p);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

template <int XACT>
int launch_wg(const float* G, const float* X, const WgParams& p, cudaStream_t st) {
  // 128-row blocks when a stage holds at most three boxes (N <= 64 and K <= 32, or N <= 32 and K <= 64) and the
  // reduction is long: the first layers of the network, M in the millions
  const int boxes = (p.N <= KB ? 1 : GB) + ceil_div(p.K < 128 ? p.K : 128, KB);
  int mb = (boxes <= 3 && p.sample_rows == 0 && p.M >= (1 << 20)) ? 128 : 64;
  if (const char* e = getenv("EAT_WG_MB")) { const int v = atoi(e); if (v == 64 || (v == 128 && boxes <= 3)) mb = v; }
  return mb == 128 ? launch_wg_mb<XACT, 128>(G, X, p, st) : launch_wg_mb<XACT, 64>(G, X, p, st);
}

int launch_wg_x(const float* G, const float* X, const WgParams& p, cudaStream_t st) {
  if (p.in_scale == nullptr) return launch_wg<-1>(G, X, p, st);
  if (p.in_act == EAT_ACT_RELU) return launch_wg<1>(G, X, p, st);
  if (p.in_act == EAT_ACT_HSWISH) return launch_wg<2>(G, X, p, st);
  return launch_wg<0>(G, X, p, st);
}

}  // namespace

extern "C" int eat_pw_tma_wgrad(const float* G, const float* X, float* dW, long long M, int N, int K, const float* in_scale,
                                const float* in_shift, int in_act, const float* gate, int rows_per_sample,
                                int per_sample, cudaStream_t st) {
  if (M == 0) return EAT_OK;
  if (K % 4 != 0 || N % 4 != 0) { eat_set_error("pw_tma_wgrad: K and N must be multiples of 4"); return EAT_ERR_ARG; }
  if (M >= (1ll << 31) - 128) { eat_set_error("pw_tma_wgrad: M too large"); return EAT_ERR_ARG; }
  if ((in_scale == nullptr) != (in_shift == nullptr)) { eat_set_error("pw_tma_wgrad: in_scale and in_shift come together"); return EAT_ERR_ARG; }
  if (in_act == EAT_ACT_SIGMOID) { eat_set_error("pw_tma_wgrad: sigmoid input activation is not offered"); return EAT_ERR_UNSUPPORTED; }
  if ((((uintptr_t)G) | ((uintptr_t)X) | ((uintptr_t)dW) | ((uintptr_t)gate)) & 15) { eat_set_error("pw_tma_wgrad: operands must be 16-byte aligned"); return EAT_ERR_ARG; }
  WgParams p{};
  p.dW = dW; p.M = (int)M; p.N = N; p.K = K;
  p.in_scale = in_scale; p.in_shift = in_shift; p.gate = gate; p.in_act = in_act;
  p.rps = rows_per_sample > 0 ? rows_per_sample : 1;
  p.sample_rows = 0;
  if (per_sample) {
    if (rows_per_sample < 1 || M % rows_per_sample != 0) { eat_set_error("pw_tma_wgrad: per-sample mode needs M = B * rows_per_sample"); return EAT_ERR_ARG; }
    p.sample_rows = rows_per_sample;
  }
  return launch_wg_x(G, X, p, st);
}
