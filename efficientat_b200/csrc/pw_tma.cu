// Pointwise (1x1) convolution / Linear for fp32 activations as a TMA-fed tcgen05 GEMM (sm_100a), bf16x3 products:
//     C[M,N] = epi( xf(A)[M,K] . W[N,K]^T  (+ R[M,N]) )          A, C, R: NHWC activation rows, fp32; W: fp32
// Same contract and reference call sites as pw_tcgen05.cu (models/mn/block_types.py:140-147,167-171;
// models/mn/model.py:160-166; every data-gradient GEMM of loss.backward(), ex_audioset.py:197).
//
// Why a second kernel: in pw_tcgen05.cu eight producer warps carry every operand byte global -> registers -> smem;
// their scalar tile walk and register-limited loads (not bytes, not MMAs) set a ~2000-cycle floor per 128-row tile
// (profiles/r01_pw_tc_role_timing.txt).  Here no thread touches global memory on the operand path:
//
//   warp 0   TMA producer : ONE thread.  cp.async.bulk.tensor (2-D tiled, SWIZZLE_128B) lands raw fp32 tiles of A
//                           [128 rows x 32 k] (and of W, and of the residual R) in a ring of shared-memory stages;
//                           ragged M / K / N edges are zero-filled by the TMA unit.
//   warps 2-5 fix-up      : on-chip pass over a landed tile: (BatchNorm affine + activation + SE gate of the producing
//                           layer for training-mode operands), then the 32 fp32 values of a row (128 bytes) are
//                           replaced IN PLACE by 32 bf16 "hi" values (64 bytes) followed by 32 bf16 "lo" values with
//                           hi + lo = v to ~2^-17 -- the row is then a 64-element bf16 K-major row of the same
//                           128-byte-swizzled tile the TMA wrote, which UMMA reads as is.  No second buffer: a
//                           pipeline stage is 16 KB, which is what lets two CTAs share an SM.
//   warp 1   MMA issuer   : one thread, tcgen05.mma.cta_group::1.kind::f16 (bf16, M=128, N<=128, K=16).  The three
//                           products hi*hi + lo*hi + hi*lo (fp32-grade, ~2^-16) are three MMAs whose A / B descriptors
//                           simply point at the hi or lo half of the row; accumulators live in TMEM, several in
//                           flight.  The residual is added BY THE TENSOR CORE: R tiles ride the same pipeline as extra
//                           k-blocks against a 32x32 identity operand, so the epilogue never reads global memory.
//   warps 6-9 epilogue    : tcgen05.ld (32 lanes x 32 columns) -> shift + activation in registers -> 128B-swizzled
//                           staging tile -> cp.async.bulk.tensor store (ragged edges clipped by the TMA unit).
//                           BatchNorm batch statistics: each lane sums one column of the staged 32x32 tile and keeps
//                           its partial sums in registers across all tiles of the CTA.
// The folded-BatchNorm scale of the epilogue is applied to the WEIGHT rows during their fix-up (W is tiny), which is
// what lets the residual go through the accumulator unscaled.  Weights whose hi/lo tiles fit stay resident in shared
// memory for the whole N tile (all layers of mn10 with M >= 512 000 rows); larger K streams W k-blocks with A.
// HBM-bound at mn10 widths: algorithmic bytes per launch = 4*(M*K + M*N (+ M*N residual) + N*K).
#include <cstdlib>

#include "tma_common.cuh"

// Optional per-role cycle accounting (scripts/timing/: built with -DEAT_TMA_TIMING into a separate library, never part of
// libeat_b200.so): each role accumulates clock64() deltas between marks; one thread per role and CTA dumps
// {4 accumulators, count} to the buffer registered with eat_debug_tma_timing().
#ifdef EAT_TMA_TIMING
__device__ long long* g_tma_timing = nullptr;
extern "C" int eat_debug_tma_timing(long long* buf) {
  return cudaMemcpyToSymbol(g_tma_timing, &buf, sizeof(buf)) == cudaSuccess ? 0 : 2;
}
#define TT_DECL long long t_acc[4] = {0, 0, 0, 0}; long long t_prev = clock64(); int t_cnt = 0;
#define TT_MARK(i) { const long long t_now = clock64(); t_acc[i] += t_now - t_prev; t_prev = t_now; }
#define TT_COUNT ++t_cnt;
#define TT_DUMP(role, cond)                                                            \
  if ((cond) && g_tma_timing != nullptr) {                                             \
    long long* d__ = g_tma_timing + ((size_t)blockIdx.x * 4 + (role)) * 8;             \
    for (int i__ = 0; i__ < 4; ++i__) d__[i__] = t_acc[i__];                           \
    d__[4] = t_cnt;                                                                    \
  }
#else
#define TT_DECL
#define TT_MARK(i)
#define TT_COUNT
#define TT_DUMP(role, cond)
#endif

namespace {
using namespace tc;
using namespace tma;

constexpr int BM = 128;
constexpr int BN_MAX = 256;           // widest N tile (one UMMA instruction covers it: M = 128, N <= 256)
constexpr int A_TILE = BM * 128;        // 16 KB
constexpr int kThreads = 320;           // TMA warp, MMA warp, 4 fix-up warps, 4 epilogue warps
constexpr int kMmaWarp = 1, kFirstFix = 2, kFirstEpi = 6;
constexpr int STG_BYTES = 32 * 128;     // one staged 32 x 32 fp32 sub-tile

struct TmaParams {
  int M, N, K;
  int BN, n_tiles, m_tiles, k_blocks, r_blocks;   // r_blocks: residual k-blocks per tile (0: no residual)
  int bnp;                                        // BN rounded up to 128 / 256: stride of the per-column shared-memory tables
  int stages, wres;                               // wres != 0: weight hi/lo tiles resident per N tile
  int n_acc, acc_cols, tmem_cols;                 // TMEM accumulators in flight (n_acc * acc_cols <= tmem_cols columns)
  int stg_bufs;                                   // staging buffers per epilogue warp (1 or 2)
  int wpre;                                       // weights arrive pre-split (bf16 hi|lo rows, scale folded): no weight fix-up
  int n_samples, tps, tile_rps;                   // M = n_samples * tile_rps rows; tps M tiles per sample (tiles never straddle
                                                  // samples; 1 sample = the whole matrix for ordinary layers)
  int wdyn;                                       // per-sample weights (DynamicConv): the weight map's third coordinate is the sample
  uint32_t stage_bytes, off_w, off_ident, off_stg, off_f, off_bar;   // shared-memory carve-up (bytes)
  int kpad;                                       // floats reserved for each of the in-transform vectors (0: none)
  const float* in_scale; const float* in_shift; const float* gate; int in_act; int rps;
  const float* scale; const float* shift;        // epilogue affine (scale is folded into W)
  double* stat_sum; double* stat_sq;
};

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 B apart (same as pw_tcgen05.cu)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D fp32, A/B bf16 (format 1), both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// EPI : 0 raw output (+ statistics), 1 + shift, 2 + shift + ReLU, 3 + shift + Hardswish   (scale lives in W)
// XACT: -1 raw operand (gate still possible), 0 affine, 1 affine + ReLU, 2 affine + Hardswish on load
template <int EPI, int XACT>
__global__ void __launch_bounds__(kThreads, 2)
pw_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
              const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapR, const TmaParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* s_w = smem + p.off_w;                   // resident weights: [k_blocks][hi | lo][BN rows x 128 B]
  unsigned char* s_ident = smem + p.off_ident;           // 32 x 32 identity, K-major, swizzled (4 KB; residual launches)
  unsigned char* s_stg = smem + p.off_stg;               // [4 epilogue warps][stg_bufs][4 KB]
  float* s_isc = reinterpret_cast<float*>(smem + p.off_f);             // [kpad] in-transform scale (0 beyond K)
  float* s_ish = s_isc + p.kpad;                                       // [kpad]
  float* s_shift = s_ish + p.kpad;                                     // [bnp] epilogue shift of the current N tile
  float* s_stat = s_shift + p.bnp;                                     // [4][2][bnp]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  const int S = p.stages;
  const uint32_t bar_full = smem_u32(bars), bar_ready = bar_full + 8 * S, bar_empty = bar_ready + 8 * S;
  const uint32_t bar_tfull = bar_empty + 8 * S, bar_tempty = bar_tfull + 64, bar_wfull = bar_tempty + 64;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 3 * S + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_ready + 8 * s, 4); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 8; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
    mbar_init(bar_wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapW)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapC)) : "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // one-time tables: identity operand for the residual MMAs, in-transform vectors (zero beyond K: act(0) = 0)
  if (p.r_blocks > 0) {
    for (int i = threadIdx.x; i < 32 * 8; i += kThreads) {             // 32 rows x 8 chunks
      const int r = i >> 3, c = i & 7;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);                              // bf16 1.0 = 0x3F80 at element k = r of row r
      if ((r >> 3) == c) {
        const uint32_t one = (r & 1) ? 0x3F800000u : 0x00003F80u;
        const int w = (r & 7) >> 1;
        if (w == 0) v.x = one; else if (w == 1) v.y = one; else if (w == 2) v.z = one; else v.w = one;
      }
      *reinterpret_cast<uint4*>(s_ident + swz(r, c)) = v;
    }
  }
  if (XACT >= 0) {
    for (int i = threadIdx.x; i < p.kpad; i += kThreads) {
      s_isc[i] = i < p.K ? p.in_scale[i] : 0.f;
      s_ish[i] = i < p.K ? p.in_shift[i] : 0.f;
    }
  }
  for (int i = threadIdx.x; i < 8 * p.bnp; i += kThreads) s_stat[i] = 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int BN = p.BN;
  const int kb_total = p.k_blocks + p.r_blocks;          // pipeline slots per tile
  const uint32_t w_tile = (uint32_t)BN * 128u;           // bytes of one [BN x 32] weight tile
  const uint32_t stage_base = smem_u32(smem);
  // tile walk without divisions (m fastest: a CTA stays on one N tile while it can)
  int nt = blockIdx.x / p.m_tiles, mt = blockIdx.x - nt * p.m_tiles;
  auto next_tile = [&]() { mt += gridDim.x; while (mt >= p.m_tiles) { mt -= p.m_tiles; ++nt; } };
  // (sample, M tile inside the sample) of M tile `m`; one sample = no division on the ordinary path
  auto split_tile = [&](int m, int& bs, int& jt) { if (p.n_samples == 1) { bs = 0; jt = m; } else { bs = m / p.tps; jt = m - bs * p.tps; } };

  if (warp == 0) {
    // ================================================================= TMA producer (one thread)
    if (lane == 0) {
      int s = 0, s_prev = 0, cur_nt = -1;
      uint32_t ph = 0, ph_prev = 0;
      bool first = true;
      TT_DECL
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int bs, jt;
        split_tile(mt, bs, jt);
        const int m0 = jt * BM, n0 = nt * BN;                    // m0: row inside the sample
        const int wb = p.wdyn ? bs : 0;
        const int wkey = nt * p.n_samples + wb;
        if (p.wres && wkey != cur_nt) {
          // every MMA that reads the old weights has retired once the most recently filled stage was released
          if (!first) mbar_wait(bar_empty + 8 * s_prev, ph_prev);
          mbar_expect_tx(bar_wfull, (uint32_t)p.k_blocks * w_tile);
          for (int kb = 0; kb < p.k_blocks; ++kb)
            tma_load_3d(&mapW, bar_wfull, smem_u32(s_w) + (uint32_t)kb * w_tile, kb * (p.wpre ? 2 * KB : KB), n0, wb);
          cur_nt = wkey;
        }
        next_tile();
        for (int kb = 0; kb < kb_total; ++kb) {
          TT_MARK(0)
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          TT_MARK(1)
          const uint32_t dst = stage_base + (uint32_t)s * p.stage_bytes;
          if (kb < p.k_blocks) {
            mbar_expect_tx(bar_full + 8 * s, A_TILE + (p.wres ? 0u : w_tile));
            tma_load_3d(&mapA, bar_full + 8 * s, dst, kb * KB, m0, bs);
            if (!p.wres) tma_load_3d(&mapW, bar_full + 8 * s, dst + A_TILE, kb * (p.wpre ? 2 * KB : KB), n0, wb);
          } else {
            mbar_expect_tx(bar_full + 8 * s, A_TILE);
            tma_load_3d(&mapR, bar_full + 8 * s, dst, n0 + (kb - p.k_blocks) * KB, m0, bs);
          }
          TT_MARK(2)
          s_prev = s; ph_prev = ph; first = false;
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        TT_COUNT
      }
      TT_DUMP(0, true)
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================================================================= MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(BN), idesc_id = idesc_bf16(32);
      const uint64_t desc0 = umma_desc(0);                   // descriptor of address 0: add (address >> 4)
      const uint64_t ident = desc0 + (smem_u32(s_ident) >> 4);
      const uint64_t wres0 = desc0 + (smem_u32(s_w) >> 4);
      const uint32_t w_tile16 = w_tile >> 4;
      int s = 0, acc = 0;
      uint32_t ph = 0, acc_phase = 0;
      TT_DECL
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        TT_MARK(0)
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
        TT_MARK(1)
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.acc_cols);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(bar_ready + 8 * s, ph);
          TT_MARK(2)
          tc_fence_after();
          const uint32_t sa = stage_base + (uint32_t)s * p.stage_bytes;
          // row = [hi: 32 bf16 | lo: 32 bf16]; one K=16 MMA step covers 32 bytes: hi steps at +0/+32 B, lo at +64/+96 B
          const uint64_t a_hi = desc0 + (sa >> 4), a_lo = a_hi + 4;
          if (kb < p.k_blocks) {
            const uint64_t w_hi = p.wres ? wres0 + (uint32_t)kb * w_tile16 : a_hi + (A_TILE >> 4);
            const uint64_t w_lo = w_hi + 4;
            const int krem = p.K - kb * KB;
            const int nk16 = krem >= KB ? 2 : (krem + 15) >> 4;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (j < nk16) {
                const uint64_t ko = (uint64_t)(j * 2);              // 16 bf16 = 32 bytes along K, in 16-byte units
                mma_bf16(tmem_d, a_hi + ko, w_hi + ko, idesc, (kb | j) ? 1u : 0u);
                mma_bf16(tmem_d, a_lo + ko, w_hi + ko, idesc, 1u);
                mma_bf16(tmem_d, a_hi + ko, w_lo + ko, idesc, 1u);
              }
            }
          } else {                 // residual block j: D[:, 32j .. 32j+31] += R_hi . I + R_lo . I
            const uint32_t tmem_r = tmem_d + (uint32_t)(kb - p.k_blocks) * 32u;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t ko = (uint64_t)(j * 2);
              mma_bf16(tmem_r, a_hi + ko, ident + ko, idesc_id, 1u);
              mma_bf16(tmem_r, a_lo + ko, ident + ko, idesc_id, 1u);
            }
          }
          tc_commit(bar_empty + 8 * s);
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        tc_commit(bar_tfull + 8 * acc);
        TT_MARK(3)
        TT_COUNT
        if (++acc == p.n_acc) { acc = 0; acc_phase ^= 1u; }
      }
      TT_DUMP(1, true)
    }
    __syncwarp();
  } else if (warp < kFirstEpi) {
    // ================================================================= fix-up warps (128 threads)
    const int ft = threadIdx.x - kFirstFix * 32;
    int s = 0, cur_nt = -1;
    uint32_t ph = 0, wphase = 0;
    const bool fold = EPI != 0 && p.scale != nullptr;
    auto do_fix_w = [&](unsigned char* w, int kb, int n0) {
      const int lg = pair_lg(p.K - kb * KB);
      for (int h = 0; h < BN; h += 128) {                      // one pass covers 128 weight rows
        unsigned char* wh = w + (size_t)h * 128;
        const int rows = min(128, BN - h);
        if (fold) {
          if (lg == 2) fix_w<2, true>(wh, ft, rows, p.scale, n0 + h, p.N);
          else fix_w<1, true>(wh, ft, rows, p.scale, n0 + h, p.N);
        } else {
          if (lg == 2) fix_w<2, false>(wh, ft, rows, nullptr, n0 + h, p.N);
          else fix_w<1, false>(wh, ft, rows, nullptr, n0 + h, p.N);
        }
      }
    };
    TT_DECL
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int bs, jt;
      split_tile(mt, bs, jt);
      const int n0 = nt * BN;
      const int rows_valid = min(BM, p.tile_rps - jt * BM);
      const int wkey = nt * p.n_samples + (p.wdyn ? bs : 0);
      if (p.wres && wkey != cur_nt) {
        mbar_wait(bar_wfull, wphase);
        wphase ^= 1u;
        if (!p.wpre)
          for (int kb = 0; kb < p.k_blocks; ++kb) do_fix_w(s_w + (size_t)kb * w_tile, kb, n0);
        cur_nt = wkey;
      }
      next_tile();
      int b0 = 0, off0 = 0;
      if (p.gate != nullptr) { const int m0 = bs * p.tile_rps + jt * BM; b0 = m0 / p.rps; off0 = m0 - b0 * p.rps; }
      for (int kb = 0; kb < kb_total; ++kb) {
        TT_MARK(0)
        mbar_wait(bar_full + 8 * s, ph);
        TT_MARK(1)
        unsigned char* tile = smem + (size_t)s * p.stage_bytes;
        if (kb >= p.k_blocks) {
          fix_a<2, -1>(tile, ft, BM, nullptr, nullptr, 0, nullptr, 0, 0, 1, 0);                   // residual tile: plain split
        } else {
          const int lg = pair_lg(p.K - kb * KB);
          const int k = kb * KB + (ft & ((1 << lg) - 1)) * 8;
          if (lg == 2) fix_a<2, XACT>(tile, ft, rows_valid, s_isc, s_ish, k, p.gate, off0, b0, p.rps, p.K);
          else fix_a<1, XACT>(tile, ft, rows_valid, s_isc, s_ish, k, p.gate, off0, b0, p.rps, p.K);
          if (!p.wres && !p.wpre) do_fix_w(tile + A_TILE, kb, n0);
        }
        TT_MARK(2)
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ready + 8 * s);
        TT_MARK(3)
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      TT_COUNT
    }
    TT_DUMP(2, ft == 0)
  } else {
    // ================================================================= epilogue (one warp per TMEM lane quadrant)
    const int q = warp & 3;                                   // TMEM lane quadrant this warp may access
    const int ew = warp - kFirstEpi;
    const int etid = threadIdx.x - kFirstEpi * 32;            // 0..127
    unsigned char* stg = s_stg + (size_t)ew * p.stg_bufs * STG_BYTES;
    constexpr int NC = BN_MAX / 32;                           // 32-column chunks of the widest tile
    const int bnp = p.bnp;
    float* my_stat = s_stat + ew * 2 * bnp;
    const bool do_stats = EPI == 0 && p.stat_sum != nullptr;
    float csum[NC], csq[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { csum[c] = 0.f; csq[c] = 0.f; }
    int cur_nt = -1, acc = 0, cb = 0;
    uint32_t acc_phase = 0;
    TT_DECL
    auto flush_stats = [&](int nt_old) {
      // lane partials -> shared, combine the four warps, one fp64 atomic per channel
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c * 32 < bnp) { my_stat[c * 32 + lane] = csum[c]; my_stat[bnp + c * 32 + lane] = csq[c]; csum[c] = 0.f; csq[c] = 0.f; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int e = etid; e < BN; e += 128) {
        const int n = nt_old * BN + e;
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) { a += s_stat[w * 2 * bnp + e]; b += s_stat[w * 2 * bnp + bnp + e]; }
        if (n < p.N) { atomicAdd(p.stat_sum + n, (double)a); atomicAdd(p.stat_sq + n, (double)b); }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int bs, jt;
      split_tile(mt, bs, jt);
      const int m0 = jt * BM, n0 = nt * BN;                      // m0: row inside the sample
      if (nt != cur_nt) {
        if (do_stats && cur_nt >= 0) flush_stats(cur_nt);
        if (EPI != 0) {
          asm volatile("bar.sync 2, 128;" ::: "memory");       // everyone is done with the previous tile's shifts
          for (int e = etid; e < BN; e += 128) s_shift[e] = (p.shift != nullptr && n0 + e < p.N) ? p.shift[n0 + e] : 0.f;
          asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        cur_nt = nt;
      }
      next_tile();
      TT_MARK(0)
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      TT_MARK(1)
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_cols);
      const int row0 = m0 + q * 32;
      const int rows_left = min(32, p.tile_rps - row0);       // <= 0: nothing of this warp's slab is inside the sample
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c * 32 < BN && n0 + c * 32 < p.N) {
          unsigned char* buf = stg + (size_t)cb * STG_BYTES;
          // the store issued from this buffer (two chunks ago, or the previous one with a single buffer) has drained
          if (p.stg_bufs == 2) { cb ^= 1; if (lane == 0) tma_wait_read<1>(); }
          else if (lane == 0) tma_wait_read<0>();
          __syncwarp();
          uint32_t raw[32];
          tc_ld32(trow + c * 32, raw);
          TT_MARK(2)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                   __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
            if (EPI != 0) {
              const float4 sh = *reinterpret_cast<const float4*>(s_shift + c * 32 + 4 * j);
              v.x = act_out<EPI>(v.x + sh.x); v.y = act_out<EPI>(v.y + sh.y);
              v.z = act_out<EPI>(v.z + sh.z); v.w = act_out<EPI>(v.w + sh.w);
            }
            *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = v;     // row = lane, chunk j, swizzled
          }
          if (do_stats) {
            __syncwarp();
            // lane = column: sum the staged column over the rows of this slab that lie inside M (conflict-free:
            // the 32 lanes of a row read 32 different banks)
            float s1 = 0.f, s2 = 0.f;
            const int cj = lane >> 2, ci = lane & 3;
            if (rows_left >= 32) {
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const float x = *reinterpret_cast<const float*>(buf + r * 128 + ((cj ^ (r & 7)) << 4) + ci * 4);
                s1 += x; s2 = fmaf(x, x, s2);
              }
            } else {
              for (int r = 0; r < rows_left; ++r) {
                const float x = *reinterpret_cast<const float*>(buf + r * 128 + ((cj ^ (r & 7)) << 4) + ci * 4);
                s1 += x; s2 = fmaf(x, x, s2);
              }
            }
            csum[c] += s1; csq[c] += s2;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && rows_left > 0) { tma_store_3d(&mapC, smem_u32(buf), n0 + c * 32, row0, bs); tma_commit(); }
          TT_MARK(3)
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      TT_COUNT
      if (++acc == p.n_acc) { acc = 0; acc_phase ^= 1u; }
    }
    TT_DUMP(3, ew == 0 && lane == 0)
    if (do_stats && cur_nt >= 0) flush_stats(cur_nt);
    if (lane == 0) tma_wait_read<0>();                        // staging buffers must outlive their stores
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols));
  }
}

// ---------------------------------------------------------------------------------------------- host side
// Weights pre-split once per launch: out[n][kb][64 bf16] = 32 hi | 32 lo values of W[n][32 kb ..] (x row_scale[n]), i.e. the
// layout the in-kernel weight fix-up produces, so that the GEMM's TMA loads land finished operand tiles.  W is [N, K], or
// [K, N] when `trans` (the data gradient uses the forward weight transposed; this replaces eat_transpose_f32 there).
__global__ void w_split_kernel(const float* __restrict__ W, const float* __restrict__ row_scale, int trans,
                               uint4* __restrict__ out, int N, int K, int k_blocks) {
  const long long items = (long long)N * k_blocks * 4;             // one item = 8 consecutive k of one row
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int cp = (int)(i & 3);
    const long long t = i >> 2;
    const int kb = (int)(t % k_blocks), n = (int)(t / k_blocks);
    const int k0 = kb * KB + cp * 8;
    float v[8];
    const float sc = row_scale != nullptr ? __ldg(row_scale + n) : 1.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      v[j] = k < K ? __ldg(trans ? W + (size_t)k * N + n : W + (size_t)n * K + k) * sc : 0.f;
    }
    uint4 hi, lo;
    split8(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), hi, lo);
    uint4* row = out + ((size_t)n * k_blocks + kb) * 8;              // 8 chunks of 16 bytes per (row, k-block)
    row[cp] = hi;
    row[4 + cp] = lo;
  }
}

// DynamicConv (reference models/dymn/dy_block.py:103-131): per-sample kernels W_b = sum_j att[b, j] * W_j, mixed in fp32 and
// written pre-split like above: out[b][n][kb][64 bf16].  W holds dyn_k kernels [dyn_k][N][K] ([dyn_k][K][N] when `trans`).
__global__ void w_mix_split_kernel(const float* __restrict__ W, const float* __restrict__ att, int dyn_k,
                                   const float* __restrict__ row_scale, int trans, uint4* __restrict__ out, int B, int N,
                                   int K, int k_blocks) {
  const long long per = (long long)N * k_blocks * 4;
  const long long items = per * B;
  const size_t bank = (size_t)N * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per);
    const long long r = i - (long long)b * per;
    const int cp = (int)(r & 3);
    const long long t = r >> 2;
    const int kb = (int)(t % k_blocks), n = (int)(t / k_blocks);
    const int k0 = kb * KB + cp * 8;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < dyn_k; ++j) a[j] = __ldg(att + (size_t)b * dyn_k + j);
    const float sc = row_scale != nullptr ? __ldg(row_scale + n) : 1.f;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = k0 + e;
      float acc = 0.f;
      if (k < K) {
        const size_t off = trans ? (size_t)k * N + n : (size_t)n * K + k;
        for (int j = 0; j < dyn_k; ++j) acc = fmaf(a[j], __ldg(W + j * bank + off), acc);
      }
      v[e] = acc * sc;
    }
    uint4 hi, lo;
    split8(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), hi, lo);
    uint4* row = out + (((size_t)b * N + n) * k_blocks + kb) * 8;
    row[cp] = hi;
    row[4 + cp] = lo;
  }
}

// [rows, cols] bf16 row-major tensor, box = box_rows x 64 columns (128 bytes), SWIZZLE_128B
int make_map_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) { eat_set_error("pw_tma: cuTensorMapEncodeTiled is not available from this driver"); return EAT_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { eat_set_error("pw_tma: cuTensorMapEncodeTiled (bf16) failed"); return EAT_ERR_CUDA; }
  return EAT_OK;
}

constexpr size_t kSmemLimit = 227 * 1024;

struct WeightWs { int trans; void* ws; size_t bytes; const float* att; int dyn_k; };   // att != nullptr: DynamicConv kernel mix

template <int EPI, int XACT>
int launch_tma(const void* A, const float* W, void* C, const void* R, TmaParams p, cudaStream_t st, WeightWs ws) {
  // ---- tiling.  Shared-memory traffic per k-block of one M tile is n_tiles * (TMA write 16 KB + fix-up 32 KB + 24 KB of A
  // operand reads) + 320 B per padded output column: wide N tiles amortise the A side, which is what bounds the K >= 80
  // layers.  Tiles above 128 columns hold ONE accumulator per CTA (256 TMEM columns when two CTAs share the SM), so they are
  // used only when the main loop is long enough to dwarf the epilogue (k_blocks >= kWideMinKb).
  int bn_max = 208, wide_min_kb = 5;      // 208: widest tile whose streamed weight stage still lets two CTAs share the SM
  if (const char* e = getenv("EAT_TMA_BNMAX")) { const int v = atoi(e); if (v >= 64 && v <= BN_MAX) bn_max = v; }
  if (const char* e = getenv("EAT_TMA_WIDE_MINKB")) wide_min_kb = atoi(e);
  if (ceil_div(p.K, KB) < wide_min_kb && bn_max > 128) bn_max = 128;
  if (p.N <= bn_max) { p.BN = ceil_div(p.N, 16) * 16; p.n_tiles = 1; }
  else {                                   // several N tiles: multiples of 32 so that no store chunk straddles two tiles
    int best = 0; long long best_cost = 0;
    for (int bn = 64; bn <= bn_max; bn += 32) {
      const long long nt = ceil_div(p.N, bn);
      const long long cost = nt * 72 * 1024 + nt * bn * 320;
      if (best == 0 || cost < best_cost || (cost == best_cost && bn > best)) { best = bn; best_cost = cost; }
    }
    p.BN = best; p.n_tiles = ceil_div(p.N, best);
  }
  p.bnp = p.BN <= 128 ? 128 : 256;
  if (p.n_samples < 1) { p.n_samples = 1; p.tile_rps = p.M; }
  p.tps = ceil_div(p.tile_rps, BM);
  p.m_tiles = p.n_samples * p.tps;
  p.k_blocks = ceil_div(p.K, KB);
  p.r_blocks = R != nullptr ? ceil_div(min(p.BN, p.N), 32) : 0;
  p.kpad = XACT >= 0 ? p.k_blocks * KB : 0;
  // ---- shared-memory carve-up: [stages][resident W][identity][staging][floats][barriers].
  // Preferred: TWO co-resident CTAs per SM (<= 113 KB each, 256 TMEM columns each): every per-tile latency chain
  // (single-thread TMA / MMA issue, fix-up, TMEM read-out) is then overlapped by a second, independent tile stream.
  // Needs resident weights and >= 2 stages; otherwise one CTA takes the whole SM.
  const size_t w_tile = (size_t)p.BN * 128;
  const size_t w_res = (size_t)p.k_blocks * w_tile;
  const size_t ident = p.r_blocks > 0 ? 4096 : 0;
  const size_t floats = (2 * (size_t)p.kpad + 9 * (size_t)p.bnp) * 4;
  const size_t barsz = (3 * 8 + 17) * 8 + 16;
  auto fixed = [&](int bufs) { return ident + (size_t)bufs * 4 * STG_BYTES + floats + barsz + 1024 /*alignment slack*/; };
  int ctas = 1;
  bool stream2 = false;
  p.stg_bufs = 2;
  const char* force = getenv("EAT_TMA_CTAS");
  const size_t half = (kSmemLimit - 2048) / 2;
  if (!(force && atoi(force) == 1)) {
    for (int bufs = 2; bufs >= 1 && ctas == 1; --bufs)
      if (fixed(bufs) + w_res + 3 * (size_t)A_TILE <= half) { ctas = 2; p.stg_bufs = bufs; }
    // large K: stream the weight k-blocks with A (stage = A tile + W tile); two CTAs still fit with >= 2 stages each
    for (int bufs = 2; bufs >= 1 && ctas == 1; --bufs)
      if (fixed(bufs) + 2 * ((size_t)A_TILE + w_tile) <= half) { ctas = 2; p.stg_bufs = bufs; stream2 = true; }
  }
  const size_t limit = ctas == 2 ? half : kSmemLimit;
  p.wres = ctas == 2 ? (stream2 ? 0 : 1) : ((fixed(2) + w_res + 4 * (size_t)A_TILE <= kSmemLimit) ? 1 : 0);
  p.stage_bytes = (uint32_t)(A_TILE + (p.wres ? 0 : w_tile));
  const size_t avail = limit - fixed(p.stg_bufs) - (p.wres ? w_res : 0);
  p.stages = (int)(avail / p.stage_bytes);
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) { eat_set_error("pw_tma: shared-memory budget exceeded (K too large for the in-transform tables)"); return EAT_ERR_UNSUPPORTED; }
  // TMEM: n_acc accumulators of acc_cols (>= BN) columns; 256 columns per CTA when two CTAs share the SM
  p.tmem_cols = ctas == 2 ? 256 : 512;
  p.acc_cols = p.BN <= 32 ? 32 : (p.BN <= 64 ? 64 : (p.BN <= 128 ? 128 : 256));
  p.n_acc = p.tmem_cols / p.acc_cols;
  if (p.n_acc > 8) p.n_acc = 8;
  if (const char* e = getenv("EAT_TMA_NACC")) { const int v = atoi(e); if (v >= 1 && v <= p.n_acc) p.n_acc = v; }
  size_t off = (size_t)p.stages * p.stage_bytes;
  p.off_w = (uint32_t)off; off += p.wres ? w_res : 0;
  p.off_ident = (uint32_t)off; off += ident;
  p.off_stg = (uint32_t)off; off += (size_t)p.stg_bufs * 4 * STG_BYTES;
  p.off_f = (uint32_t)off; off += floats;
  off = (off + 7) & ~(size_t)7;
  p.off_bar = (uint32_t)off; off += (3 * (size_t)p.stages + 17) * 8 + 16;
  const size_t smem = off;
  if (smem > limit) { eat_set_error("pw_tma: shared-memory carve-up exceeds its budget"); return EAT_ERR_UNSUPPORTED; }
  // ---- tensor maps
  CUtensorMap mA, mW, mC, mR;
  const long long ns = p.n_samples, rps = p.tile_rps;
  if (int rc = make_map3(&mA, A, ns, rps, p.K, BM, 4)) return rc;
  p.wpre = 0;
  p.wdyn = 0;
  if (ws.ws != nullptr) {
    // pre-split the weights (scale folded, optionally transposed, per sample for DynamicConv) into the caller's workspace
    // and read THAT through TMA
    const long long wsamples = ws.att != nullptr ? ns : 1;
    const size_t need = (size_t)wsamples * p.N * p.k_blocks * 128;
    if (ws.bytes < need || (((uintptr_t)ws.ws) & 127)) { eat_set_error("pw_tma: weight workspace too small or not 128-byte aligned"); return EAT_ERR_ARG; }
    const long long items = wsamples * p.N * p.k_blocks * 4;
    const int grid = (int)min((long long)148 * 8, ceil_div_ll(items, 256));
    const float* fold = (EPI != 0) ? p.scale : nullptr;
    if (ws.att != nullptr) {
      w_mix_split_kernel<<<grid, 256, 0, st>>>(W, ws.att, ws.dyn_k, fold, ws.trans, reinterpret_cast<uint4*>(ws.ws), (int)ns, p.N, p.K, p.k_blocks);
      p.wdyn = 1;
    } else {
      w_split_kernel<<<grid, 256, 0, st>>>(W, fold, ws.trans, reinterpret_cast<uint4*>(ws.ws), p.N, p.K, p.k_blocks);
    }
    EAT_CHECK_LAUNCH();
    p.wpre = 1;
    if (int rc = make_map3(&mW, ws.ws, wsamples, p.N, (long long)p.k_blocks * 64, p.BN, 2)) return rc;
  } else {
    if (ws.trans || ws.att != nullptr) { eat_set_error("pw_tma: transposed / per-sample weights need the weight workspace"); return EAT_ERR_ARG; }
    if (int rc = make_map3(&mW, W, 1, p.N, p.K, p.BN, 4)) return rc;
  }
  if (int rc = make_map3(&mC, C, ns, rps, p.N, 32, 4)) return rc;
  if (int rc = make_map3(&mR, R != nullptr ? R : C, ns, rps, p.N, BM, 4)) return rc;
  static unsigned long long attr_mask = 0;
  if (int rc = eat_opt_in_smem(pw_tma_kernel<EPI, XACT>, kSmemLimit, attr_mask)) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = p.m_tiles * p.n_tiles;
  const int slots = sms * ctas;
  pw_tma_kernel<EPI, XACT><<<tiles < slots ? tiles : slots, kThreads, smem, st>>>(mA, mW, mC, mR, p);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

template <int EPI>
int launch_tma_x(const void* A, const float* W, void* C, const void* R, const TmaParams& p, cudaStream_t st, WeightWs ws) {
  if (p.in_scale == nullptr) return launch_tma<EPI, -1>(A, W, C, R, p, st, ws);
  if (p.in_act == EAT_ACT_RELU) return launch_tma<EPI, 1>(A, W, C, R, p, st, ws);
  if (p.in_act == EAT_ACT_HSWISH) return launch_tma<EPI, 2>(A, W, C, R, p, st, ws);
  return launch_tma<EPI, 0>(A, W, C, R, p, st, ws);
}

}  // namespace

extern "C" int eat_pw_tma_fwd(const float* A, const float* W, int w_trans, float* C, long long M, int N, int K,
                              const float* in_scale, const float* in_shift, int in_act, const float* gate,
                              int rows_per_sample, const float* scale, const float* shift, int act, const float* residual,
                              double* stat_sum, double* stat_sq, void* w_ws, long long w_ws_bytes, cudaStream_t st) {
  if (M == 0) return EAT_OK;
  if (act == EAT_ACT_SIGMOID || in_act == EAT_ACT_SIGMOID) { eat_set_error("pw_tma: sigmoid epilogues run on the CUDA-core GEMM (eat_gemm_simt_fwd)"); return EAT_ERR_UNSUPPORTED; }
  const bool aff = scale != nullptr || shift != nullptr || act != 0;
  if (residual != nullptr && act != EAT_ACT_NONE) { eat_set_error("pw_tma: the residual is accumulated before the activation; residual + activation is not offered"); return EAT_ERR_UNSUPPORTED; }
  if (stat_sum != nullptr && (aff || residual != nullptr)) {
    eat_set_error("pw_tma: batch statistics are produced by the raw-output variant only (no affine/activation/residual)");
    return EAT_ERR_UNSUPPORTED;
  }
  if ((in_scale == nullptr) != (in_shift == nullptr)) { eat_set_error("pw_tma: in_scale and in_shift come together"); return EAT_ERR_ARG; }
  if (K % 4 != 0 || N % 4 != 0) { eat_set_error("pw_tma: K and N must be multiples of 4 (16-byte row pitch for TMA)"); return EAT_ERR_ARG; }
  if (M >= (1ll << 31) - BM) { eat_set_error("pw_tma: M too large"); return EAT_ERR_ARG; }
  if ((((uintptr_t)A) | ((uintptr_t)W) | ((uintptr_t)C) | ((uintptr_t)residual) | ((uintptr_t)gate)) & 15) { eat_set_error("pw_tma: operands must be 16-byte aligned"); return EAT_ERR_ARG; }
  TmaParams p{};
  p.M = (int)M; p.N = N; p.K = K;
  p.in_scale = in_scale; p.in_shift = in_shift; p.gate = gate; p.in_act = in_act; p.rps = rows_per_sample > 0 ? rows_per_sample : 1;
  p.scale = scale; p.shift = shift; p.stat_sum = stat_sum; p.stat_sq = stat_sq;
  const WeightWs ws{w_trans, w_ws, (size_t)(w_ws_bytes > 0 ? w_ws_bytes : 0), nullptr, 0};
  if (!aff) return launch_tma_x<0>(A, W, C, residual, p, st, ws);
  if (act == EAT_ACT_RELU) return launch_tma_x<2>(A, W, C, residual, p, st, ws);
  if (act == EAT_ACT_HSWISH) return launch_tma_x<3>(A, W, C, residual, p, st, ws);
  return launch_tma_x<1>(A, W, C, residual, p, st, ws);
}

// DynamicConv 1x1 (reference models/dymn/dy_block.py:103-131) on the TMA kernel: W holds dyn_k kernels [dyn_k][N][K]
// ([dyn_k][K][N] with w_trans = 1, the data gradient); sample b uses sum_j att[b, j] * W[j], mixed and pre-split once per
// launch into w_ws (B * N * ceil(K/32) * 128 bytes).  M = B * rows_per_sample; tiles, loads and stores never cross a
// sample (3-D tensor maps).
extern "C" int eat_pw_tma_dyn_fwd(const float* A, const float* W, const float* att, int dyn_k, int w_trans, float* C,
                                  long long M, int N, int K, int rows_per_sample, const float* scale, const float* shift,
                                  int act, const float* residual, double* stat_sum, double* stat_sq, void* w_ws,
                                  long long w_ws_bytes, cudaStream_t st) {
  if (M == 0) return EAT_OK;
  if (dyn_k < 1 || dyn_k > 4) { eat_set_error("pw_tma_dyn: 1..4 kernels supported"); return EAT_ERR_UNSUPPORTED; }
  if (rows_per_sample < 1 || M % rows_per_sample != 0) { eat_set_error("pw_tma_dyn: M must be B * rows_per_sample"); return EAT_ERR_ARG; }
  if (act == EAT_ACT_SIGMOID) { eat_set_error("pw_tma_dyn: sigmoid epilogue is not offered"); return EAT_ERR_UNSUPPORTED; }
  const bool aff = scale != nullptr || shift != nullptr || act != 0;
  if (residual != nullptr && act != EAT_ACT_NONE) { eat_set_error("pw_tma_dyn: residual + activation is not offered"); return EAT_ERR_UNSUPPORTED; }
  if (stat_sum != nullptr && (aff || residual != nullptr)) { eat_set_error("pw_tma_dyn: statistics come from the raw-output variant only"); return EAT_ERR_UNSUPPORTED; }
  if (K % 4 != 0 || N % 4 != 0) { eat_set_error("pw_tma_dyn: K and N must be multiples of 4"); return EAT_ERR_ARG; }
  if (M >= (1ll << 31) - BM) { eat_set_error("pw_tma_dyn: M too large"); return EAT_ERR_ARG; }
  if (w_ws == nullptr || att == nullptr) { eat_set_error("pw_tma_dyn: attention weights and the weight workspace are required"); return EAT_ERR_ARG; }
  if ((((uintptr_t)A) | ((uintptr_t)W) | ((uintptr_t)C) | ((uintptr_t)residual)) & 15) { eat_set_error("pw_tma_dyn: operands must be 16-byte aligned"); return EAT_ERR_ARG; }
  TmaParams p{};
  p.M = (int)M; p.N = N; p.K = K;
  p.rps = 1;
  p.scale = scale; p.shift = shift; p.stat_sum = stat_sum; p.stat_sq = stat_sq;
  p.n_samples = (int)(M / rows_per_sample); p.tile_rps = rows_per_sample;
  const WeightWs ws{w_trans, w_ws, (size_t)(w_ws_bytes > 0 ? w_ws_bytes : 0), att, dyn_k};
  if (!aff) return launch_tma_x<0>(A, W, C, residual, p, st, ws);
  if (act == EAT_ACT_RELU) return launch_tma_x<2>(A, W, C, residual, p, st, ws);
  if (act == EAT_ACT_HSWISH) return launch_tma_x<3>(A, W, C, residual, p, st, ws);
  return launch_tma_x<1>(A, W, C, residual, p, st, ws);
}
