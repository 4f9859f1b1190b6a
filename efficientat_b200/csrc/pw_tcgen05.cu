// Pointwise (1x1) convolution / Linear as a tcgen05 GEMM for sm_100a:
//     C[M,N] = epi( xf(A)[M,K] . W[N,K]^T )        A, C: NHWC activation rows (fp32 or bf16), W: fp32
// Replaces the 1x1 ConvNormActivation layers of the reference (models/mn/block_types.py:140-147,
// 167-171; models/mn/model.py:160-166) -- 90 % of mn10's MACs.
//
// Design (one persistent CTA per SM, 9 warps, warp-specialised, mbarrier pipelines):
//   warps 0-3  producers : global -> registers -> (BatchNorm affine + activation + SE gate of the
//                          producing layer, fused on load) -> bf16 -> 128B-swizzled K-major smem tiles.
//                          fp32 activations/weights are split x = hi + lo (two bf16) so that three
//                          MMAs (hi*hi + lo*hi + hi*lo) reproduce fp32 products to ~2^-16.
//   warp  4    MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256,
//                          K=16) on UMMA smem descriptors; accumulators live in TMEM (2 x 256 columns,
//                          double buffered so the epilogue of tile i overlaps the MMAs of tile i+1).
//   warps 5-8  epilogue  : tcgen05.ld (32 lanes x 32 columns) -> BN affine / activation / residual ->
//                          per-warp smem transpose -> 128-byte coalesced stores; per-channel batch
//                          statistics (sum, sum of squares) accumulated in registers across tiles.
// The kernel is HBM-bound for mn10 widths (arithmetic intensity below the ridge); algorithmic bytes per
// launch = M*K*sizeof(A) + M*N*sizeof(C) (+ residual) + N*K*4.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // bf16 elements per k-block (= one 128-byte swizzle row)
constexpr int kProducerThreads = 128;
constexpr int kThreads = 288;          // 4 producer warps + 1 MMA warp + 4 epilogue warps
constexpr int kMmaWarp = 4;
constexpr int kFirstEpiWarp = 5;
constexpr int A_TILE_BYTES = BM * 128; // 16 KB

struct TcParams {
  const void* A;
  const float* W;
  void* C;
  const void* residual;
  int M, N, K;
  InXform xf;
  const float* scale;
  const float* shift;
  int act;
  double* stat_sum;
  double* stat_sq;
  int BN;        // columns per N tile (multiple of 16)
  int n_tiles, m_tiles, k_blocks;
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout=2 [61,64))
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major), 16 B
  d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                     // SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ uint32_t swz(int row, int chunk) {   // byte offset of a 16-byte chunk in a [rows][128 B] tile
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int NP>
__device__ __forceinline__ void store_chunk(unsigned char* hi_tile, unsigned char* lo_tile, uint32_t off, const float (&v)[8]) {
  uint4 h;
  h.x = pack_bf16(v[0], v[1]); h.y = pack_bf16(v[2], v[3]); h.z = pack_bf16(v[4], v[5]); h.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(hi_tile + off) = h;
  if (NP == 2) {
    float r[8];
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(hh[i]);
      r[2 * i] = v[2 * i] - f.x;
      r[2 * i + 1] = v[2 * i + 1] - f.y;
    }
    uint4 l;
    l.x = pack_bf16(r[0], r[1]); l.y = pack_bf16(r[2], r[3]); l.z = pack_bf16(r[4], r[5]); l.w = pack_bf16(r[6], r[7]);
    *reinterpret_cast<uint4*>(lo_tile + off) = l;
  }
}

template <typename T>
__device__ __forceinline__ void load_chunk(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load_chunk<float>(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load_chunk<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// ------------------------------------------------------------------------------------------ kernel
// T: activation storage type.  NP = 1: operands rounded to bf16 (bf16 mode); NP = 2: hi/lo split (fp32 mode).
template <typename T, int NP, int STAGES, int BN_MAX>
__global__ void __launch_bounds__(kThreads, 1) pw_tc_kernel(TcParams p) {
  constexpr int B_TILE_BYTES = BN_MAX * 128;
  constexpr int STAGE_BYTES = NP * (A_TILE_BYTES + B_TILE_BYTES);
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* stage_base = smem;
  float* s_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);      // 4 warps x 32 x 33 floats
  float* s_scale = s_stage + 4 * 32 * 33;                                      // [BN_MAX]
  float* s_shift = s_scale + BN_MAX;                                           // [BN_MAX]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + BN_MAX);              // full[S], empty[S], tfull[2], tempty[2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 2 * STAGES), bar_tempty = smem_u32(bars + 2 * STAGES + 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, kProducerThreads); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int K = p.K, N = p.N, BN = p.BN;

  if (warp < 4) {
    // ================================================================= producers
    const int ptid = threadIdx.x;
    const int kc = ptid & 7, r0 = ptid >> 3;        // this thread's 16-byte chunk column and first row
    const T* __restrict__ A = reinterpret_cast<const T*>(p.A);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;       // m fastest: a CTA stays on one N tile
      const long long m0 = (long long)mt * BM;
      const int n0 = nt * BN;
      const float* gate_row[8];
      if (p.xf.gate != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          long long m = m0 + r0 + 16 * i;
          gate_row[i] = m < p.M ? p.xf.gate + (m / p.xf.rows_per_sample) * K : nullptr;
        }
      }
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        unsigned char* sA_hi = stage_base + stage * STAGE_BYTES;
        unsigned char* sA_lo = sA_hi + A_TILE_BYTES;                       // only used when NP == 2
        unsigned char* sB_hi = sA_hi + NP * A_TILE_BYTES;
        unsigned char* sB_lo = sB_hi + B_TILE_BYTES;
        const int k = kb * BK + kc * 8;
        const bool kok = k < K;
        float isc[8], ish[8];
        if (p.xf.scale != nullptr && kok) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { isc[j] = __ldg(p.xf.scale + k + j); ish[j] = __ldg(p.xf.shift + k + j); }
        }
        // ---- A: 128 rows x 8 chunks, 8 rows per thread; loads first, then transform + store
        float av[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long m = m0 + r0 + 16 * i;
          if (kok && m < p.M) load_chunk<T>(A + m * K + k, av[i]);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) av[i][j] = 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long m = m0 + r0 + 16 * i;
          if (kok && m < p.M) {
            if (p.xf.scale != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) av[i][j] = act_fwd(fmaf(av[i][j], isc[j], ish[j]), p.xf.act);
            }
            if (p.xf.gate != nullptr) {
              const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate_row[i] + k));
              const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate_row[i] + k) + 1);
              av[i][0] *= g0.x; av[i][1] *= g0.y; av[i][2] *= g0.z; av[i][3] *= g0.w;
              av[i][4] *= g1.x; av[i][5] *= g1.y; av[i][6] *= g1.z; av[i][7] *= g1.w;
            }
          }
          store_chunk<NP>(sA_hi, sA_lo, swz(r0 + 16 * i, kc), av[i]);
        }
        // ---- B: BN rows (output channels) x 8 chunks of the fp32 weight matrix
        for (int r = r0; r < BN; r += 16) {
          const int n = n0 + r;
          float wv[8];
          if (kok && n < N) load_chunk<float>(p.W + (size_t)n * K + k, wv);
          else {
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[j] = 0.f;
          }
          store_chunk<NP>(sB_hi, sB_lo, swz(r, kc), wv);
        }
        fence_proxy_async();
        mbar_arrive(bar_full + 8 * stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================================================= MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN_MAX;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sA_hi = smem_u32(stage_base + stage * STAGE_BYTES);
          const uint32_t sA_lo = sA_hi + A_TILE_BYTES;
          const uint32_t sB_hi = sA_hi + NP * A_TILE_BYTES;
          const uint32_t sB_lo = sB_hi + B_TILE_BYTES;
          const int krem = K - kb * BK;
          const int nk16 = krem >= BK ? BK / 16 : (krem + 15) / 16;
          for (int k16 = 0; k16 < nk16; ++k16) {
            const uint32_t ko = k16 * 32;                         // 16 bf16 = 32 bytes along K inside the swizzle row
            const uint32_t first = (kb == 0 && k16 == 0) ? 0u : 1u;
            tc_mma(tmem_d, umma_desc(sA_hi + ko), umma_desc(sB_hi + ko), idesc, first);
            if (NP == 2) {
              tc_mma(tmem_d, umma_desc(sA_lo + ko), umma_desc(sB_hi + ko), idesc, 1u);
              tc_mma(tmem_d, umma_desc(sA_hi + ko), umma_desc(sB_lo + ko), idesc, 1u);
            }
          }
          tc_commit(bar_empty + 8 * stage);                       // frees the smem stage when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_tfull + 8 * acc);                           // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ================================================================= epilogue
    const int q = warp & 3;                                       // TMEM lane quadrant this warp may access
    float* stg = s_stage + q * 32 * 33;
    T* __restrict__ C = reinterpret_cast<T*>(p.C);
    const T* __restrict__ R = reinterpret_cast<const T*>(p.residual);
    const int etid = threadIdx.x - kFirstEpiWarp * 32;            // 0..127
    constexpr int NCH = BN_MAX / 32;
    float ssum[NCH], ssq[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { ssum[c] = 0.f; ssq[c] = 0.f; }
    int cur_nt = -1;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool do_stats = p.stat_sum != nullptr;
    auto flush_stats = [&](int nt) {
      if (!do_stats || nt < 0) return;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int n = nt * BN + c * 32 + lane;
        if (c * 32 < BN && n < N && n < (nt + 1) * BN) {
          atomicAdd(p.stat_sum + n, (double)ssum[c]);
          atomicAdd(p.stat_sq + n, (double)ssq[c]);
        }
        ssum[c] = 0.f; ssq[c] = 0.f;
      }
    };
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;
      const long long m0 = (long long)mt * BM;
      const int n0 = nt * BN;
      if (nt != cur_nt) {
        flush_stats(cur_nt);
        cur_nt = nt;
        // stage this N tile's epilogue affine in smem (named barrier over the 128 epilogue threads)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = etid; i < BN; i += 128) {
          const int n = n0 + i;
          s_scale[i] = (p.scale != nullptr && n < N) ? p.scale[n] : 1.f;
          s_shift[i] = (p.shift != nullptr && n < N) ? p.shift[n] : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN_MAX;
      const long long mrow0 = m0 + q * 32;
#pragma unroll 1
      for (int c = 0; c * 32 < BN; ++c) {
        uint32_t raw[32];
        tc_ld32(trow + c * 32, raw);
        // lane == row (mrow0 + lane), raw[j] == column n0 + c*32 + j
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(raw[j]);
        __syncwarp();
        const int n = n0 + c * 32 + lane;                          // now lane == column
        const bool nok = (c * 32 + lane < BN) && n < N;
        const float sc = s_scale[c * 32 + lane], sh = s_shift[c * 32 + lane];
        float cs = 0.f, cq = 0.f;
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const long long m = mrow0 + r;
          if (m >= p.M) break;
          float v = stg[r * 33 + lane];
          cs += v;
          cq = fmaf(v, v, cq);
          if (nok) {
            v = act_fwd(fmaf(v, sc, sh), p.act);
            if (R != nullptr) v += to_f32<T>(R[m * N + n]);
            C[m * N + n] = from_f32<T>(v);
          }
        }
        if (do_stats) {
#pragma unroll
          for (int cc = 0; cc < NCH; ++cc)
            if (cc == c) { ssum[cc] += cs; ssq[cc] += cq; }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    flush_stats(cur_nt);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

template <typename T, int NP, int STAGES, int BN_MAX>
int launch_tc(const TcParams& p0, cudaStream_t st) {
  TcParams p = p0;
  p.n_tiles = ceil_div(p.N, BN_MAX);
  p.BN = ceil_div(ceil_div(p.N, p.n_tiles), 16) * 16;
  p.n_tiles = ceil_div(p.N, p.BN);
  p.m_tiles = ceil_div(p.M, BM);
  p.k_blocks = ceil_div(p.K, BK);
  constexpr size_t smem = (size_t)STAGES * NP * (A_TILE_BYTES + BN_MAX * 128) + 4 * 32 * 33 * sizeof(float) +
                          2 * BN_MAX * sizeof(float) + (2 * STAGES + 4) * sizeof(uint64_t) + 16;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(pw_tc_kernel<T, NP, STAGES, BN_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { eat_set_error(cudaGetErrorString(e)); return EAT_ERR_CUDA; }
    attr_done = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < sms ? tiles : sms;
  pw_tc_kernel<T, NP, STAGES, BN_MAX><<<grid, kThreads, smem, st>>>(p);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

}  // namespace

extern "C" int eat_pw_tc_fwd(const void* A, int a_dtype, const float* W, int w_trans, void* C, int c_dtype, long long M,
                             int N, int K, const float* in_scale, const float* in_shift, int in_act, const float* gate,
                             int rows_per_sample, const float* scale, const float* shift, int act, const void* residual,
                             double* stat_sum, double* stat_sq, cudaStream_t st) {
  if (M == 0) return EAT_OK;
  if (w_trans) { eat_set_error("pw_tc: transposed weights are not supported (pre-transpose with eat_transpose_f32)"); return EAT_ERR_UNSUPPORTED; }
  if (a_dtype != c_dtype) { eat_set_error("pw_tc: A and C must share the storage dtype"); return EAT_ERR_UNSUPPORTED; }
  if (K % 8 != 0 || N % 8 != 0) { eat_set_error("pw_tc: K and N must be multiples of 8"); return EAT_ERR_ARG; }
  if (M >= (1ll << 31) - BM) { eat_set_error("pw_tc: M too large"); return EAT_ERR_ARG; }
  if ((((uintptr_t)A) | ((uintptr_t)W) | ((uintptr_t)C)) & 15) { eat_set_error("pw_tc: operands must be 16-byte aligned"); return EAT_ERR_ARG; }
  TcParams p;
  p.A = A; p.W = W; p.C = C; p.residual = residual; p.M = (int)M; p.N = N; p.K = K;
  p.xf = InXform{in_scale, in_shift, gate, in_act, rows_per_sample > 0 ? rows_per_sample : 1};
  p.scale = scale; p.shift = shift; p.act = act; p.stat_sum = stat_sum; p.stat_sq = stat_sq;
  if (a_dtype == EAT_BF16) return launch_tc<__nv_bfloat16, 1, 4, 256>(p, st);
  return launch_tc<float, 2, 3, 128>(p, st);
}

// [R, Cc] fp32 -> [Cc, R]  (weights for the data-gradient GEMM)
namespace {
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
  __shared__ float tile[32][33];
  int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
  for (int j = 0; j < 32; j += 8)
    if (x < Cc && y + j < R) tile[threadIdx.y + j][threadIdx.x] = in[(size_t)(y + j) * Cc + x];
  __syncthreads();
  x = blockIdx.y * 32 + threadIdx.x;
  y = blockIdx.x * 32 + threadIdx.y;
  for (int j = 0; j < 32; j += 8)
    if (x < R && y + j < Cc) out[(size_t)(y + j) * R + x] = tile[threadIdx.x][threadIdx.y + j];
}
}  // namespace

extern "C" int eat_transpose_f32(const float* in, float* out, int rows, int cols, cudaStream_t st) {
  if (rows == 0 || cols == 0) return EAT_OK;
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(in, out, rows, cols);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}
