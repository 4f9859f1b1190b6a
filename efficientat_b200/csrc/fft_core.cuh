// Radix-8 Stockham passes for a 512-point complex FFT (= 1024-point real FFT after the
// split step).  Written __host__ __device__ so the exact index/twiddle logic is unit-tested
// on the CPU (tests/test_fft_core.py compiles this header with g++).
#pragma once
#if defined(__CUDACC__)
#define EAT_HD __host__ __device__ __forceinline__
#else
#define EAT_HD inline
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

EAT_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
EAT_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
EAT_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
EAT_HD float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

EAT_HD void bfly2(float2& a, float2& b) { float2 t = a; a = cadd(t, b); b = csub(t, b); }

// in-register 8-point DFT (forward, e^{-i...}); result in natural order in v[0..7]
EAT_HD void fft8(float2 (&v)[8]) {
  const float h = 0.70710678118654752440f;
  bfly2(v[0], v[4]); bfly2(v[1], v[5]); bfly2(v[2], v[6]); bfly2(v[3], v[7]);
  v[5] = cmul(v[5], make_float2(h, -h));
  v[6] = cmul_mi(v[6]);
  v[7] = cmul(v[7], make_float2(-h, -h));
  // two 4-point DFTs on (0,1,2,3) and (4,5,6,7)
  bfly2(v[0], v[2]); bfly2(v[1], v[3]); v[3] = cmul_mi(v[3]); bfly2(v[0], v[1]); bfly2(v[2], v[3]);
  bfly2(v[4], v[6]); bfly2(v[5], v[7]); v[7] = cmul_mi(v[7]); bfly2(v[4], v[5]); bfly2(v[6], v[7]);
  // outputs are bit-reversed: X0=v0 X1=v4 X2=v2 X3=v6 X4=v1 X5=v5 X6=v3 X7=v7
  float2 t1 = v[1], t3 = v[3], t4 = v[4], t6 = v[6];
  v[1] = t4; v[3] = t6; v[4] = t1; v[6] = t3;
}

// One Stockham pass, work item j in [0,64): twiddle + butterfly on 8 values already in v
// (v[r] = in[j + 64 r]), Ns in {1, 8, 64}; tw[m] = exp(-2 pi i m / 512).
EAT_HD void stockham8_compute(float2 (&v)[8], int j, int Ns, const float2* tw) {
  if (Ns > 1) {
    int k = j & (Ns - 1);
    int step = 512 / (Ns * 8);
#pragma unroll
    for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], tw[r * k * step]);
  }
  fft8(v);
}
// destination index of output r of work item j
EAT_HD int stockham8_dst(int j, int Ns, int r) {
  int k = j & (Ns - 1);
  return (j - k) * 8 + k + r * Ns;
}

// Real-FFT split: given Z = FFT512(x[2n] + i x[2n+1]), returns X[k] of the 1024-point real
// FFT for k in [0,512];  w1024 = exp(-2 pi i k / 1024).
EAT_HD float2 rfft_split(float2 zk, float2 zmk, float2 w1024) {
  // E = (Zk + conj(Zmk))/2 ; O = -i (Zk - conj(Zmk))/2 ; X = E + w O
  float2 e = make_float2(0.5f * (zk.x + zmk.x), 0.5f * (zk.y - zmk.y));
  float2 d = make_float2(0.5f * (zk.x - zmk.x), 0.5f * (zk.y + zmk.y));
  float2 o = make_float2(d.y, -d.x);
  return cadd(e, cmul(w1024, o));
}
