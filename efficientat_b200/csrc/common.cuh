// Shared device helpers for the EfficientAT B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/eat_b200.h"   // the C ABI: declarations here are checked against definitions

#define EAT_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) { eat_set_error(cudaGetErrorString(e__)); return EAT_ERR_CUDA; } \
  } while (0)

void eat_set_error(const char* msg);

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == EAT_ACT_RELU) return fmaxf(v, 0.f);
  if (act == EAT_ACT_HSWISH) return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
  if (act == EAT_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}
// derivative of the activation w.r.t. its input, evaluated at pre-activation v
// (torch: hardswish' = 0 for v<-3, 1 for v>3, (2v+3)/6 otherwise; relu' = v>0)
__device__ __forceinline__ float act_bwd(float v, int act) {
  if (act == EAT_ACT_RELU) return v > 0.f ? 1.f : 0.f;
  if (act == EAT_ACT_HSWISH) return v < -3.f ? 0.f : (v <= 3.f ? (2.f * v + 3.f) * (1.f / 6.f) : 1.f);
  return 1.f;
}
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// ---- vector load/store of V contiguous channels as fp32, for fp32 (V=4) and bf16 (V=8) storage.
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ __forceinline__ static void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ __forceinline__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ __forceinline__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Per-channel affine + activation + per-(sample,channel) gate applied to an operand as it is loaded:
//   v' = act(v * scale[c] + shift[c]) * gate[b * C + c]
// scale == nullptr -> no affine/act;  gate == nullptr -> no gate.
struct InXform {
  const float* scale;
  const float* shift;
  const float* gate;
  int act;
  int rows_per_sample;   // F*T of the tensor the gate indexes (rows -> sample index)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
