// Small kernels of the DyMN context path (reference models/dymn/dy_block.py):
//   ContextGen pooling (:236-237), sequence average pooling (:227-233,249), DynamicConv attention
//   softmax(Linear(h_c)/T) (:104-107) and the per-sample depthwise weight mix (:111-117).
#include "common.cuh"

namespace {

// out[b, pos, c]: pos < F -> mean over t of x[b, pos, :, c];  pos >= F -> mean over f of x[b, :, pos-F, c]
template <typename T>
__global__ void __launch_bounds__(256) ctx_pool_kernel(const T* __restrict__ x, float* __restrict__ out, int F, int Tn,
                                                       int C) {
  constexpr int V = Vec<T>::N;
  const int cv = C / V;
  const int b = blockIdx.y;
  const int items = (F + Tn) * cv;
  const T* xb = x + (size_t)b * F * Tn * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += gridDim.x * blockDim.x) {
    const int cvi = i % cv, pos = i / cv;
    const int c0 = cvi * V;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    if (pos < F) {
      const T* p = xb + (size_t)pos * Tn * C + c0;
      for (int t = 0; t < Tn; ++t) {
        float v[V];
        Vec<T>::load(p + (size_t)t * C, v);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += v[k];
      }
      const float inv = 1.f / Tn;
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] *= inv;
    } else {
      const T* p = xb + (size_t)(pos - F) * C + c0;
      for (int f = 0; f < F; ++f) {
        float v[V];
        Vec<T>::load(p + (size_t)f * Tn * C, v);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += v[k];
      }
      const float inv = 1.f / F;
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] *= inv;
    }
    float* o = out + ((size_t)b * (F + Tn) + pos) * C + c0;
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = acc[k];
  }
}

// AvgPool (kernel 3, stride s, pad 1, count_include_pad) along a sequence, or a plain copy when stride == 1.
// in: rows [row0, row0 + L) of each sample of a [B, Ltot, H] tensor -> out [B, Lo, H]
__global__ void seq_pool_kernel(const float* __restrict__ in, float* __restrict__ out, int Ltot, int row0, int L, int Lo,
                                int H, int stride) {
  const int b = blockIdx.y;
  const int items = Lo * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += gridDim.x * blockDim.x) {
    const int h = i % H, lo = i / H;
    const float* base = in + ((size_t)b * Ltot + row0) * H + h;
    float v;
    if (stride == 1) v = base[(size_t)lo * H];
    else {
      float acc = 0.f;
#pragma unroll
      for (int d = -1; d <= 1; ++d) {
        const int l = lo * stride + d;
        if (l >= 0 && l < L) acc += base[(size_t)l * H];
      }
      v = acc * (1.f / 3.f);
    }
    out[((size_t)b * Lo + lo) * H + h] = v;
  }
}

// att[b, :] = softmax((Wr h_c[b] + br) / temperature);  one warp per sample, k <= 4
__global__ void dyconv_att_kernel(const float* __restrict__ hc, const float* __restrict__ wr, const float* __restrict__ br,
                                  float inv_temp, float* __restrict__ att, int B, int H, int k) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  float logit[4];
  for (int j = 0; j < k; ++j) {
    float acc = 0.f;
    for (int h = lane; h < H; h += 32) acc = fmaf(__ldg(wr + (size_t)j * H + h), hc[(size_t)b * H + h], acc);
    logit[j] = (warp_sum(acc) + br[j]) * inv_temp;
  }
  if (lane == 0) {
    float mx = logit[0];
    for (int j = 1; j < k; ++j) mx = fmaxf(mx, logit[j]);
    float e[4], s = 0.f;
    for (int j = 0; j < k; ++j) { e[j] = expf(logit[j] - mx); s += e[j]; }
    for (int j = 0; j < k; ++j) att[(size_t)b * k + j] = e[j] / s;
  }
}

// wt[b, tap, c] = sum_j att[b, j] * W[j, c, tap]     (W: [k][C][kk] as stored by DynamicConv.weight)
__global__ void dyconv_mix_dw_kernel(const float* __restrict__ w, const float* __restrict__ att, float* __restrict__ wt,
                                     int C, int kk, int k) {
  const int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * kk; i += gridDim.x * blockDim.x) {
    const int c = i / kk, tap = i % kk;
    float acc = 0.f;
    for (int j = 0; j < k; ++j) acc = fmaf(att[(size_t)b * k + j], __ldg(w + (size_t)j * C * kk + i), acc);
    wt[((size_t)b * kk + tap) * C + c] = acc;
  }
}

}  // namespace

extern "C" {

int eat_ctx_pool(const void* x, int dtype, float* out, int B, int F, int T, int C, cudaStream_t st) {
  if (B == 0) return EAT_OK;
  const int V = dtype == EAT_BF16 ? 8 : 4;
  if (C % V != 0) { eat_set_error("ctx_pool: channels must be a multiple of the vector width"); return EAT_ERR_ARG; }
  dim3 grid(ceil_div((F + T) * (C / V), 256), B);
  if (dtype == EAT_BF16) ctx_pool_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, out, F, T, C);
  else ctx_pool_kernel<float><<<grid, 256, 0, st>>>((const float*)x, out, F, T, C);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

int eat_seq_pool(const float* in, float* out, int B, int Ltot, int row0, int L, int H, int stride, cudaStream_t st) {
  if (B == 0) return EAT_OK;
  const int Lo = stride == 1 ? L : (L + 2 - 3) / stride + 1;
  dim3 grid(ceil_div(Lo * H, 256), B);
  seq_pool_kernel<<<grid, 256, 0, st>>>(in, out, Ltot, row0, L, Lo, H, stride);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

int eat_dyconv_att(const float* hc, const float* wr, const float* br, float temperature, float* att, int B, int H, int k,
                   cudaStream_t st) {
  if (B == 0) return EAT_OK;
  if (k < 1 || k > 4) { eat_set_error("dyconv_att: 1..4 kernels supported"); return EAT_ERR_UNSUPPORTED; }
  dyconv_att_kernel<<<ceil_div(B, 4), 128, 0, st>>>(hc, wr, br, 1.f / temperature, att, B, H, k);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

int eat_dyconv_mix_dw(const float* w, const float* att, float* wt, int B, int C, int ksize, int k, cudaStream_t st) {
  if (B == 0) return EAT_OK;
  dim3 grid(ceil_div(C * ksize * ksize, 256), B);
  dyconv_mix_dw_kernel<<<grid, 256, 0, st>>>(w, att, wt, C, ksize * ksize, k);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

}  // extern "C"
