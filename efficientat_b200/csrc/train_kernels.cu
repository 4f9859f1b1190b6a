// Training-step kernels outside the network proper (reference ex_audioset.py:135-199):
// spectrogram mixup (:145-146), BCE-with-logits hard-label + knowledge-distillation loss with its
// gradient (:149-189), and a fused Adam/AdamW step over a flat parameter arena (:86-91,198).
#include "common.cuh"

namespace {

// out[b, :] = x[b, :] * lam[b] + x[perm[b], :] * (1 - lam[b])
__global__ void mixup_kernel(const float* __restrict__ x, const int* __restrict__ perm, const float* __restrict__ lam,
                             float* __restrict__ out, int B, long long per) {
  const long long n4 = per / 4;
  const int b = blockIdx.y;
  const float l = lam[b];
  const float4* xa = reinterpret_cast<const float4*>(x + (size_t)b * per);
  const float4* xb = reinterpret_cast<const float4*>(x + (size_t)perm[b] * per);
  float4* o = reinterpret_cast<float4*>(out + (size_t)b * per);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = xa[i], c = xb[i];
    o[i] = make_float4(a.x * l + c.x * (1.f - l), a.y * l + c.y * (1.f - l), a.z * l + c.z * (1.f - l),
                       a.w * l + c.w * (1.f - l));
  }
  if (blockIdx.x == 0) {
    for (long long i = n4 * 4 + threadIdx.x; i < per; i += blockDim.x)
      out[(size_t)b * per + i] = x[(size_t)b * per + i] * l + x[(size_t)perm[b] * per + i] * (1.f - l);
  }
}

__device__ __forceinline__ float bce_logits(float z, float t) {
  return fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)));
}

// loss = kd * mean BCE(z, y_mix) + (1 - kd) * mean_b known[b] * BCE(z, t_mix), with *_mix the mixup blend of row b
// and row perm[b] (BCE is linear in the target, so blending the targets equals blending the two losses as
// ex_audioset.py:170-174 does); known[b] = 0 zeroes the WHOLE distillation row of a clip without teacher predictions
// (ex_audioset.py:166-178: `soft_targets_loss[unknown_indices] *= 0` before the mean over all B rows).
// dlogits = (kd * (sigmoid(z) - y_mix) + (1 - kd) * known[b] * (sigmoid(z) - t_mix)) / (B * C).
// loss_acc[0] += label term (already weighted), loss_acc[1] += distillation term (weighted).
__global__ void bce_kd_kernel(const float* __restrict__ z, const float* __restrict__ y, const float* __restrict__ teacher,
                              const float* __restrict__ known, const int* __restrict__ perm,
                              const float* __restrict__ lam, float kd, int B, int C,
                              float* __restrict__ dz, double* __restrict__ loss_acc) {
  const long long n = (long long)B * C;
  const float inv = 1.f / (float)n;
  float l_hard = 0.f, l_soft = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / C), c = (int)(i % C);
    const float l = lam != nullptr ? lam[b] : 1.f;
    const int pb = perm != nullptr ? perm[b] : b;
    const float zz = z[i];
    const float ym = y[i] * l + y[(size_t)pb * C + c] * (1.f - l);
    const float sg = sigmoidf_(zz);
    float grad = sg - ym;
    l_hard += bce_logits(zz, ym);
    if (teacher != nullptr) {
      const float kn = known != nullptr ? known[b] : 1.f;
      const float tm = teacher[i] * l + teacher[(size_t)pb * C + c] * (1.f - l);
      l_soft += kn * bce_logits(zz, tm);
      grad = kd * (sg - ym) + (1.f - kd) * kn * (sg - tm);
    }
    if (dz != nullptr) dz[i] = grad * inv;
  }
  l_hard = warp_sum(l_hard);
  l_soft = warp_sum(l_soft);
  if ((threadIdx.x & 31) == 0) {
    const float wh = teacher != nullptr ? kd : 1.f;
    atomicAdd(loss_acc, (double)(l_hard * inv * wh));
    if (teacher != nullptr) atomicAdd(loss_acc + 1, (double)(l_soft * inv * (1.f - kd)));
  }
}

// torch.optim.Adam / AdamW semantics on a flat arena; grad_scale folds the 1/world_size of DDP averaging.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                            int adamw, float bc1, float bc2_sqrt, float grad_scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i], gi = g[i] * grad_scale;
    if (wd != 0.f) {
      if (adamw) pi *= 1.f - lr * wd; else gi = fmaf(wd, pi, gi);
    }
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = pi - (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

}  // namespace

extern "C" {

int eat_mixup(const float* x, const int* perm, const float* lam, float* out, int B, long long per_sample,
              cudaStream_t st) {
  if (B == 0 || per_sample == 0) return EAT_OK;
  if (((uintptr_t)x | (uintptr_t)out) & 15 || per_sample % 4 != 0) {
    eat_set_error("mixup: per-sample size must be a multiple of 4 floats and 16-byte aligned"); return EAT_ERR_ARG;
  }
  dim3 grid((unsigned)min((long long)max(1, (148 * 8) / B), ceil_div_ll(per_sample / 4, 256)), B);
  mixup_kernel<<<grid, 256, 0, st>>>(x, perm, lam, out, B, per_sample);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

int eat_bce_kd_loss(const float* logits, const float* y, const float* teacher, const float* teacher_known,
                    const int* perm, const float* lam, float kd_lambda, int B, int C, float* dlogits, double* loss_acc,
                    cudaStream_t st) {
  if (B == 0) return EAT_OK;
  if (kd_lambda < 0.f || kd_lambda > 1.f) { eat_set_error("bce_kd_loss: kd_lambda must be in [0, 1] (ex_audioset.py:100)"); return EAT_ERR_ARG; }
  int grid = (int)min((long long)148 * 2, ceil_div_ll((long long)B * C, 256));
  bce_kd_kernel<<<grid, 256, 0, st>>>(logits, y, teacher, teacher_known, perm, lam, kd_lambda, B, C, dlogits, loss_acc);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

int eat_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int adamw, int step, float grad_scale, cudaStream_t st) {
  if (n == 0) return EAT_OK;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  int grid = (int)min((long long)148 * 8, ceil_div_ll(n, 256));
  adam_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, adamw, bc1, bc2s, grad_scale);
  EAT_CHECK_LAUNCH();
  return EAT_OK;
}

}  // extern "C"
