/* libeat_b200 -- C ABI of the B200-native EfficientAT hot path (mel front end + MobileNetV3 /
 * DyMN forward & backward).  sm_100a only; no CPU fallback.
 *
 * The reference (fschmid56/EfficientAT) has no FFI: its hot path is Python nn.Modules calling
 * cuFFT / cuDNN / cuBLAS through PyTorch.  Each entry point below therefore cites the reference
 * *module code* (file:line under the reference tree) whose library calls it replaces.  The host
 * side (efficientat_b200/models/...) mirrors the reference's nn.Module surface and reaches these
 * functions through ctypes with raw device pointers + the current CUDA stream.
 *
 * Conventions
 *   - every function returns 0 (EAT_OK) or an EAT_ERR_* code; eat_last_error() has the text.
 *   - all pointers are DEVICE pointers unless stated; the caller owns all memory.
 *   - work is enqueued on `stream` and is asynchronous; no function synchronises or allocates.
 *   - activations are NHWC: [B, F, T, C] with C innermost, dtype EAT_F32 or EAT_BF16;
 *     parameters, BatchNorm vectors, statistics and gates are always fp32 (statistics: fp64).
 *   - "in_scale/in_shift/in_act": optional per-channel affine + activation applied to the input
 *     operand as it is loaded (the BatchNorm+activation of the producing layer);  NULL = none.
 *   - "scale/shift/act": optional per-channel affine + activation applied to the result.
 *   - "stat_sum/stat_sq": optional fp64 per-channel sum / sum of squares of the RAW result
 *     (BatchNorm batch statistics), accumulated with atomics; caller zeroes them.
 */
#ifndef EAT_B200_H
#define EAT_B200_H

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define EAT_OK 0
#define EAT_ERR_ARG 1
#define EAT_ERR_CUDA 2
#define EAT_ERR_UNSUPPORTED 3

#define EAT_ACT_NONE 0
#define EAT_ACT_RELU 1
#define EAT_ACT_HSWISH 2
#define EAT_ACT_SIGMOID 3   /* forward-only (DyMN coordinate attention / DyReLU coefficient nets) */

#define EAT_F32 0
#define EAT_BF16 1

const char* eat_last_error(void);
int eat_abi_version(void);

/* Host-only (no GPU work): the launch plan of the sliding-window depthwise kernels (csrc/dw_slide.cu) for a layer.
 * kind: 0 forward, 1 weight gradient, 2 stride-2 data gradient.  per_sample != 0: blockIdx.y must be the sample
 * (squeeze-excitation pooling, DynamicConv per-sample weights).  plan[6] = {channel chunks, channel vectors per chunk,
 * output rows (row pairs for kind 2) per segment, CTA groups per chunk, gridDim.y, strip width}.  Exposed so the
 * host logic is testable without a device (tests/test_cabi.py). */
int eat_dw_plan(int kind, int dtype, int B, int F, int T, int C, int k, int stride, int per_sample, int* plan);
int eat_device_check(int device);

/* Fused log-mel front end.  Replaces AugmentMelSTFT.forward, models/preprocess.py:40-67
 * (conv1d pre-emphasis :41, torch.stft :42-43, power :44, mel matmul :56-57, log :59, affine :65).
 * wave [B,N] fp32 -> out [B, n_mels, 1+(N-1)/hop] fp32.  twiddle: 1024 float2 (exp(-2 pi i m/512),
 * m<512, then exp(-2 pi i k/1024), k<512).  Filterbank as bands: fb_start/fb_len [n_mels],
 * fb_w [max_len][n_mels] (tap-major). */
int eat_mel_fwd(const float* wave, int B, int N, const float* window, int win_length, int hop, int n_fft,
                const float* twiddle, const int* fb_start, const int* fb_len, const float* fb_w, int max_len,
                int n_mels, float preemph, float* out, cudaStream_t stream);

/* Kaldi mel filterbank in banded form, built on the device (training-mode fmin/fmax jitter,
 * models/preprocess.py:45-55 + torchaudio.compliance.kaldi.get_mel_banks).  fb_w holds cap x n_mels floats. */
int eat_mel_filterbank(int n_mels, int n_fft, float sample_rate, double fmin, double fmax, int* fb_start,
                       int* fb_len, float* fb_w, int cap, cudaStream_t stream);

/* SpecAugment masking of the log-mel (training only): torchaudio Frequency/TimeMasking with
 * iid_masks=True, models/preprocess.py:31-38,61-63.  spec [B,F,T]; band [start,end) per example. */
int eat_mel_mask(float* spec, int B, int F, int T, const int* f_start, const int* f_end, const int* t_start,
                 const int* t_end, float fill, cudaStream_t stream);

/* Stem 3x3 conv on the 1-channel spectrogram.  Replaces ConvNormActivation(1, C, k=3, s=2) at
 * models/mn/model.py:125-133 (dymn/model.py:80-87).  x [B,F,T] fp32 -> out [B,Fo,To,C]. */
int eat_stem_fwd(const float* x, const float* w, void* out, int out_dtype, int B, int F, int T, int C, int stride,
                 const float* scale, const float* shift, int act, double* stat_sum, double* stat_sq,
                 cudaStream_t stream);

/* Depthwise weights [C,1,k,k] -> tap-major [k*k][C]. */
int eat_dw_repack(const float* w, float* wt, int C, int k, cudaStream_t stream);

/* Depthwise k x k conv (k in {3,5}, stride in {1,2}, pad (k-1)/2) + BN/act (+ SE squeeze sums).
 * Replaces the depthwise ConvNormActivation at models/mn/block_types.py:150-162 and the mean of
 * SqueezeExcitation._scale :73 (pool [B,C] += per-sample channel sums of the result). */
int eat_dw_conv_fwd(const void* in, const float* wt, void* out, int dtype, int B, int F, int T, int C, int k,
                    int stride, const float* in_scale, const float* in_shift, int in_act, const float* scale,
                    const float* shift, int act, float* pool, double* stat_sum, double* stat_sq,
                    cudaStream_t stream);

/* Eval-mode BatchNorm folding: scale = gamma / sqrt(rvar + eps), shift = beta - rmean * scale. */
int eat_bn_fold(const float* gamma, const float* beta, const float* rmean, const float* rvar, float eps,
                float* scale, float* shift, int C, cudaStream_t stream);

/* Training-mode BatchNorm: batch statistics -> scale/shift, saved mean/invstd, running-stat update
 * (nn.BatchNorm2d(eps=1e-3, momentum=0.01), models/mn/model.py:114-115). rmean/rvar/nbt may be NULL. */
int eat_bn_finalize(const double* sum, const double* sq, double count, const float* gamma, const float* beta,
                    float eps, float momentum, float* rmean, float* rvar, long long* nbt, float* scale,
                    float* shift, float* save_mean, float* save_invstd, int C, cudaStream_t stream);

/* y = act(z * scale + shift) (+ res), elementwise on [rows, C]  (BN apply + residual,
 * models/mn/block_types.py:177-181). */
int eat_bn_apply(const void* z, const float* scale, const float* shift, int act, const void* res, void* y,
                 int dtype, long long rows, int C, cudaStream_t stream);

/* pool[b,c] += mul * sum_p act(z[b,p,c] * scale[c] + shift[c])   (SE squeeze / global average pool,
 * models/mn/block_types.py:73, models/mn/model.py:220). */
int eat_bn_act_pool(const void* z, const float* scale, const float* shift, int act, float* pool, float mul,
                    int dtype, int B, int P, int C, cudaStream_t stream);

/* Squeeze-excitation MLP: gate = sigmoid(W2 relu(W1 (pool*inv_count) + b1) + b2)
 * (models/mn/block_types.py:72-83).  hidden_out [B,S] optional (saved for backward). */
int eat_se_fc_fwd(const float* pool, float inv_count, const float* w1, const float* b1, const float* w2,
                  const float* b2, float* gate, float* hidden_out, int B, int C, int S, cudaStream_t stream);

/* Exact-fp32 CUDA-core GEMM: C[M,N] = epi(xf(A)[M,K] . W[N,K]^T).  1x1 convs on NHWC rows
 * (models/mn/block_types.py:140-147,167-171) and the classifier Linear layers
 * (models/mn/model.py:187-194).  gate [B,K]: SE gate of sample row/rows_per_sample.
 * w_trans = 1 reads W as [K,N] (data gradient: dA[M,Cin] = G[M,Cout] . W[Cout,Cin]). */
int eat_gemm_simt_fwd(const void* A, int a_dtype, const float* W, int w_trans, void* C, int c_dtype, long long M,
                      int N, int K, const float* in_scale, const float* in_shift, int in_act, const float* gate,
                      int rows_per_sample, const float* scale, const float* shift, int act, const void* residual,
                      double* stat_sum, double* stat_sq, cudaStream_t stream);

/* tcgen05 tensor-core version of the same GEMM contract (w_trans must be 0; A and C share the dtype;
 * K, N multiples of 8).  bf16 storage: one bf16 MMA per product; fp32 storage: hi/lo split, three MMAs
 * (fp32-grade products, ~2^-16 relative).  Same reference call sites as eat_gemm_simt_fwd. */
int eat_pw_tc_fwd(const void* A, int a_dtype, const float* W, int w_trans, void* C, int c_dtype, long long M, int N,
                  int K, const float* in_scale, const float* in_shift, int in_act, const float* gate,
                  int rows_per_sample, const float* scale, const float* shift, int act, const void* residual,
                  double* stat_sum, double* stat_sq, cudaStream_t stream);
/* fp32-storage version of the same contract, fed by TMA (cp.async.bulk.tensor tiles of A, W and the residual) with
 * TF32x3 products (hi/lo split on chip, ~2^-19 relative) and TMA stores; w_trans = 0, K and N multiples of 4.
 * The residual is accumulated before shift/activation, so residual != NULL requires act == EAT_ACT_NONE (the only
 * combination the reference has: block_types.py:167-171,179-180).  w_ws (optional, 128-byte aligned, at least
 * N * ceil(K/32) * 128 bytes): scratch into which the weights are pre-split once per launch (bf16 hi|lo rows, epilogue
 * scale folded, transposed when w_trans = 1, i.e. W given as [K, N]) so that no CTA repeats that work per tile; without
 * it the split happens on chip per tile and w_trans must be 0.  eat_pw_tc_fwd forwards fp32 launches here (without
 * workspace) unless the environment says EAT_PW_IMPL=tc. */
int eat_pw_tma_fwd(const float* A, const float* W, int w_trans, float* C, long long M, int N, int K,
                   const float* in_scale, const float* in_shift, int in_act, const float* gate, int rows_per_sample,
                   const float* scale, const float* shift, int act, const float* residual, double* stat_sum,
                   double* stat_sq, void* w_ws, long long w_ws_bytes, cudaStream_t stream);
/* DynamicConv 1x1 (models/dymn/dy_block.py:103-131) on the same TMA kernel: W = dyn_k kernels [dyn_k][N][K] ([dyn_k][K][N]
 * with w_trans = 1), sample b uses sum_j att[b, j] * W[j]; the per-sample kernels are mixed + pre-split once per launch
 * into w_ws (>= B * N * ceil(K/32) * 128 bytes, 128-byte aligned).  M = B * rows_per_sample. */
int eat_pw_tma_dyn_fwd(const float* A, const float* W, const float* att, int dyn_k, int w_trans, float* C, long long M,
                       int N, int K, int rows_per_sample, const float* scale, const float* shift, int act,
                       const float* residual, double* stat_sum, double* stat_sq, void* w_ws, long long w_ws_bytes,
                       cudaStream_t stream);
/* out[cols, rows] = in[rows, cols]^T (fp32); used to feed W^T to the data-gradient GEMM. */
int eat_transpose_f32(const float* in, float* out, int rows, int cols, cudaStream_t stream);

/* ---- DyMN (reference models/dymn/dy_block.py) ---- */

/* DynamicConv 1x1 on tensor cores (dy_block.py:103-131): W = dyn_k kernels [dyn_k][N][K]; sample b uses
 * sum_k att[b,k]*W[k], mixed while the weight tile is staged (never materialised).  M = B*rows_per_sample. */
int eat_pw_tc_dyn_fwd(const void* A, int dtype, const float* W, const float* att, int dyn_k, void* C, long long M,
                      int N, int K, int rows_per_sample, const float* in_scale, const float* in_shift, int in_act,
                      const float* scale, const float* shift, int act, const void* residual, double* stat_sum,
                      double* stat_sq, cudaStream_t stream);
/* DynamicConv depthwise + BN affine + DyReLU-B (dy_block.py:172-188) + CoordAtt (:195-201) in one kernel.
 * wt: per-sample tap-major weight tables (stride wt_bstride floats); theta [B,C,4] = sigmoid(coef_net(h_c));
 * lam/init: DyReLU buffers; ca_f [B,Fo,C], ca_t [B,To,C] = sigmoid(g_cf), sigmoid(g_ct). */
int eat_dw_conv_fwd_dy(const void* in, const float* wt, long long wt_bstride, void* out, int dtype, int B, int F, int T,
                       int C, int k, int stride, const float* in_scale, const float* in_shift, int in_act,
                       const float* scale, const float* shift, const float* theta, const float* lam,
                       const float* init, const float* ca_f, const float* ca_t, double* stat_sum, double* stat_sq,
                       cudaStream_t stream);
/* ContextGen pooling (dy_block.py:236-240): out [B, F+T, C] fp32 = [mean over T | mean over F]. */
int eat_ctx_pool(const void* x, int dtype, float* out, int B, int F, int T, int C, cudaStream_t stream);
/* Sequence pooling of ContextGen (dy_block.py:227-233,249): AvgPool(3, stride, pad 1) or copy (stride 1) of rows
 * [row0, row0+L) of each sample of in [B, Ltot, H] -> out [B, Lo, H]. */
int eat_seq_pool(const float* in, float* out, int B, int Ltot, int row0, int L, int H, int stride, const float* scale,
                 const float* shift, int act, cudaStream_t stream);   /* optional affine+act applied to `in` on load */
/* att [B,k] = softmax((Wr h_c + br)/temperature)  (dy_block.py:104-107). */
int eat_dyconv_att(const float* hc, const float* wr, const float* br, float temperature, float* att, int B, int H, int k,
                   cudaStream_t stream);
/* per-sample depthwise weight tables wt [B][ksize^2][C] = sum_j att[b,j] * W[j] (dy_block.py:111-117). */
int eat_dyconv_mix_dw(const float* w, const float* att, float* wt, int B, int C, int ksize, int k, cudaStream_t stream);

/* DyMN training path.  p = DyReLU-B(BN(z)) * ca_f * ca_t materialised for the projection conv, and its backward:
 * du = d/d(BN output), dcaf [B,Fo,C] / dcoef [B,C,4] accumulated (caller zeroes), dcat [B,To,C] overwritten. */
int eat_dy_act_fwd(const void* z, void* out, int dtype, const float* scale, const float* shift, const float* theta,
                   const float* lam, const float* init, const float* ca_f, const float* ca_t, int B, int Fo, int To,
                   int C, cudaStream_t stream);
int eat_dy_act_bwd(const void* dp, const void* z, void* du, int dtype, const float* scale, const float* shift,
                   const float* theta, const float* lam, const float* init, const float* ca_f, const float* ca_t,
                   float* dcaf, float* dcat, float* dcoef, int B, int Fo, int To, int C, cudaStream_t stream);
/* dpre = dcoef * lam * 2 s (1-s) with s = theta (DyReLU coefficient net); out = g * s * (1-s) (sigmoid backward). */
int eat_dyrelu_coef_bwd(const float* dcoef, const float* theta, const float* lam, float* dpre, long long n,
                        cudaStream_t stream);
int eat_sigmoid_bwd(const float* g, const float* s, float* out, long long n, cudaStream_t stream);
/* softmax(Linear(h_c)/T) backward: dWr/dbr accumulated, dh_c[b,:] += dlogit . Wr. */
int eat_dyconv_att_bwd(const float* datt, const float* att, float temperature, const float* hc, const float* wr,
                       float* dwr, float* dbr, float* dhc, int B, int H, int k, cudaStream_t stream);
int eat_seq_pool_bwd(const float* dout, float* dsrc, int B, int Ltot, int row0, int L, int H, int stride,
                     cudaStream_t stream);
int eat_ctx_pool_bwd(const float* dg, void* dx, int dtype, int B, int F, int T, int C, cudaStream_t stream);
/* Per-sample weight gradients S [B,N,K] of a 1x1 conv on tensor cores, and the DynamicConv bank / attention
 * gradients derived from per-sample gradients S [B,n]: dW[k] += sum_b att[b,k] S[b]; datt[b,k] = <S[b], W[k]>. */
int eat_pw_tc_wgrad_persample(const void* G, const void* A, int dtype, float* S, long long M, int N, int K,
                              int rows_per_sample, cudaStream_t stream);
int eat_dyn_wgrad_mix(const float* S, const float* att, const float* W, float* dW, float* datt, int B, long long n, int k,
                      cudaStream_t stream);

/* ---- backward (training step: ex_audioset.py:197 loss.backward() over the modules above) ---- */

/* Weight gradient of a 1x1 conv / Linear: dW[N,K] += G[M,N]^T . xf(A)[M,K]; db[N] += colsum(G) (db may be
 * NULL).  dW/db are fp32 and must be zeroed by the caller once per step (atomically accumulated). */
int eat_gemm_simt_wgrad(const void* G, int g_dtype, const void* A, int a_dtype, float* dW, float* db, long long M,
                        int N, int K, const float* in_scale, const float* in_shift, int in_act, const float* gate,
                        int rows_per_sample, cudaStream_t stream);

/* tcgen05 version of the weight gradient (db must be NULL; G and A share the dtype; K, N multiples of 8):
 * MN-major UMMA operands, reduction over pixels split across CTAs, vector atomics into dW. */
int eat_pw_tc_wgrad(const void* G, int g_dtype, const void* A, int a_dtype, float* dW, float* db, long long M, int N,
                    int K, const float* in_scale, const float* in_shift, int in_act, const float* gate,
                    int rows_per_sample, cudaStream_t stream);

/* fp32-storage successor of eat_pw_tc_wgrad, fed by TMA (cp.async.bulk.tensor boxes of G and X, bf16 hi/lo split in
 * place on chip, MN-major UMMA, two CTAs per SM).  per_sample != 0: S[b] = G_b^T . X_b for DynamicConv (M = B *
 * rows_per_sample, dW = S [B, N, K]).  eat_pw_tc_wgrad / eat_pw_tc_wgrad_persample forward fp32 launches here unless the
 * environment says EAT_WG_IMPL=tc. */
int eat_pw_tma_wgrad(const float* G, const float* X, float* dW, long long M, int N, int K, const float* in_scale,
                     const float* in_shift, int in_act, const float* gate, int rows_per_sample, int per_sample,
                     cudaStream_t stream);

/* BatchNorm backward, pass 1: s1[c] += sum dy, s2[c] += sum dy*xhat with dy = g * act'(z*scale+shift),
 * g = gA * gate[b,c] + dpool[b,c] (gA / gate / dpool each optional).  z, gA: [B, P, C]. */
int eat_bn_bwd_reduce(const void* gA, const float* gate, const float* dpool, const void* z, const float* scale,
                      const float* shift, const float* mean, const float* invstd, int act, int dtype, int B, int P,
                      int C, double* s1, double* s2, cudaStream_t stream);
/* dgamma += s2, dbeta += s1 (either may be NULL), c1 = s1/count, c2 = s2/count. */
int eat_bn_bwd_finalize(const double* s1, const double* s2, double count, float* dgamma, float* dbeta, float* c1,
                        float* c2, int C, cudaStream_t stream);
/* pass 2: dz = scale * (dy - c1 - xhat * c2). */
int eat_bn_bwd_apply(const void* gA, const float* gate, const float* dpool, const void* z, const float* scale,
                     const float* shift, const float* mean, const float* invstd, int act, const float* c1,
                     const float* c2, void* dz, int dtype, int B, int P, int C, cudaStream_t stream);

/* SE backward: dgate[b,c] += sum_p dp[b,p,c] * act(z[b,p,c]*scale[c]+shift[c]). */
int eat_se_bwd_reduce(const void* dp, const void* z, const float* scale, const float* shift, int act, float* dgate,
                      int dtype, int B, int P, int C, cudaStream_t stream);
/* SE block (block_types.py:72-83,177-181; autograd of `scale * input` and of the BatchNorm in front of it): the
 * squeeze-excitation reduce and the BatchNorm-backward reduce of the depthwise output in ONE pass over (dp, z).
 * dgate as eat_se_bwd_reduce; part [parts][4][B][C] fp32 (no zero fill needed) receives, per slice of the pixels,
 * sum dp*act'(v), sum dp*act'(v)*(z-mean), sum act'(v), sum act'(v)*(z-mean) with v = z*scale+shift.  Grid = parts x B. */
int eat_se_bn_bwd_reduce(const void* dp, const void* z, const float* scale, const float* shift, const float* mean, int act,
                         float* dgate, float* part, int parts, int dtype, int B, int P, int C, cudaStream_t stream);
/* ... and, once the SE MLP backward has produced dpool: s1[c] += sum_b gate*part0 + dpool*part2,
 * s2[c] += invstd[c] * sum_b gate*part1 + dpool*part3 -- the sums eat_bn_bwd_reduce(gA = dp, gate, dpool) yields. */
int eat_se_bn_bwd_combine(const float* part, int parts, const float* gate, const float* dpool, const float* invstd, int B,
                          int C, double* s1, double* s2, cudaStream_t stream);
/* SE MLP backward (per sample): du2 = dgate*gate*(1-gate), du1 = (W2^T du2)*(hidden>0),
 * dpool = (W1^T du1) * inv_count.  Weight gradients follow from du2/du1 via eat_gemm_simt_wgrad. */
int eat_se_fc_bwd(const float* dgate, const float* gate, const float* hidden, const float* w1, const float* w2,
                  float inv_count, float* du2, float* du1, float* dpool, int B, int C, int S, cudaStream_t stream);

/* Depthwise conv backward: data gradient (+ optional residual add into din) and weight gradient
 * (dw [C,1,k,k] fp32, atomically accumulated; in may carry the producing layer's BN+act as in_*). */
int eat_dw_conv_dgrad(const void* dz, const float* wt, long long wt_bstride, const void* res, void* din, int dtype, int B,
                      int F, int T, int C, int k, int stride, cudaStream_t stream);
int eat_dw_conv_dgrad_s1(const void* dz, const float* wt, long long wt_bstride, const void* res, void* din, int dtype,
                         int B, int F, int T, int C, int k, cudaStream_t stream);   /* stride-1 fast path */
/* Stride-2 data gradient whose output din [B,F,T,C] is the upstream gradient of a BatchNorm + activation with raw input
 * z [B,F,T,C] (the expand stage of an InvertedResidual, block_types.py:140-147): the BatchNorm-backward reduce
 * (s1[c] += sum g, s2[c] += invstd[c] * sum g*(z-mean[c]), g = din * act'(z*scale+shift) -- what eat_bn_bwd_reduce(gA = din)
 * yields) is taken in the epilogue while din is still in registers.  fp32 storage, k in {3,5}; anything else returns
 * EAT_ERR_UNSUPPORTED and the caller runs eat_dw_conv_dgrad + eat_bn_bwd_reduce. */
int eat_dw_conv_dgrad_bnred(const void* dz, const float* wt, const void* res, void* din, const void* z, const float* zscale,
                            const float* zshift, const float* zmean, const float* zinvstd, int zact, double* s1, double* s2,
                            int dtype, int B, int F, int T, int C, int k, int stride, cudaStream_t stream);
int eat_dw_conv_wgrad(const void* dz, const void* in, const float* in_scale, const float* in_shift, int in_act,
                      float* dw, long long dw_bstride, int dtype, int B, int F, int T, int C, int k, int stride,
                      cudaStream_t stream);   /* wt_bstride / dw_bstride: floats between per-sample tables (0: shared) */
/* Stem weight gradient (the spectrogram itself needs no gradient). */
int eat_stem_wgrad(const void* dz, int dtype, const float* x, float* dw, int B, int F, int T, int C, int stride,
                   cudaStream_t stream);
/* dpre = dh * mask * act'(pre), fp32 vectors (classifier Hardswish + Dropout backward). */
int eat_act_bwd(const float* dh, const float* pre, const float* mask, int act, float* dpre, long long n,
                cudaStream_t stream);

/* ---- training step around the network (ex_audioset.py:135-199) ---- */

/* Spectrogram mixup, ex_audioset.py:145-146: out[b] = x[b]*lam[b] + x[perm[b]]*(1-lam[b]). */
int eat_mixup(const float* x, const int* perm, const float* lam, float* out, int B, long long per_sample,
              cudaStream_t stream);
/* Hard-label + knowledge-distillation BCE-with-logits and its gradient, ex_audioset.py:149-189.
 * loss_acc (fp64[2], caller-zeroed) += {kd*label_loss, (1-kd)*distillation_loss}; teacher/perm/lam optional
 * (NULL teacher: plain BCE, weight 1).  teacher_known [B] (optional, 1/0): 0 zeroes the distillation loss of a clip
 * without teacher predictions, ex_audioset.py:166-178.  dlogits = d(total loss)/d(logits), may be NULL. */
int eat_bce_kd_loss(const float* logits, const float* y, const float* teacher, const float* teacher_known,
                    const int* perm, const float* lam, float kd_lambda, int B, int C, float* dlogits,
                    double* loss_acc, cudaStream_t stream);
/* torch.optim.Adam (adamw = 0) / AdamW (adamw = 1) on flat fp32 arenas, ex_audioset.py:86-91,198;
 * grad_scale multiplies the gradient first (1/world_size after a sum all-reduce). step counts from 1. */
int eat_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int adamw, int step, float grad_scale, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EAT_B200_H */
