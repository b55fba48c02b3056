/*
 * ctl_b200.h -- C ABI of the B200-native cooperative-training hot path.
 *
 * Drop-in boundary.  The reference has no FFI: its boundary is four Python
 * callables.  This library exports the device work those callables perform; the Python mirror in
 * cooperative_training_and_latent_space_data_augmentation_b200/ binds it with ctypes and keeps the
 * reference's names, argument order, defaults and exceptions.  INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; buffers are caller-allocated
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are asynchronous
 *   - return value: CTL_OK or an error code; ctl_last_error() gives a thread-local message
 *   - no allocation, no global state, no CPU fallback: without a CUDA device every compute entry
 *     point returns CTL_ERR_CUDA
 *   - tensors are dense NCHW (masking) exactly as the reference passes them
 */
#ifndef CTL_B200_H_
#define CTL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTL_B200_VERSION 201 /* 0.2.1: + ctl_bn_apply_from_sums_c8 */

enum ctl_status {
  CTL_OK = 0,
  CTL_ERR_INVALID = 1,     /* bad argument (NULL pointer, non-positive size, unknown enum)          */
  CTL_ERR_INDEX = 2,       /* k >= n: the reference raises IndexError (model_util.py:231-232)        */
  CTL_ERR_UNSUPPORTED = 3, /* shape outside what the kernels handle (documented per entry point)     */
  CTL_ERR_CUDA = 4         /* CUDA runtime error (message in ctl_last_error)                         */
};

enum ctl_dtype { CTL_F32 = 0, CTL_BF16 = 1 };
enum ctl_mode { CTL_MODE_CHANNEL = 0, CTL_MODE_SPATIAL = 1 };
enum ctl_act { CTL_ACT_NONE = 0, CTL_ACT_LRELU = 1 /* slope 0.2 */, CTL_ACT_RELU = 2, CTL_ACT_SIGMOID = 3 };

int ctl_version(void);
const char* ctl_last_error(void);
/* number of SMs of the current device (grid sizing is derived from it); <0 on error */
int ctl_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1 -- latent saliency.  Replaces
 *   torch.mean(gradient.view(N, C, -1), dim=2)              medseg/models/model_util.py:224-225
 *   torch.mean(gradient, dim=1, keepdim=True).squeeze()...  medseg/models/model_util.py:285-286
 * g: [N,C,HW] (g_dtype), s_out: fp32 [N,C] (channel) or [N,HW] (spatial).
 * Accumulates in fp64 and rounds once (order-independent; see oracle/masking_oracle.py).
 */
int ctl_saliency_reduce(const void* g, int g_dtype, int64_t N, int64_t C, int64_t HW, int mode,
                        float* s_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2 -- per-sample top-p threshold, mask build and 128-bit apply.  Replaces
 *   torch.sort(s, descending=True)[0][:, k]; torch.where(s > thr, 0.5*rand_like(s) | 0, 1);
 *   code * mask.view(N,C,1,1) | (N,1,H,W)       medseg/models/model_util.py:231-249, :293-312
 * s: fp32 [N,n], n = C (channel) or HW (spatial).  k = int(n*p) is computed by the caller on the host
 * exactly as the reference does; k >= n returns CTL_ERR_INDEX, k < 0 CTL_ERR_INVALID.
 * soft != 0: masked entries get 0.5*u with u = rand[i,j] when `rand` is non-NULL (the caller's
 * torch.rand_like(s) draw -- reference-compatible mode), else u = Philox4x32-10(seed, offset,
 * (first_sample+i)*n + j) (native, shard-invariant mode; first_sample = global index of row 0).
 * mask_out: fp32 [N,n] (viewed by the caller as [N,C,1,1] or [N,1,H,W]); thr_out: fp32 [N] or NULL.
 * z: [N,C,HW] (z_dtype) -> z_out same shape (out_dtype; the reference always produces fp32).
 * Two launches: a one-CTA-per-sample select kernel, then the streaming apply kernel.
 */
int ctl_topp_mask_apply(const float* s, const void* z, int z_dtype, int64_t N, int64_t C, int64_t HW,
                        int mode, int64_t k, int soft, const float* rand, uint64_t seed,
                        uint64_t offset, int64_t first_sample, float* mask_out, float* thr_out,
                        void* z_out, int out_dtype, void* stream);

/* The whole tail of mask_latent_code_{channel,spatial}_wise as one chain of three launches on `stream`:
 * K1 (saliency), a one-CTA-per-sample select/mask-build kernel and K2 (128-bit streaming apply).  The
 * second and third are launched with programmatic stream serialization, so K2's loads of z are already
 * in flight while K1 drains and the select runs.  s_scratch: fp32 [N,n] caller scratch, holds s afterwards.
 */
int ctl_saliency_mask_apply(const void* g, int g_dtype, const void* z, int z_dtype, int64_t N,
                            int64_t C, int64_t HW, int mode, int64_t k, int soft, const float* rand,
                            uint64_t seed, uint64_t offset, int64_t first_sample, float* s_scratch,
                            float* mask_out, float* thr_out, void* z_out, int out_dtype,
                            void* stream);

/* ---------------------------------------------------------------------------------------------
 * Random channel dropout.  Replaces F.dropout2d(z, p) + the full-size `where(masked == z, 1, 0)`
 *   medseg/models/advanced_triplet_recon_segmentation_model.py:332-336
 * keep: fp32 [N,C] of 0/1 (the caller's bernoulli_(1-p) draw) or NULL for native Philox
 * (keep <=> u >= p).  z_out = z * (keep * scale), scale = fp32(1/(1-p)) supplied by the caller so the
 * host controls the rounding (p == 1: scale 0; p == 0: pass scale 1 and keep NULL/ones).
 * mask_out: fp32 [N,C,HW] = (z_out == z) or NULL to skip the reference's quirk mask.
 * keep_out: fp32 [N,C] or NULL (the drawn pattern, for tests).
 */
int ctl_channel_dropout(const void* z, int z_dtype, int64_t N, int64_t C, int64_t HW, float p,
                        float scale, const float* keep, uint64_t seed, uint64_t offset,
                        int64_t first_sample, void* z_out, int out_dtype, float* mask_out,
                        float* keep_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CUDA-graph replay forms of the two entry points above.  A captured launch freezes its by-value
 * arguments, but the reference draws a new percentile (hence k) and new random numbers every step
 *   medseg/models/model_util.py:224-231 (np.random.rand() * percentile -> k), :238-241 (rand_like)
 * so these variants read the per-step values from DEVICE memory at run time:
 *   step_params: int64 [3] = {k, philox offset, global index of this rank's first sample}
 * (k is ignored by the dropout form).  The host validates k < n before it uploads the values -- the
 * reference's IndexError is raised there -- and the kernel clamps k into [0, n-1].
 */
int ctl_saliency_mask_apply_dyn(const void* g, int g_dtype, const void* z, int z_dtype, int64_t N,
                                int64_t C, int64_t HW, int mode, int soft, const float* rand,
                                uint64_t seed, const int64_t* step_params, float* s_scratch,
                                float* mask_out, float* thr_out, void* z_out, int out_dtype,
                                void* stream);
int ctl_channel_dropout_dyn(const void* z, int z_dtype, int64_t N, int64_t C, int64_t HW, float p,
                            float scale, uint64_t seed, const int64_t* step_params, void* z_out,
                            int out_dtype, float* mask_out, float* keep_out, void* stream);

/* Philox4x32-10 uniform draw, exposed for parity tests of the native RNG: out[i] = u(first_index+i). */
int ctl_philox_uniform(uint64_t seed, uint64_t offset, uint64_t first_index, int64_t count,
                       float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3 -- conv blocks of the FTN/STN encoders/decoders as bf16 implicit GEMM on tcgen05 tensor cores
 * (TMA-fed, TMEM accumulators).  Replaces the nn.Conv2d / ConvTranspose2d(k2,s2) + BatchNorm2d (folded:
 * eval running statistics, or batch statistics computed by ctl_bn_batch_affine_c8) + LeakyReLU(0.2) | ReLU |
 * Sigmoid + residual-add sequences of
 *   res_convdown   medseg/models/ebm/encoder_decoder.py:19-68
 *   res_up_family  medseg/models/ebm/encoder_decoder.py:285-348
 *   MyEncoder / MyDecoder / Dual_Branch_Encoder  :351-415, :418-453, :456-503
 *
 * Activation layout "C8": bf16 [N][C/8][H][W][8] (channels in groups of eight, each group a dense plane of
 * 16-byte pixels) -- one TMA box per tile lands in shared memory in the UMMA canonical layout.
 * taps = 9: 3x3, padding 1; taps = 1: 1x1.  subsample = 2: the 3x3 stride-2 padding-1 convolution of `down`
 * (output H/2 x W/2).  up2x = 1 (taps 1): ConvTranspose2d(kernel 2, stride 2): Cout = 4*out_channels GEMM
 * columns ordered (dy*2+dx)*out_channels + co, scattered to output pixel (2y+dy, 2x+dx) of [N][oc/8][2H][2W][8].
 * w_packed, NT = ctl_conv2d_n_tile(Cin, Cout, taps):
 *   ctl_conv2d_vpacked(...) == 0 ("tap-major": 1x1, 3x3 stride 2, 3x3 with Cin <= 64): bf16 [Cout/NT][taps][Cin/8][NT][8], element
 *           (t, tap, q, n, j) = weight[t*NT + n][q*8 + j][tap/3][tap%3];
 *   ctl_conv2d_vpacked(...) == 1 ("vertically packed": the three vertical taps of a filter column share one MMA, N = 3*NT):
 *           bf16 [Cout/NT][3][Cin/8][3*NT][8], element (t, s, q, r*NT + n, j) = weight[t*NT + n][q*8 + j][r][s].
 * out = act( conv(x) * scale[c] + shift[c] + (res * res_scale[c] + res_shift[c]) ); res has the output's shape
 * and layout; res and every per-GEMM-column fp32 vector may be NULL (identity).
 * Cin in {16,32,64,128}, Cout %% 16 == 0.
 * stats (may be NULL): double [2][Cout], the per-channel sum and sum of squares of the STORED (bf16-rounded) outputs are
 * accumulated into it by the epilogue (zero it first) -- train-mode BatchNorm statistics without a second pass over the
 * tensor; needs ctl_conv2d_n_tile(...) <= 32 and up2x == 0.  Finalise with ctl_bn_affine_from_sums.
 */
int ctl_conv2d_n_tile(int Cin, int Cout, int taps);
/* 1 when ctl_conv2d_c8_bf16 expects the vertically packed 3x3 layout for this layer class (today: 3x3, stride 1, 128 input
 * channels), 0 for the tap-major layout. */
int ctl_conv2d_vpacked(int Cin, int Cout, int taps, int subsample);
/* fp32 nn.Conv2d weight [Cout][Cin][k][k] (taps = k*k) -> the packed bf16 w_packed above.  transposed = 1 packs the weight of
 * the INPUT-gradient convolution instead, w'[ci][co][r][s] = w[co][ci][k-1-r][k-1-s] (its Cout' = Cin, Cin' = Cout).
 * tap_major = 1: tap-major layout; 0: vertically packed 3x3 layout -- pass !ctl_conv2d_vpacked(Cin', Cout', taps, subsample). */
int ctl_pack_conv_weight(const float* weight, int64_t Cout, int64_t Cin, int taps, int transposed, int tap_major, void* out,
                         void* stream);
/* The same for many weights in ONE launch (a training step repacks every conv weight after the optimizers ran).
 * jobs: DEVICE int64 [n_jobs][8] = {weight ptr, out ptr, Cout, Cin, taps, ctl_conv2d_n_tile of the packed view,
 * transposed, tap_major}; max_elements = the largest Cout*Cin*taps among them (grid sizing). */
int ctl_pack_conv_weights_batched(const int64_t* jobs, int64_t n_jobs, int64_t max_elements, void* stream);
int ctl_conv2d_c8_bf16(const void* x, int64_t N, int64_t H, int64_t W, int64_t Cin, const void* w_packed,
                       int64_t Cout, int taps, int subsample, int up2x, const float* scale,
                       const float* shift, const void* res, const float* res_scale,
                       const float* res_shift, int act, void* out, double* stats, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CUDA-core kernels around K3 (all HBM-bound streaming passes over C8 tensors).
 */
/* NCHW (fp32 | bf16) <-> C8 bf16; C %% 8 == 0.  Used for the latent codes, which the masking API holds in NCHW. */
int ctl_nchw_to_c8(const void* x, int x_dtype, int64_t N, int64_t C, int64_t H, int64_t W, void* y, void* stream);
int ctl_c8_to_nchw(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, void* y, int y_dtype, void* stream);
/* Stem: 3x3 pad-1 conv from a planar fp32 NCHW input with Cin in {1,4} to 16 C8 channels, y = act(acc*scale+shift)
 * (MyEncoder.inc[0] + norm + LeakyReLU, encoder_decoder.py:370-378).  in_mode 0: x as is; 1: softmax(x/temperature)
 * over the Cin channels; 2: one-hot of the int64 label map `labels` [N,H,W] -- modes 1/2 fuse construct_input
 * (basic_operations.py:110-158) into the first STN convolution.  weight: fp32 [16][Cin][3][3]. */
int ctl_stem_conv3x3_c8(const float* x, const int64_t* labels, int in_mode, float temperature, int64_t N,
                        int64_t Cin, int64_t H, int64_t W, const float* weight, int64_t Cout,
                        const float* scale, const float* shift, int act, void* y, void* stream);
/* Head: 1x1 conv from 16 C8 channels to Cout in [1,4] planar fp32 NCHW (+ bias, optional sigmoid)
 * (MyDecoder.final_conv + last_act, encoder_decoder.py:439-452).  weight: fp32 [Cout][16]. */
int ctl_head_conv1x1_c8(const void* x, int64_t N, int64_t Cin, int64_t H, int64_t W, const float* weight,
                        const float* bias, int64_t Cout, int act, float* y, void* stream);
/* nn.UpsamplingNearest2d(scale_factor=2) on a C8 tensor (encoder_decoder.py:294-296). */
int ctl_upsample2x_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, void* y, void* stream);
/* Train-mode BatchNorm2d statistics of a C8 tensor folded into y = x*scale + shift:
 * scale = gamma*rsqrt(var_biased + eps), shift = beta - mean*scale (fp64 accumulation).  When running_mean/var are
 * given they are updated like nn.BatchNorm2d (momentum, unbiased variance) -- pass NULL to reproduce
 * _disable_tracking_bn_stats (model_util.py:414-451).  mean_out / var_out (biased) are what the backward needs.
 * workspace: ctl_bn_workspace_bytes(N, C) bytes. */
size_t ctl_bn_workspace_bytes(int64_t N, int64_t C);
int ctl_bn_batch_affine_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, const float* gamma,
                           const float* beta, float eps, void* workspace, float* scale, float* shift,
                           float* mean_out, float* var_out, float* running_mean, float* running_var,
                           float momentum, void* stream);
/* ctl_bn_batch_affine_c8 from sums accumulated by the conv epilogue (sums: double [2][C], count = N*H*W of that tensor) */
int ctl_bn_affine_from_sums(const double* sums, int64_t C, int64_t count, const float* gamma, const float* beta, float eps,
                            float* scale, float* shift, float* mean_out, float* var_out, float* running_mean,
                            float* running_var, float momentum, void* stream);
/* y = act(x*scale[c] + shift[c]) on C8 tensors (BatchNorm apply + LeakyReLU in one pass). */
int ctl_scale_shift_act_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, const float* scale,
                           const float* shift, int act, void* y, void* stream);
/* y = act(x*scale[c] + shift[c] + up2(low)): x, y C8 [N,C/8,H,W,8], low C8 [N,C/8,H/2,W/2,8] (nearest x2 on the fly).
 * Tail of res_up_family with nn.UpsamplingNearest2d (encoder_decoder.py:294-296, :334-337) once the 1x1 shortcut is taken
 * at the low resolution: conv1x1(up(x)) == up(conv1x1(x)).  scale / shift may be NULL (1 / 0); H, W even. */
int ctl_scale_shift_upadd_act_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, const float* scale,
                                 const float* shift, const void* low, int act, void* y, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward of the conv blocks (the reference leaves all of it to torch autograd over the modules of
 * medseg/models/ebm/encoder_decoder.py:19-68, :285-348, :351-415, :418-453, :456-503).
 * Input gradients of 3x3 / 1x1 convolutions run on ctl_conv2d_c8_bf16 itself with transposed + flipped weights.
 */
/* K3w: dW(tap, ci, co) += sum_p x[p + tap - pad][ci] * dy[p][co] on tcgen05 (both operands MN-major from C8 tiles).
 * x: C8 [N,Cin/8,H,W,8], dy: C8 [N,Cout/8,H,W,8] (same H, W; 3x3 pad 1 or 1x1).  dW: fp32, ACCUMULATED into (zero it
 * first); element (tap, ci, co) lives at dW[tap*stride_tap + ci*stride_ci + co*stride_co] -- (1, Cout, Cin*Cout) is the
 * kernel's natural order, (Cin*taps, taps, 1) writes nn.Conv2d's [Cout][Cin][k][k] directly, (4, 4*Cout, -) with dW
 * offset by d writes tap d of a ConvTranspose2d [Cin][Cout][2][2].  Cin in {16,32,64,128}, Cout %% 16 == 0. */
int ctl_conv_wgrad_c8_bf16(const void* x, const void* dy, int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                           int taps, float* dW, int64_t stride_co, int64_t stride_ci, int64_t stride_tap, void* stream);
/* workspace bytes of the per-channel reductions below */
size_t ctl_reduce_workspace_bytes(int64_t N, int64_t C);
/* per-channel sum / sum of squares of a C8 tensor (bias gradients of convolutions without BatchNorm) */
int ctl_channel_sums_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, void* workspace, float* sum_out,
                        float* sumsq_out, void* stream);
/* BatchNorm(train) + activation backward, stage 1.  dv = dy * act'(h) (h = the activation OUTPUT; NULL: dv = dy);
 * reduces sum(dv), sum(dv*a) per channel (a = the BatchNorm input) and emits the coefficients of
 *   da = coef[0][c]*dv + coef[1][c]*a + coef[2][c]      (coef: fp32 [3][C])
 * plus dgamma / dbeta (either may be NULL).  dv_out (C8, may be NULL) materialises dv.  mean / var: the batch statistics
 * of the forward (ctl_bn_batch_affine_c8 mean_out / var_out).  act in {NONE, LRELU, RELU}. */
int ctl_bn_bwd_reduce_c8(const void* dy, const void* h, const void* a, int64_t N, int64_t C, int64_t H, int64_t W,
                         int act, const float* mean, const float* var, float eps, const float* gamma, void* workspace,
                         void* dv_out, float* coef, float* dgamma, float* dbeta, const float* act_scale,
                         const float* act_shift, void* stream);
/* stage 2: da = coef0*dv + coef1*a + coef2 with dv = dy * act'(h) (h NULL: dy is already dv).
 * Both stages: when h is NULL and (act_scale, act_shift) -- the forward's folded BatchNorm affine, fp32 [C] each -- are
 * given, h = act(a*act_scale + act_shift) is not read at all: act' only needs the sign of that expression (LReLU / ReLU),
 * which is recomputed from `a` (one tensor read less per stage). */
int ctl_bn_bwd_apply_c8(const void* dy, const void* h, const void* a, int64_t N, int64_t C, int64_t H, int64_t W,
                        int act, const float* coef, void* da, const float* act_scale, const float* act_shift,
                        void* stream);
/* The two BatchNorm-backward stages in ONE call without a finalisation launch: the reduction leaves sum dv / sum dv*a in
 * `totals` (double [2][C], ZEROED by the caller; one fp64 atomic per value and CTA) and the apply pass forms its
 * coefficients from them in its prologue; dgamma / dbeta (may be NULL) are written by its first CTA.  Arguments as
 * ctl_bn_bwd_reduce_c8 / ctl_bn_bwd_apply_c8; C <= 256. */
int ctl_bn_bwd_c8(const void* dy, const void* h, const void* a, int64_t N, int64_t C, int64_t H, int64_t W, int act,
                  const float* mean, const float* var, float eps, const float* gamma, double* totals, void* dv_out, void* da,
                  float* dgamma, float* dbeta, const float* act_scale, const float* act_shift, void* stream);
/* The BatchNorm-backward REDUCTION fused into the convolution that produces dy (3x3, stride 1, Cin <= 64, N tile <= 32,
 * vertically-unpacked layers): out = conv(x) is an activation gradient that flows into the backward of
 * h = act(BatchNorm(bn_a)); the epilogue reads bn_a at the output's pixels and accumulates totals[0][c] += sum dv,
 * totals[1][c] += sum dv * bn_a with dv = out * act'(bn_a * bn_scale + bn_shift) (double [2][Cout], zeroed by the caller;
 * bn_act = CTL_ACT_LRELU | CTL_ACT_RELU).  ctl_bn_bwd_apply_totals_c8 is the apply pass of ctl_bn_bwd_c8 alone. */
int ctl_conv2d_c8_bf16_bnbwd(const void* x, int64_t N, int64_t H, int64_t W, int64_t Cin, const void* w_packed, int64_t Cout,
                             const void* bn_a, const float* bn_scale, const float* bn_shift, int bn_act, void* out,
                             double* totals, void* stream);
int ctl_bn_bwd_apply_totals_c8(const void* dy, const void* a, int64_t N, int64_t C, int64_t H, int64_t W, int act,
                               const float* mean, const float* var, float eps, const float* gamma, const double* totals,
                               void* da, float* dgamma, float* dbeta, const float* act_scale, const float* act_shift,
                               void* stream);
/* Train-mode BatchNorm finalisation + apply in ONE launch: y = act(x * scale + shift [+ nearest_up2(low)]) where scale / shift
 * are formed in every CTA's prologue from the per-channel sums a convolution epilogue accumulated (sums: double [2][C] =
 * sum x | sum x^2 over N*H*W values); the first CTA also writes scale / shift / mean / var (for the backward) and applies
 * the running-statistics update of nn.BatchNorm2d (running_* may be NULL).  low: NULL, or the C8 tensor
 * [N][C/8][H/2][W/2][8] of ctl_scale_shift_upadd_act_c8.  Replaces ctl_bn_affine_from_sums + ctl_scale_shift[_upadd]_act_c8.
 * C <= 256. */
int ctl_bn_apply_from_sums_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, const double* sums,
                              const float* gamma, const float* beta, float eps, const void* low, int act, void* y,
                              float* scale_out, float* shift_out, float* mean_out, float* var_out, float* running_mean,
                              float* running_var, float momentum, void* stream);
/* dv = dy * act'(h) */
int ctl_act_bwd_c8(const void* dy, const void* h, int64_t N, int64_t C, int64_t H, int64_t W, int act, void* dv,
                   void* stream);
/* backward of nearest x2 up-sampling: dy C8 [N,C/8,2H,2W,8] -> dx C8 [N,C/8,H,W,8] (2x2 sums); H, W = low resolution */
int ctl_downsample2x_sum_c8(const void* dy, int64_t N, int64_t C, int64_t H, int64_t W, void* dx, void* stream);
/* x C8 [N,C/8,H,W,8] -> y C8 [N,C/8,2H,2W,8] with y[2i][2j] = x[i][j], zeros elsewhere (dy of a stride-2 conv at full
 * resolution, so that its input / weight gradients are ordinary stride-1 calls) */
int ctl_zero_stuff2x_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, void* y, void* stream);
/* x C8 [N,C/8,2H,2W,8] -> y C8 [4][N,C/8,H,W,8], y[d] = x[2i + d/2][2j + d%2] (dy of ConvTranspose2d k2 s2 per kernel tap) */
int ctl_split_parity2x2_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, void* y, void* stream);
/* backward of ctl_head_conv1x1_c8: dy planar fp32 [N,Cout,H,W], y = the forward output (sigmoid only, else NULL);
 * dx C8 bf16 [N,2,H,W,8]; dW fp32 [Cout][16] and db fp32 [Cout] are ACCUMULATED into.  Cout in {1,4}. */
int ctl_head_bwd_c8(const float* dy, const float* y, const void* x, int64_t N, int64_t Cin, int64_t H, int64_t W,
                    const float* weight, int64_t Cout, int act, void* dx, float* dW, float* db, void* stream);
/* ---------------------------------------------------------------------------------------------
 * Fused 2-D cross entropy with label-map targets (no mask, no class weights).  Replaces
 *   log_softmax -> NHWC transpose -> nll_loss(sum) -> / region   medseg/models/custom_loss.py:706-741 (via :8-19)
 *   log_softmax -> nll_loss(sum) / (numel + 1e-10)               medseg/models/model_util.py:104-135
 * logits: planar fp32 [N,C,H,W], C in {2,3,4,8}; labels: int64 [N,H,W] (values outside [0,C) are ignored);
 * loss_out[0] = scale * sum over pixels of (logsumexp(x) - x[label]).  workspace16: 16 bytes of device memory that
 * are ZERO on entry and left zero on exit (fp64 sum + CTA ticket), so one buffer serves every call of a stream.
 * ctl_ce2d_bwd: dlogits = grad_out[0] * scale * (softmax(x) - onehot(label)); grad_out is a DEVICE scalar (NULL = 1). */
int ctl_ce2d_fwd(const float* logits, const int64_t* labels, int64_t N, int64_t C, int64_t H, int64_t W, double scale,
                 void* workspace16, float* loss_out, void* stream);
int ctl_ce2d_bwd(const float* logits, const int64_t* labels, int64_t N, int64_t C, int64_t H, int64_t W, double scale,
                 const float* grad_out, float* dlogits, void* stream);
/* The stem's input as a 16-channel C8 bf16 tensor [N,2,H,W,8] (in_mode as above), so that MyEncoder.inc[0] and its
 * weight gradient run on ctl_conv2d_c8_bf16 / ctl_conv_wgrad_c8_bf16 with Cin = 16 -- the tensor-core route of the
 * training path; Cin in {1,4}.  Input channel c is stored as a bf16 pair (fp32 value v): channel c = hi = bf16(v),
 * channel Cin+c = lo = bf16(v - hi), channel 2*Cin+c = hi again; the other channels are zero.  With the weight laid out
 * as [16][16][3][3] = (w_hi | w_hi | w_lo | 0) over those channel groups the convolution keeps ~16 mantissa bits of
 * both operands; the weight gradient is the sum of the first two channel groups of the 16-channel result. */
int ctl_stem_input_c8(const float* x, const int64_t* labels, int in_mode, float temperature, int64_t N, int64_t Cin,
                      int64_t H, int64_t W, void* out, void* stream);
/* backward of ctl_stem_conv3x3_c8 w.r.t. the weight: dy C8 (16 channels, gradient of the raw conv output);
 * x / labels / in_mode / temperature as in the forward; dW fp32 [16][Cin][3][3] ACCUMULATED into. */
int ctl_stem_wgrad_c8(const void* dy, const float* x, const int64_t* labels, int in_mode, float temperature, int64_t N,
                      int64_t Cin, int64_t H, int64_t W, float* dW, void* stream);
/* ... and w.r.t. the input (in_mode 0: plain, 1: through softmax(x / temperature)); dx planar fp32 [N,Cin,H,W] */
int ctl_stem_dgrad_c8(const void* dy, const float* x, int in_mode, float temperature, int64_t N, int64_t Cin, int64_t H,
                      int64_t W, const float* weight, float* dx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-tensor Adam over flat buffers.  Replaces the five optim.Adam(...).step() calls of
 *   medseg/models/advanced_triplet_recon_segmentation_model.py:774-785 (set_optimizers / optimize_all_params),
 *   medseg/train_adv_supervised_segmentation_triplet.py:230-231
 * params / grads / exp_avg / exp_avg_sq: flat fp32 DEVICE buffers of equal length (16-byte aligned, length padded to a
 * multiple of 4); seg_bounds_host: HOST array of n_segments (<= 8) [begin, end) element pairs, begin %% 4 == 0, ascending
 * -- one segment per optimizer of the reference; seg_mask bit s: segment s takes a step; steps: DEVICE fp32
 * [n_segments] step counters, incremented for the stepped segments before use (CUDA-graph replayable).
 * g' = grads*grad_scale (+ weight_decay*p); m += (g'-m)(1-beta1); v = beta2*v + (1-beta2) g'^2;
 * p -= lr/(1-beta1^t) * m / (sqrt(v)/sqrt(1-beta2^t) + eps)      (torch.optim.Adam, amsgrad=False).
 * zero_grad != 0: the gradients of the stepped segments are cleared in the same pass (the next step accumulates into
 * zeros: optimizer.zero_grad() of advanced...model.py:755-758 without another sweep). */
int ctl_adam_flat(float* params, float* grads, float* exp_avg, float* exp_avg_sq, const int64_t* seg_bounds_host,
                  int n_segments, unsigned seg_mask, float* steps, double lr, double beta1, double beta2, double eps,
                  double weight_decay, double grad_scale, int zero_grad, void* stream);
/* ---------------------------------------------------------------------------------------------
 * Sum of squared errors and its gradient (row f2): loss_out[0] = scale * sum (pred - target)^2 over n fp32 elements;
 * dpred = grad_out[0] * 2 * scale * (pred - target).  Replaces 0.5 * nn.MSELoss()(recon, image)
 * (advanced...model.py:443-447, scale = 0.5/n) and torch.mean((decoder(code) - gt)**2) (model_util.py:207-208).
 * workspace16 / grad_out as in ctl_ce2d_fwd / ctl_ce2d_bwd. */
int ctl_sse_fwd(const float* pred, const float* target, int64_t n, double scale, void* workspace16, float* loss_out,
                void* stream);
int ctl_sse_bwd(const float* pred, const float* target, int64_t n, double scale, const float* grad_out, float* dpred,
                void* stream);
/* ---------------------------------------------------------------------------------------------
 * On-device evaluation metric (row f4).  Replaces pred.max(1)[1].cpu().numpy() + runningScore._fast_hist / update /
 * get_scores (medseg/common_utils/metrics.py:18-53, advanced...model.py:656-659).
 * Prediction per pixel = argmax over classes of planar fp32 logits [N,C,HW] (first maximum), or pred_labels int64 [N,HW]
 * when logits is NULL.  gt: int64 [N,HW] (pixels with gt outside [0,C) are skipped, as _fast_hist's mask does);
 * hist: uint64 [C][C] (row = gt, column = prediction) ACCUMULATED into; labels_out: optional uint8 [N,HW] label map.
 * gt == hist == NULL: only the label map is produced.  C in {2,3,4,8}.
 * ctl_confusion_scores: scores_out fp64 [4 + C] = overall acc, mean acc (nanmean), frequency-weighted acc, mean IoU
 * (nanmean), IoU per class (NaN for an absent class). */
int ctl_confusion_update(const float* logits, const int64_t* pred_labels, const int64_t* gt, int64_t N, int64_t C,
                         int64_t HW, void* hist, void* labels_out, void* stream);
int ctl_confusion_scores(const void* hist, int64_t C, double* scores_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Latent saliency reduced ON-CHIP in the epilogue of the decoder's last input-gradient convolution, and the masking tail
 * that consumes it.  Replaces, on the training path, the materialised gradient of
 *   gradient = torch.autograd.grad(loss, [code])  +  torch.mean(gradient...)    medseg/models/model_util.py:217-225, :285-286
 * and the layout conversion in front of decoder_inference (advanced...model.py:396-412).
 * ctl_conv2d_c8_bf16_saliency: the 1x1 convolution (Cin 64 or 128) out = conv(x) + res of ctl_conv2d_c8_bf16 whose
 * epilogue also accumulates (fp64 atomics, buffer zeroed by the caller) per-sample sums of the bf16-rounded outputs:
 * sal_sums[N][Cout] summed over pixels (CTL_MODE_CHANNEL) or sal_sums[N][H*W] summed over channels (CTL_MODE_SPATIAL);
 * store_out == 0: `out` is not written (may be NULL) -- dL/dz never reaches HBM.
 * ctl_saliency_sums_mask_apply: s = fp32(sums / count) (count = HW or C: the value K1 computes from the materialised
 * tensor), per-sample top-k threshold and mask as ctl_topp_mask_apply, z~ = z * mask written as NCHW fp32 (z_out) and,
 * when z_c8_out != NULL, as the blocked bf16 tensor [N][C/8][HW][8] the decoder's first convolution reads.
 * step_params: NULL, or DEVICE int64[3] = {k, offset, first_sample} as in ctl_saliency_mask_apply_dyn.  z: fp32, C %% 8 == 0. */
int ctl_conv2d_c8_bf16_saliency(const void* x, int64_t N, int64_t H, int64_t W, int64_t Cin, const void* w_packed,
                                int64_t Cout, const void* res, void* out, double* sal_sums, int sal_mode, int store_out,
                                void* stream);
int ctl_saliency_sums_mask_apply(const double* sums, const float* z, int64_t N, int64_t C, int64_t HW, int mode, int64_t k,
                                 int soft, const float* rand, uint64_t seed, uint64_t offset, int64_t first_sample,
                                 const int64_t* step_params, float* s_out, float* mask_out, float* thr_out, float* z_out,
                                 void* z_c8_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTL_B200_H_ */
