/* bcp_b200 -- C ABI of the B200-native BCP training-step kernels (libbcp_b200.so).
 *
 * The reference (DeepMed-Lab-ECNU/BCP) has no FFI seam: every operator on its hot path is a PyTorch
 * library call made from Python.  The entry points below are what a binding for that
 * path has to reach; each one names the reference call site(s) it replaces (paths relative to
 * /root/reference/code).  Host code (bcp_b200/*.py) reaches them through ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C symbols, POD arguments, no torch types; every call takes the cudaStream_t to enqueue on
 *   - return 0 on success, <0 on error (-1 invalid argument, -2 unsupported shape, -3 CUDA error);
 *     bcp_last_error() returns a thread-local message.  Never throws, never aborts, never synchronises.
 *   - the CALLER owns all memory, including workspaces (query bcp_*_workspace_* first); no hidden allocation
 *     and no mutable library state beyond immutable per-shape tiling plans memoised behind a mutex and the
 *     thread-local error string, so calls are re-entrant (autograd's backward thread, one process per GPU); kernel
 *     selection never reads the environment
 *   - activations: channel-blocked bf16 "CB8"  [N][ceil(C/8)][X][Y][Z][8]  (2-D nets use X = 1)
 *     network input / logits: planar fp32 [N][C][X][Y][Z] (PyTorch NCDHW); labels: uint8 [N][X][Y][Z]
 *     statistics, master weights, gradients of weights: fp32
 */
#ifndef BCP_B200_H_
#define BCP_B200_H_

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCP_B200_ABI_VERSION 1

const char* bcp_last_error(void);
int bcp_abi_version(void);
int bcp_device_sm_count(void);

/* ---- box mask-mix: out = a*M + b*(1-M), M = 0 inside box [b, b+p) else 1.  Bit-exact vs the tensor
 * expression at LA_BCP_train.py:155,248-249; ACDC_BCP_train.py:244,372-373; pancreas/train_pancreas.py:86,155-156.
 * a, b, out: fp32 [n][c][X][Y][Z].  box6_dev = {x0,y0,z0,px,py,pz} int32 in DEVICE memory (so a captured CUDA graph can
 * be replayed with a fresh box); it is clipped to the volume like Python slicing.  Same convention for label_mix and
 * mix_loss (box6). */
int bcp_mask_mix(const float* a, const float* b, float* out, int n, int c, int X, int Y, int Z,
                 const int* box6_dev, cudaStream_t stream);

/* uint8 label maps mixed with the same box (label_batch at LA_BCP_train.py:156, ACDC_BCP_train.py:245). */
int bcp_label_mix(const unsigned char* a, const unsigned char* b, unsigned char* out, int n, int X, int Y, int Z,
                  const int* box6_dev, cudaStream_t stream);

/* ---- pseudo labels from planar fp32 logits [n][c][v] -> uint8 [n][v].
 * mode 0: (softmax(x,1) >= thr)[:,1]     LA_BCP_train.py:57-60, pancreas/pancreas_utils.py:275-278   (c == 2)
 * mode 1: argmax_c softmax(x,1)          ACDC_BCP_train.py:112-114                                   (c == 2|4) */
int bcp_pseudo_label(const float* logits, unsigned char* out, int n, int c, long long v, int mode, float thr,
                     cudaStream_t stream);

/* ---- largest connected component per (sample, class 1..3); LA_BCP_train.py:65-77, pancreas_utils.py:284-296,
 * ACDC_BCP_train.py:89-109 (skimage.measure.label + bincount on the CPU in the reference).
 * seg: uint8 [n][X][Y][Z] with values 0..3; outputs (either may be NULL): uint8 and/or float32 of the same shape
 * holding class value where the voxel belongs to the largest component of its class, else 0. */
long long bcp_largest_cc_workspace_bytes(int n, long long v);
int bcp_largest_cc(const unsigned char* seg, unsigned char* out_u8, float* out_f32, void* workspace, int n, int X, int Y,
                   int Z, int connectivity, cudaStream_t stream);

/* ---- fused mask-weighted Dice + CE.  form 0: utils/BCP_utils.py:58-69 + utils/losses.py:47-77 (LA) and
 * pancreas/losses.py:82-141; form 1: ACDC_BCP_train.py:167-179 + utils/losses.py:102-134.
 * Pre-train losses (LA_BCP_train.py:159-161) are the empty-box case.  `mask` (optional uint8 [n][v], 1 = image
 * region) overrides the box for callers that hold an explicit loss mask (utils/losses.py:47 mask_DiceLoss API).
 * ctx (device, bcp_mix_loss_ctx_floats): [0]=(dice+ce)/2, [1]=dice, [2]=ce, rest = backward coefficients.
 * bwd: grad3 (device) = upstream gradients of ctx[0..2]; dlogits = (g1+g0/2)*d(dice)/dx + (g2+g0/2)*d(ce)/dx. */
long long bcp_mix_loss_ctx_floats(int n, int c);
long long bcp_mix_loss_workspace_floats(int n, int c, long long v);
int bcp_mix_loss_fwd(const float* logits, const unsigned char* lab_img, const unsigned char* lab_patch,
                     const unsigned char* mask, float* ctx, float* workspace, int n, int c, int X, int Y, int Z, const int* box6, int form, float w_img,
                     float w_patch, cudaStream_t stream);
int bcp_mix_loss_bwd(const float* logits, const unsigned char* lab_img, const unsigned char* lab_patch,
                     const unsigned char* mask, const float* ctx, const float* grad3, float* dlogits,
                     int n, int c, int X, int Y, int Z, const int* box6, cudaStream_t stream);

/* ---- optimiser + EMA over flat fp32 arenas.
 * SGD: torch.optim.SGD(momentum, weight_decay) at LA_BCP_train.py:218 / ACDC_BCP_train.py:334 fused with
 *      update_ema_variables (utils/BCP_utils.py:78-81) / update_model_ema (ACDC_BCP_train.py:123-129).
 *      hyper (device) = {lr, momentum, weight_decay, ema_alpha, grad_scale, 1-ema_alpha}
 * Adam: torch.optim.Adam(lr) at pancreas/dataloaders.py:182 + pancreas/pancreas_utils.py:299-302.
 *      hyper (device) = {lr, beta1, beta2, eps, ema_alpha, grad_scale, 1-beta1^t, sqrt(1-beta2^t), 1-ema_alpha,
 *                        lr/(1-beta1^t), 1-beta1, 1-beta2}  (12 floats)
 * elements [n_train, n_total) are EMA-only; ema may be NULL. */
int bcp_sgd_ema_step(float* params, const float* grads, float* momentum, float* ema, const float* hyper,
                     long long n_train, long long n_total, cudaStream_t stream);
int bcp_adam_ema_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema,
                      const float* hyper, long long n_train, long long n_total, cudaStream_t stream);
/* advances the DEVICE step counter and writes the bias corrections {hyper[6], hyper[7]} and step_size hyper[9] = lr/bc1
 * (double precision, like torch.optim.Adam's Python scalars); enqueue once before each bcp_adam_ema_step. */
int bcp_adam_tick(float* hyper, long long* step, cudaStream_t stream);
int bcp_ema_i64(long long* ema, const long long* model, int n, float alpha, float one_minus_alpha, cudaStream_t stream);

/* ---- weight repack: fp32 master weights -> bf16 operand layouts (one launch per network).
 * source tensor is [dim_a][dim_b][taps] fp32 at arena + src_off (floats); destination at packed + dst_off (bf16 elems)
 * kind 0: [T][dim_b/8][dim_a][8] (inner = b)          conv fwd / stride-2 gather operand
 * kind 1: [T][dim_a/8][dim_b][8] (inner = a), taps reversed     conv dgrad operand
 * kind 2: [T][dim_a/8][dim_b][8] (inner = a)          stride-2 scatter operand (CUDA-core kernel)
 * kind 3: [dim_a/8][T][dim_b][8] (inner = a)          stride-2 scatter operand (tcgen05 kernel) */
typedef struct bcp_repack_job {
  long long src_off;
  long long dst_off;
  int dim_a, dim_b, taps, kind;
} bcp_repack_job;
int bcp_weights_repack(const float* arena, void* packed, const bcp_repack_job* jobs_dev, int njobs, cudaStream_t stream);

/* ---- layout converts planar fp32 <-> CB8 bf16 */
int bcp_planar_to_cb8(const float* in, void* out, int n, int c, long long s, cudaStream_t stream);
int bcp_cb8_to_planar(const void* in, float* out, int n, int c, long long s, cudaStream_t stream);

/* ---- normalisation (train-mode nn.BatchNorm3d/2d: networks/VNet.py:19, networks/unet.py:21,25;
 * nn.InstanceNorm3d: pancreas/Vnet.py:25,49,76) fused with ReLU/LeakyReLU, Dropout3d/Dropout and the skip add.
 * groups of `spg` consecutive samples share statistics (BatchNorm of one reference forward call = one group;
 * InstanceNorm: spg = 1).  stat/coef: fp32 [n/spg][c][2] = {mean, invstd} / {scale, shift}.
 * `counter`: one device int, zero before the first call; the last block of the reduction grid finalises in fixed order
 * and resets it (deterministic; lets one launch replace partial + finalize kernels). */
int bcp_norm_chunks(int n, int c, long long s);
long long bcp_norm_workspace_floats(int n, int c, long long s);
int bcp_norm_stats(const void* y, const float* gamma, const float* beta, float* running_mean, float* running_var,
                   long long* num_batches_tracked, float* stat, float* coef, float* workspace, int* counter,
                   int n, int c, long long s, int spg, float eps, float momentum, cudaStream_t stream);
int bcp_norm_eval_coef(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float* stat, float* coef, int c, int groups, float eps, cudaStream_t stream);
int bcp_norm_apply(const void* y, void* out, const float* coef, const float* chan_scale, const unsigned char* elem_keep,
                   float elem_scale, const void* residual, int n, int c, long long s, int spg, float slope,
                   cudaStream_t stream);
int bcp_norm_bwd(const void* dact, const void* y, void* dy, const float* stat, const float* coef, const float* chan_scale,
                 const unsigned char* elem_keep, float elem_scale, float* dgamma, float* dbeta, float* sums,
                 float* workspace, int* counter, int n, int c, long long s, int spg, float slope, int stats_grad,
                 int accumulate, cudaStream_t stream);

/* Single-launch variants for layers of at most 64 Ki voxels per statistics group (csrc/norm_fused.cu): one thread-block
 * cluster per (group, channel octet) reduces through distributed shared memory, so statistics + normalise/activation
 * (forward) and both reductions + the input gradient (backward) are ONE kernel each.  Same semantics and argument meaning
 * as bcp_norm_stats + bcp_norm_apply / bcp_norm_bwd; `workspace` >= (n/spg)*c*2 floats; `counter`: one zero device int of
 * its own (self-resetting).  bcp_norm_fused_supported tells whether a shape is eligible. */
int bcp_norm_fused_supported(int n, int c, long long s, int spg);
int bcp_norm_fused_fwd(const void* y, void* out, const float* gamma, const float* beta, float* running_mean, float* running_var,
                       long long* num_batches_tracked, float* stat, float* coef, float* workspace, int* counter,
                       const float* chan_scale, const unsigned char* elem_keep, float elem_scale, const void* residual,
                       int n, int c, long long s, int spg, float eps, float momentum, float slope, cudaStream_t stream);
int bcp_norm_fused_bwd(const void* dact, const void* y, void* dy, const float* stat, const float* coef, const float* chan_scale,
                       const unsigned char* elem_keep, float elem_scale, float* dgamma, float* dbeta, float* sums, int* counter,
                       int n, int c, long long s, int spg, float slope, int stats_grad, int accumulate, cudaStream_t stream);

/* ---- convolutions (nn.Conv3d / nn.Conv2d / nn.ConvTranspose3d call sites: networks/VNet.py:17,74,101,210;
 * networks/unet.py:20,24,49,102; pancreas/Vnet.py:19,43,70,128).  dims/kernel/stride/pad are int[3] (x,y,z).
 * conv_tc_*: tcgen05 + TMA implicit GEMM for 3x3x3 / 1x3x3, stride 1, 'same' padding, channels % 16 == 0.
 * conv_direct_*: CUDA-core kernels for everything else (see csrc/conv_direct.cu).
 * Every weight-gradient entry point takes `accumulate`: 0 = overwrite dw, 1 = dw += (lets the host point dw straight
 * into the flat gradient arena instead of adding a temporary). */
int bcp_conv_direct_fwd(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                        const int* in_dims, const int* kernel, const int* stride, const int* pad, int transposed,
                        cudaStream_t stream);
long long bcp_conv_wgrad_workspace_floats(int n, int cin, int cout, const int* out_dims, const int* kernel);
int bcp_conv_direct_wgrad(const void* in, const void* outgrad, float* dw, float* workspace, int n, int cin, int cout,
                          const int* in_dims, const int* kernel, const int* stride, const int* pad, int accumulate,
                          cudaStream_t stream);
long long bcp_chan_sum_workspace_floats(int n, int c, long long s);
int bcp_chan_sum(const void* x, float* out, float* workspace, int n, int c, long long s, cudaStream_t stream);
int bcp_conv_first_fwd(const float* in, const float* w, const float* bias, void* out, int n, int cout,
                       const int* dims, const int* kernel, cudaStream_t stream);
long long bcp_conv_first_wgrad_workspace_floats(int n, int cout, const int* dims, const int* kernel);
/* > 0 (the number of per-CTA partials) when bcp_conv_first_wgrad takes the TMA-staged kernel (conv_first_tma.cu) for this
 * shape, 0 when it uses the register-window kernel (Z % 4 != 0, more than two channel octets, no tensor-map encoder). */
int bcp_conv_first_wgrad_tma_chunks(int n, int cout, const int* dims, const int* kernel);
int bcp_conv_first_wgrad(const float* in, const void* outgrad, float* dw, float* workspace, int n, int cout,
                         const int* dims, const int* kernel, int accumulate, cudaStream_t stream);
int bcp_head_fwd(const void* in, const float* w, const float* bias, float* logits, int n, int cin, int ncls,
                 const int* dims, const int* kernel, cudaStream_t stream);
int bcp_head_dgrad(const float* dlogits, const float* w, void* din, int n, int cin, int ncls, const int* dims,
                   const int* kernel, cudaStream_t stream);
long long bcp_head_wgrad_workspace_floats(int n, int cin, int ncls, const int* dims, const int* kernel);
int bcp_head_wgrad(const void* in, const float* dlogits, float* dw, float* db, float* workspace, int n, int cin, int ncls,
                   const int* dims, const int* kernel, int accumulate, cudaStream_t stream);

/* tcgen05 implicit-GEMM convolution, stride 1, 'same' zero padding, kernel (kx,3,3) with kx in {1,3}.
 * wpack is the kind-0 (forward) or kind-1 (dgrad) pack.  Returns -2 for shapes it does not take. */
int bcp_conv_tc_supported(int cin, int cout, const int* dims, const int* kernel);
/* chosen tiling for inspection: {BX,BY,BZ,MT,SA,SB,AS,nbricks,tmem_cols,smem_bytes} */
int bcp_conv_tc_plan(int n, int cin, int cout, const int* dims, const int* kernel, int* plan10);
int bcp_conv_tc_fwd(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                    const int* dims, const int* kernel, cudaStream_t stream);
/* The same convolution with the train-mode normalisation statistics of its OUTPUT fused into the epilogue (replaces a
 * following bcp_norm_stats call: identical stat/coef/running-statistic semantics, statistics taken over the bf16-rounded
 * tensor that is stored).  bcp_conv_tc_stats_workspace_bytes returns 0 when the layer is not eligible (then call
 * bcp_conv_tc_fwd + bcp_norm_stats); `counter` as for bcp_norm_stats. */
long long bcp_conv_tc_stats_workspace_bytes(int n, int cin, int cout, const int* dims, const int* kernel, int spg);
int bcp_conv_tc_fwd_stats(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                          const int* dims, const int* kernel, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, long long* num_batches_tracked, float* stat, float* coef, void* workspace,
                          int* counter, int spg, float eps, float momentum, cudaStream_t stream);
/* Sliding-window validation, reference utils/test_3d_patch.py:82-141 (parity: tests/test_gpu_networks.py::test_sliding_window_validation).
 * bcp_window_accumulate adds softmax(logits)[cls] of ONE window (logits planar fp32 [c][patch]) into score/count
 * ([vol] fp32) at `origin3`; call it window by window in the reference's x,y,z order.  bcp_window_finalize turns score into
 * the mean probability in place and writes label = (mean > threshold). */
int bcp_window_accumulate(const float* logits, float* score, float* count, int c, int cls, const int* patch3, const int* vol3,
                          const int* origin3, cudaStream_t stream);
int bcp_window_finalize(float* score, const float* count, unsigned char* label, long long n, float threshold, cudaStream_t stream);

/* The same convolution for Cout in {16, 32} with the three dz taps folded into the MMA N dimension;
 * the default forward / data-gradient kernel of those layers.  Same arguments and packs as bcp_conv_tc_fwd. */
int bcp_conv_tc_fold_supported(int cin, int cout, const int* dims, const int* kernel);
int bcp_conv_tc_fold_plan(int n, int cin, int cout, const int* dims, const int* kernel, int* plan10);
int bcp_conv_tc_fold_fwd(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                         const int* dims, const int* kernel, cudaStream_t stream);
/* debug only (tools/debug_conv_tc.py): the instrumented template instance of bcp_conv_tc_fwd (fold = 0) or
 * bcp_conv_tc_fold_fwd (fold = 8, 12 or 16 = number of epilogue warps); `prof` (device, 16 x uint64 per CTA, >= #SMs CTAs) receives per-role wait cycles. */
int bcp_conv_tc_fwd_profiled(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                             const int* dims, const int* kernel, int fold, void* prof, cudaStream_t stream);

/* tcgen05 weight gradient for the same family: dw[cout][cin][taps] fp32 (PyTorch layout), deterministic.  ONE launch
 * (grid <= #SMs, one CTA per SM, so all CTAs are co-resident): every CTA accumulates its share of the voxels in TMEM, writes a partial to `workspace`, meets the others at a
 * device-wide barrier and then all CTAs reduce the partials in fixed order into dw.  `counter`: TWO device ints, zero
 * before the first call (the kernel resets them), private to the stream. */
int bcp_conv_tc_wgrad_supported(int cin, int cout, const int* dims, const int* kernel);
long long bcp_conv_tc_wgrad_workspace_floats(int n, int cin, int cout, const int* dims, const int* kernel);
int bcp_conv_tc_wgrad(const void* a, const void* dy, float* dw, float* workspace, int* counter, int n, int cin, int cout,
                      const int* dims, const int* kernel, int accumulate, cudaStream_t stream);
/* plan exploration (tools/debug_conv_tc.py): the same launch with the split-K factor capped at max_splits (0 = planner's choice) */
int bcp_conv_tc_wgrad_capped(const void* a, const void* dy, float* dw, float* workspace, int* counter, int n, int cin, int cout,
                      const int* dims, const int* kernel, int accumulate, int max_splits,
                             cudaStream_t stream);

/* tcgen05 stride-2 family (nn.Conv3d(k=2,s=2) networks/VNet.py:74, nn.ConvTranspose3d(k=2,s=2) networks/VNet.py:101).
 * half_dims = half-resolution grid.  mode 1 gather: in = full-res [cin], wpack kind 0, out = half-res [cout].
 * mode 2 scatter: in = half-res [cin], wpack kind 3 ([cin/8][8][cout][8]), out = full-res [cout].
 * wgrad: dw[c_half][c_full][8] = sum_i half[i][c_half] * full[2i+t][c_full]. */
int bcp_conv_tc_s2_supported(int cin, int cout, const int* half_dims, int mode);
int bcp_conv_tc_s2_fwd(const void* in, const void* wpack, const float* bias, void* out, int n, int cin, int cout,
                       const int* half_dims, int mode, cudaStream_t stream);
int bcp_conv_tc_s2_wgrad_supported(int c_half, int c_full, const int* half_dims);
long long bcp_conv_tc_s2_wgrad_workspace_floats(int n, int c_half, int c_full, const int* half_dims);
int bcp_conv_tc_s2_wgrad(const void* full, const void* half, float* dw, float* workspace, int* counter, int n, int c_half,
                         int c_full, const int* half_dims, int accumulate, cudaStream_t stream);

/* ---- DiceLoss on probabilities (utils/losses.py:113-134, called with softmax=False from ACDC_BCP_train.py:170,175):
 * probs fp32 [n][c][v] (already soft-maxed by the caller), target uint8 [n][v], mask uint8 [n][v] or NULL.
 * ctx[0] = loss (mean over classes of 1 - batch-global Dice, smooth 1e-10); bwd writes dL/dprobs * grad_out[0]. */
long long bcp_dice_prob_ctx_floats(int c);
long long bcp_dice_prob_workspace_floats(int n, int c, long long v);
int bcp_dice_prob_fwd(const float* probs, const unsigned char* target, const unsigned char* mask, float* ctx, float* workspace,
                      int n, int c, long long v, cudaStream_t stream);
int bcp_dice_prob_bwd(const float* probs, const unsigned char* target, const unsigned char* mask, const float* ctx,
                      const float* grad_out, float* dprobs, int n, int c, long long v, cudaStream_t stream);

/* ---- device-side input pipeline.  One sample: RandomRotFlip (np.rot90 by k in the (x,y) plane, then
 * np.flip along flip_axis), RandomCrop's zero padding `pad` per side and crop at `origin` (in the padded frame), ToTensor
 * (dataloaders/dataset.py:52-60,173-225,267-277).  img fp32 / lab uint8 [W][H][D] = src_dims resident in HBM; outputs
 * [OX][OY][OZ] = out_dims (one sample of the batch tensor).  Bit-exact vs the numpy transforms. */
int bcp_aug_crop_rotflip(const float* img, const unsigned char* lab, float* out_img, unsigned char* out_lab, const int* src_dims,
                         const int* out_dims, int k, int flip_axis, const int* pad, const int* origin, cudaStream_t stream);

/* ---- resampling (networks/unet.py:37 MaxPool2d(2); :50 Upsample(bilinear, align_corners=True);
 * networks/VNet.py:249 MaxPool3d(3, stride=2)).  planes = n * ceil(c/8) * X. */
int bcp_maxpool2_fwd(const void* in, void* out, long long planes, int y, int z, cudaStream_t stream);
int bcp_maxpool2_bwd(const void* in, const void* dout, void* din, long long planes, int y, int z, cudaStream_t stream);
int bcp_upsample2_fwd(const void* in, void* out, long long planes, int y, int z, cudaStream_t stream);
int bcp_upsample2_bwd(const void* dout, void* din, long long planes, int y, int z, cudaStream_t stream);
int bcp_maxpool3d_k3s2(const void* in, float* out, int n, int c, int x, int y, int z, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BCP_B200_H_ */
