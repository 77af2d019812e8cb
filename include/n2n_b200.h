/*
 * n2n_b200.h — C-ABI of libn2n_b200.so: the B200 (sm_100a) kernels behind the
 * Neighbor2Neighbor hot path of lmh9507/image_denoising.
 *
 * The reference is 100 % Python on stock PyTorch and has no FFI / plugin layer
 * (SURVEY.md §0.1, §8b); its boundary is the Python call surface.  Each entry
 * point below therefore cites the reference *Python* interface whose arithmetic
 * it replaces (file:line into the reference repo).  The ctypes binding a
 * maintainer would add is shown in INTEGRATION.md and implemented in
 * image_denoising_b200/_ext.py.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; n2n_last_error() gives a
 *     thread-local message;
 *   - the caller owns every buffer (device pointers, normally torch tensors);
 *     the library never allocates or frees device memory: scratch comes from a
 *     caller-provided workspace sized by the matching *_workspace_bytes query;
 *   - every call is asynchronous on the passed stream (a cudaStream_t passed as
 *     void*); no call synchronises the device;
 *   - "NCHW" tensors are contiguous; dtype tags: N2N_F32 = 0, N2N_BF16 = 1
 *     (for activations/compute), element sizes for the copy-only sub-sampler;
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with an error code.
 */
#ifndef N2N_B200_H
#define N2N_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define N2N_F32 0
#define N2N_BF16 1

#define N2N_OK 0
#define N2N_ERR_ARG (-1)
#define N2N_ERR_CUDA (-2)
#define N2N_ERR_UNSUPPORTED (-3)

const char* n2n_last_error(void);
int n2n_version(void);
/* 1 if a CUDA device with compute capability 10.x is present. */
int n2n_device_ok(void);
/* kernels launched so far by the calling host thread (bench.py's gpu_launches). */
long long n2n_launch_count(void);
/* Per-launch CUDA-event timing of the GEMM kernel classes on the launching stream (bench.py's
 * roofline leg): begin() arms it for the calling thread; end() stops it, waits for the recorded
 * events and fills out[6] = {ms, executed FLOPs, launches} for the tap-GEMM kernel (conv / deconv
 * forward + input gradient) followed by the same three for the weight-gradient kernel. */
int n2n_profile_begin(void);
int n2n_profile_active(void);   /* 1 between begin() and end() on the calling thread */
int n2n_profile_end(double* out);
/* Per-launch variant: fills rows of {class, ms, executed FLOPs} in launch order, returns the row count. */
int n2n_profile_end_list(double* out, int max_rows);

/* ------------------------------------------------------------------------- *
 * Neighbour sub-sampler — train.py:141-190 (generate_mask_pair,
 * generate_subimages), training_script.md:137-144.
 * ------------------------------------------------------------------------- */

/* train.py:151-172: rd_idx (int64, one value in [0,8) per 2x2 cell, cells in
 * (n,i,j) order) -> two flat bool masks of 4*cells bytes each (one 1 per cell)
 * and/or a packed selector (k1 | k2<<2, one byte per cell).  Any output may be
 * NULL.  The random draw itself stays with the caller's torch generator. */
int n2n_mask_pair_from_rdidx(const int64_t* rd_idx, int64_t cells,
                             uint8_t* mask1, uint8_t* mask2, uint8_t* packed_sel,
                             void* stream);

/* train.py:175-190: out[n,c,i,j] = img[n,c,2i+k/2,2j+k%2], k = position of the 1
 * in cell (n,i,j) of `mask` (flat bool, 4 bytes per cell).  elem_size in
 * {1,2,4,8}: a pure copy, bit-exact in any dtype.  img is NCHW [n,c,h,w]
 * contiguous, out is [n,c,h/2,w/2]. */
int n2n_subsample(const void* img, const uint8_t* mask, void* out,
                  int n, int c, int h, int w, int elem_size, void* stream);

/* Fused pair (the form the N2N step uses): one pass over img producing both
 * sub-images.  Selector source: either the two reference bool masks
 * (mask1/mask2 non-NULL) or the packed selector (packed_sel non-NULL). */
int n2n_subsample_pair(const void* img, const uint8_t* mask1, const uint8_t* mask2,
                       const uint8_t* packed_sel, void* out1, void* out2,
                       int n, int c, int h, int w, int elem_size, void* stream);

/* train.py:134-138 (space_to_depth): F.unfold(x, bs, stride=bs) viewed as [n, c*bs*bs, h/bs, w/bs]:
 * y[n, c*bs*bs + ky*bs + kx, i, j] = x[n, c, i*bs+ky, j*bs+kx].  Pure copy, elem_size in {1,2,4,8}.
 * (generate_subimages never materialises it; kept because it is a public name of the reference.) */
int n2n_space_to_depth(const void* x, void* y, int n, int c, int h, int w, int block_size,
                       int elem_size, void* stream);

/* ------------------------------------------------------------------------- *
 * Single layers (NCHW fp32 at the boundary; compute in `dtype`) —
 * arch_unet.py:113-190 (nn.Conv2d 3x3 s1 p1 / 1x1, LeakyReLU(0.2) in place),
 * arch_unet.py:57-62 (ConvTranspose2d k2 s2), arch_unet.py:120-136 (MaxPool2d(2)).
 * These are the per-op parity surface; the network-level entry points below run
 * the same kernels without leaving the blocked device layout.
 * ------------------------------------------------------------------------- */
size_t n2n_conv2d_workspace_bytes(int n, int cin, int cout, int h, int w, int ksize, int dtype);

/* y = act(conv2d(x, w, b)); ksize in {1,3}, stride 1, pad ksize/2.
 * w: [cout,cin,k,k] fp32, b: [cout] or NULL; act_slope < 0 => no activation,
 * otherwise y = v>0 ? v : act_slope*v (0.2 LeakyReLU, 0 ReLU). */
int n2n_conv2d_fwd(const float* x, const float* w, const float* b, float* y,
                   int n, int cin, int cout, int h, int w_, int ksize, float act_slope,
                   int dtype, void* workspace, void* stream);
/* dx = conv2d_input_grad(dy, w) (no activation handling). */
int n2n_conv2d_dgrad(const float* dy, const float* w, float* dx,
                     int n, int cin, int cout, int h, int w_, int ksize,
                     int dtype, void* workspace, void* stream);
/* dw = conv2d_weight_grad(x, dy) [cout,cin,k,k], db = sum(dy) [cout] (db may be NULL). */
int n2n_conv2d_wgrad(const float* x, const float* dy, float* dw, float* db,
                     int n, int cin, int cout, int h, int w_, int ksize,
                     int dtype, void* workspace, void* stream);

size_t n2n_deconv2x2_workspace_bytes(int n, int cin, int cout, int h, int w, int dtype);
/* y[n,co,2i+a,2j+b] = b[co] + sum_ci x[n,ci,i,j] * w[ci,co,a,b]; w: [cin,cout,2,2]. */
int n2n_deconv2x2_fwd(const float* x, const float* w, const float* b, float* y,
                      int n, int cin, int cout, int h, int w_, int dtype, void* workspace, void* stream);
int n2n_deconv2x2_dgrad(const float* dy, const float* w, float* dx,
                        int n, int cin, int cout, int h, int w_, int dtype, void* workspace, void* stream);
int n2n_deconv2x2_wgrad(const float* x, const float* dy, float* dw, float* db,
                        int n, int cin, int cout, int h, int w_, int dtype, void* workspace, void* stream);

size_t n2n_pool_workspace_bytes(int n, int c, int h, int w, int dtype);
/* y = maxpool2x2(x);  x: [n,c,h,w] -> y: [n,c,h/2,w/2]. */
int n2n_maxpool2_fwd(const float* x, float* y, int n, int c, int h, int w, int dtype,
                     void* workspace, void* stream);
/* dx = lrelu'(x; slope) * unpool(dy) with ATen's first-max tie rule; x is the
 * (already activated) pool input.  slope = 1 gives the plain max-pool backward. */
int n2n_maxpool2_bwd(const float* x, const float* dy, float* dx, int n, int c, int h, int w,
                     float slope, int dtype, void* workspace, void* stream);

/* ------------------------------------------------------------------------- *
 * UNet — arch_unet.py:100-260 (non-blindspot).  A plan is bound to
 * (in_nc,out_nc,n_feature,N,H,W,dtype); H and W must be multiples of 32.
 * params/grads: the 50 tensors in state_dict order (arch_unet.py:114-192),
 * fp32, PyTorch layouts (Conv2d [co,ci,k,k], ConvTranspose2d [ci,co,2,2]).
 * ------------------------------------------------------------------------- */
typedef struct n2n_unet_plan n2n_unet_plan;
#define N2N_UNET_NUM_PARAMS 50

int n2n_unet_plan_create(n2n_unet_plan** plan, int in_nc, int out_nc, int n_feature,
                         int n, int h, int w, int dtype, int with_backward);
/* arch_unet.RESNET (arch_unet.py:263-409): the same convolutions at full resolution, no pooling / up-sampling, global
 * residual out = net(x) + x (:409).  Returns the same plan type: n2n_unet_forward / _backward / _workspace_bytes /
 * _plan_destroy apply, with params / grads = the 42 tensors in state_dict order (:279-347; up5.deconv.* are registered by
 * the reference but unused: ignored here, their gradient slots are not written).  Needs out_nc == in_nc. */
#define N2N_RESNET_NUM_PARAMS 42
int n2n_resnet_plan_create(n2n_unet_plan** plan, int in_nc, int out_nc, int n_feature,
                           int n, int h, int w, int dtype, int with_backward);
void n2n_unet_plan_destroy(n2n_unet_plan* plan);
size_t n2n_unet_workspace_bytes(const n2n_unet_plan* plan);
/* Diagnostic / layer-level parity tests: copy activation buffer `buffer` of the last forward on `workspace` out as fp32
 * NCHW [N][16*blocks][H_l][W_l] (padding channels included).  buffer: 0..4 = concat buffers of levels 0..4 ([up | skip]),
 * 5..10 = enc_conv0..5 outputs, 11 = pool5, 12 = enc_conv6, 13.. = dec_conv5a, 5b, 4a, 4b, 3a, 3b, 2a, 2b, 1a, 1b,
 * 23 = nin_a, 24 = nin_b.  dims (may be NULL) = {channels, H_l, W_l}; out may be NULL (size query).  Returns the element
 * count (0: the buffer is not used by this plan), < 0 on error.  Buffers a fused launch skips hold stale data. */
long long n2n_unet_read_activation(const n2n_unet_plan* plan, void* workspace, int buffer, float* out, int* dims, void* stream);
/* number of kernel launches one forward / backward issues (for gpu_launches). */
int n2n_unet_launches(const n2n_unet_plan* plan, int backward);
/* y = UNet(x): x [N,in_nc,H,W] fp32, y [N,out_nc,H,W] fp32.  With a
 * with_backward plan the activations needed by n2n_unet_backward stay in the
 * workspace until the next forward on the same workspace. */
int n2n_unet_forward(n2n_unet_plan* plan, const float* const* params, const float* x, float* y,
                     void* workspace, void* stream);
/* One N2N step runs the network twice on the same weights (the no-grad denoise of the
 * full image and the training forward on the sub-image, training_script.md:139-146).
 * After this call `plan` reads its forward weights / padded biases from `donor`'s
 * workspace instead of repacking them: the caller guarantees that donor's
 * n2n_unet_forward with the same `params` precedes plan's on the same stream.
 * Sharing is per layer: a layer whose launch form differs between the two plans
 * (deepest levels of small inputs) is still packed locally.  Returns 0 when shared,
 * 1 when the plans are not the same network (nothing changed); donor = NULL ends it. */
int n2n_unet_share_weights(n2n_unet_plan* plan, const n2n_unet_plan* donor, const void* donor_workspace);
/* The pack step of n2n_unet_forward on its own (parameters -> packed weights / padded biases in `workspace`); the NEXT
 * n2n_unet_forward of this plan then skips it.  Lets a caller that runs two plans on two streams (the no-grad
 * full-resolution pass and the training forward of one N2N iteration, training_script.md:139-146, are independent) pack
 * once, fork, and still share the packed weights (n2n_unet_share_weights).  No-op for the RESNET plan. */
int n2n_unet_pack_weights(n2n_unet_plan* plan, const float* const* params, void* workspace, void* stream);
/* grads[i] (fp32, same shapes as params) are OVERWRITTEN with dL/dparam for the
 * last forward; dy is dL/dy [N,out_nc,H,W].  dx (may be NULL) receives dL/dx. */
int n2n_unet_backward(n2n_unet_plan* plan, const float* const* params, const float* dy,
                      float* const* grads, float* dx, void* workspace, void* stream);

/* ------------------------------------------------------------------------- *
 * Output adapter — adapter.py:5-26 (OutputAdapter.forward), :59-67.
 * out = base_out + conv3x3(relu(conv3x3(cat[noisy, base_out]))).
 * params/grads: net.0.weight [hid,2C,3,3], net.0.bias, net.2.weight [C,hid,3,3], net.2.bias.
 * C in {1, 3} with hid = 16 and W % 4 == 0 (the reference's configurations) runs as direct fp32
 * convolutions on the CUDA cores (csrc/adapter_fused.cu: weights in the constant bank, exact in both
 * precision modes); other shapes run as 16-channel-block tap GEMMs on the engine `dtype` selects.
 * The backward reads the hidden activations the forward of the SAME plan + workspace left behind and
 * takes the forward's inputs again (noisy, base_out: the concat operand of conv1's weight gradient).
 * ------------------------------------------------------------------------- */
typedef struct n2n_adapter_plan n2n_adapter_plan;
int n2n_adapter_plan_create(n2n_adapter_plan** plan, int channels, int hidden,
                            int n, int h, int w, int dtype, int with_backward);
void n2n_adapter_plan_destroy(n2n_adapter_plan* plan);
size_t n2n_adapter_workspace_bytes(const n2n_adapter_plan* plan);
int n2n_adapter_forward(n2n_adapter_plan* plan, const float* const* params,
                        const float* noisy, const float* base_out, float* out,
                        void* workspace, void* stream);
int n2n_adapter_backward(n2n_adapter_plan* plan, const float* const* params, const float* noisy,
                         const float* base_out, const float* dout, float* const* grads, void* workspace,
                         void* stream);

/* ------------------------------------------------------------------------- *
 * Losses
 * ------------------------------------------------------------------------- */
size_t n2n_loss_workspace_bytes(int64_t count);
/* training_script.md:146-153: diff = out-sub2, exp = den1-den2,
 * loss = mean(diff^2) + lam*mean((diff-exp)^2).  Writes loss3 = {loss_all, loss1,
 * loss2} (fp32, device) and, if grad != NULL, grad = grad_scale * dloss/dout. */
int n2n_loss_n2n_fwdbwd(const float* out, const float* sub2, const float* den1, const float* den2,
                        float lam, float grad_scale, int64_t count, float* loss3, float* grad,
                        void* workspace, void* stream);
/* finetune.py:153-162, :283-285: loss = L1(p,t) + lambda_grad*(L1(dx p,dx t)+L1(dy p,dy t)).
 * loss3 = {loss, loss_l1, loss_grad}; grad (may be NULL) = grad_scale * dloss/dpred. */
int n2n_loss_l1grad_fwdbwd(const float* pred, const float* target, int n, int c, int h, int w,
                           float lambda_grad, float grad_scale, float* loss3, float* grad,
                           void* workspace, void* stream);

/* util.py:41-70 (Structure_loss, the criterion of the fork's live loop train.py:322, :361-363):
 * loss = alpha*L1(pred, target) + beta*(L1(dy pred2) + L1(dx pred2))/2 + gamma*L1(pred2, target), pred = network(noisy),
 * pred2 = network(clean).  loss4 = {loss, pixel, TV, consistency}; grad_pred / grad_pred2 (may be NULL) receive
 * grad_scale * dloss/dpred and dloss/dpred2. */
int n2n_loss_structure_fwdbwd(const float* pred, const float* pred2, const float* target, int n, int c, int h, int w,
                              float alpha, float beta, float gamma, float grad_scale, float* loss4,
                              float* grad_pred, float* grad_pred2, void* workspace, void* stream);

/* finetune_iqsl.py:291-383 (iqsl_loss): 3-class intensity-quantised Dice + ce_factor * soft CE on
 * single-channel pred / target in [0,1] (count = all elements).  workspace: n2n_loss_iqsl_workspace_bytes() bytes, zero on
 * first use.  loss3 = {total, dice term, CE term}; grad (may be NULL) = grad_scale * dloss/dpred. */
size_t n2n_loss_iqsl_workspace_bytes(void);
int n2n_loss_iqsl_fwdbwd(const float* pred, const float* target, int64_t count, float t1, float t2, float tau, float margin,
                         float ce_factor, float eps, float grad_scale, float* loss3, float* grad, void* workspace, void* stream);

/* ------------------------------------------------------------------------- *
 * ImprovedUNet executor — arch_unet.py:475-531 (ImprovedUNet.forward with its RDB / ResBlock / UpBlock
 * sub-modules, :420-472).  One call runs the whole network with the activations resident in the engines' blocked
 * layout (dense concats and skip concats are block ranges written in place).  params: the module's parameters in
 * state_dict order (n2n_improved_num_params() of them; noise_estimator.*, downs.*, bottle.*, ups.*, final.*).
 * x: [n, in_nc, h, w] fp32, y: [n, out_nc, h, w] fp32; h, w multiples of 2^depth.  Workspace: caller-owned,
 * n2n_improved_workspace_bytes(plan) bytes.
 * ------------------------------------------------------------------------- */
typedef struct n2n_improved_plan n2n_improved_plan;
int n2n_improved_plan_create(n2n_improved_plan** plan, int in_nc, int out_nc, int n_feature, int depth, int noise,
                             int n, int h, int w, int dtype, int with_backward);
void n2n_improved_plan_destroy(n2n_improved_plan* plan);
size_t n2n_improved_workspace_bytes(const n2n_improved_plan* plan);
int n2n_improved_num_params(const n2n_improved_plan* plan);
int n2n_improved_launches(const n2n_improved_plan* plan, int backward);
int n2n_improved_forward(n2n_improved_plan* plan, const float* const* params, const float* x, float* y,
                         void* workspace, void* stream);
/* Backward of the last n2n_improved_forward of a plan created with_backward (which keeps every intermediate tensor and the
 * GroupNorm statistics in `workspace`): dy = dL/dy [n, out_nc, h, w], y = the forward's output (sigmoid'), grads[i] (fp32,
 * shapes of params[i]) are OVERWRITTEN with dL/dparam.  The operators are walked in reverse; every gradient tensor is
 * accumulated into a zero-initialised mirror of the activation buffers (dense concats collect the contributions of all
 * their consumers), LeakyReLU' is applied once per tensor after its last consumer, weight gradients run as engine-sized
 * channel chunks through one partial buffer (fixed-order reduction). */
/* Layer-level parity hook (tests): buffer `buf` of the plan's workspace (or, grad != 0, its gradient mirror after a
 * backward) as fp32 NCHW over all of its 16-channel blocks; out == NULL only returns the element count and dims. */
long long n2n_improved_read_buffer(const n2n_improved_plan* plan, const void* workspace, int buf, int grad, float* out,
                                   int* dims, void* stream);
int n2n_improved_backward(n2n_improved_plan* plan, const float* const* params, const float* dy, const float* y,
                          float* const* grads, void* workspace, void* stream);

/* ------------------------------------------------------------------------- *
 * Non-GEMM operators of arch_unet.ImprovedUNet — arch_unet.py:420-531, fp32 NCHW.
 * GroupNorm = norm2d('gn', c, 32) (arch_unet.py:7-15): n2n_groupnorm_groups() applies the reference's
 * "largest divisor of c that is <= groups" rule.  y = GN(x) * gamma + beta, then LeakyReLU(act_slope) when
 * act_slope >= 0, or + residual when given (the two forms ResBlock uses, :421-432; exclusive).  mean_rstd
 * [n][groups][2] (may be NULL in no-grad passes) is what the backward needs; biased variance, eps inside the
 * square root (torch.nn.GroupNorm).  workspace: n2n_groupnorm_workspace_bytes(n, c).  Backward: dy is the
 * gradient w.r.t. the (activated) output, y the forward output (only read when act_slope >= 0).
 * ------------------------------------------------------------------------- */
int n2n_groupnorm_groups(int channels, int groups);
size_t n2n_groupnorm_workspace_bytes(int n, int c);
int n2n_groupnorm_fwd(const float* x, const float* gamma, const float* beta, const float* residual, float* y,
                      float* mean_rstd, int n, int c, int hw, int groups, float eps, float act_slope,
                      void* workspace, void* stream);
int n2n_groupnorm_bwd(const float* x, const float* gamma, const float* y, const float* dy, const float* mean_rstd,
                      float* dx, float* dgamma, float* dbeta, int n, int c, int hw, int groups, float act_slope,
                      void* workspace, void* stream);
/* kind 1: LeakyReLU(slope) (arch_unet.py:424, :441, :464), kind 2: Sigmoid (:486, :530).  The backward reads the
 * OUTPUT (the reference's activations are in place): dx = dy * (y > 0 ? 1 : slope) or dy * y * (1 - y). */
int n2n_act_fwd(const float* x, float* y, int64_t count, int kind, float slope, void* stream);
int n2n_act_bwd(const float* y, const float* dy, float* dx, int64_t count, int kind, float slope, void* stream);
/* out = a + b — the residual connections of RDB / ResBlock (arch_unet.py:432, :449). */
int n2n_add_f32(const float* a, const float* b, float* out, int64_t count, void* stream);
/* nn.PixelShuffle(2) (arch_unet.py:456, :461): src [n][4*c_out][h][w] -> dst [n][c_out][2h][2w] with
 * dst[n,c,2y+i,2x+j] = src[n,4c+2i+j,y,x]; inverse != 0 runs the same map backwards (its gradient). */
int n2n_pixel_shuffle2(const float* src, float* dst, int n, int c_out, int h, int w, int inverse, void* stream);

/* ------------------------------------------------------------------------- *
 * Adam — train.py:332 (torch.optim.Adam defaults), finetune.py:260-263.
 * One launch over a table of tensors.  table (device, int64) holds, per tensor t
 * of ntensors: [p_ptr, g_ptr, m_ptr, v_ptr, numel]; blocks (device, int32) holds
 * per CUDA block: [tensor index, chunk index] (chunk = 2048 elements).
 * g is multiplied by grad_scale before use (1/world for data parallel).
 * ------------------------------------------------------------------------- */
#define N2N_ADAM_CHUNK 2048
int n2n_adam_multi(const int64_t* table, int ntensors, const int32_t* blocks, int nblocks,
                   float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                   void* stream);

/* CUDA-graph forms of the two per-step kernels whose scalars change every iteration
 * (Lambda = epoch/n_epoch*ratio, training_script.md:148; Adam's bias corrections, train.py:368):
 * a captured training step reads them from dev_scalars[4] = {Lambda, lr/(1-beta1^step),
 * sqrt(1-beta2^step), 0}, which n2n_set_step_scalars refreshes before each replay. */
int n2n_set_step_scalars(float* dev_scalars, float lam, float lr, float beta1, float beta2, int step,
                         void* stream);
int n2n_loss_n2n_fwdbwd_dev(const float* out, const float* sub2, const float* den1, const float* den2,
                            const float* dev_scalars, float grad_scale, int64_t count, float* loss3,
                            float* grad, void* workspace, void* stream);
int n2n_adam_multi_dev(const int64_t* table, int ntensors, const int32_t* blocks, int nblocks,
                       const float* dev_scalars, float beta1, float beta2, float eps, float grad_scale,
                       void* stream);

/* ------------------------------------------------------------------------- *
 * Evaluation — evaluation.py:82-83, evaluation_704.py:57-120, utils_eval.py:19-53.
 * ------------------------------------------------------------------------- */
/* pred (fp32 planes, [count]) -> uint8: clamp(0,1) then clip(p*255 + bias, 0, 255)
 * truncated; bias = 0.5 (evaluation.py:83) or 0 (evaluation_704.py:120). */
int n2n_quantize_u8(const float* pred, uint8_t* out, int64_t count, float bias, void* stream);
/* evaluation_704.py:105-112: acc[r0+y, c0+x] += clamp(pred[y,x],0,1) * wm[y,x], cnt += wm for the
 * valid (th x tw) part of a ps x ps tile; acc/cnt are H x W fp32 planes. */
int n2n_tile_accumulate(const float* pred_tile, int ps, const float* weight_mask,
                        float* acc, float* cnt, int H, int W, int r0, int c0, int th, int tw,
                        void* stream);
/* evaluation_704.py:114-120: out = clip(acc / (cnt==0 ? 1 : cnt) * 255, 0, 255) truncated. */
int n2n_tile_finalize_u8(const float* acc, const float* cnt, uint8_t* out, int64_t count, void* stream);

/* Batched device form of the same tiling (evaluation_704.py:82-120) for `batch` equally sized uint8 images [batch][H][W]:
 * gather = cut tile (ty, tx) at (ty*stride, tx*stride), /255, extend to ps x ps as np.pad(mode='reflect') does (periodic
 * when the pad exceeds the patch) -> tiles fp32 [batch * T][ps][ps], T = ceil(H/stride) * ceil(W/stride), row-major;
 * blend = per output pixel, the weighted sum over its covering tiles in that order, / weight sum, *255, truncated. */
int n2n_tile_gather_u8(const uint8_t* images, int batch, int h, int w, int ps, int stride, float* tiles, void* stream);
int n2n_tile_blend_u8(const float* pred_tiles, const float* weight_mask, int batch, int h, int w, int ps, int stride,
                      uint8_t* out, void* stream);

size_t n2n_psnr_ssim_workspace_bytes(int batch, int h, int w, int channels);
/* utils_eval.py:19-53 on `batch` pairs of H x W x C uint8 images (interleaved
 * HWC as PIL/numpy hold them, C in {1,3}).  result: [batch][2] doubles =
 * {psnr (dB, +inf when identical), ssim}. */
int n2n_psnr_ssim_u8(const uint8_t* a, const uint8_t* b, int batch, int h, int w, int channels,
                     double* result, void* workspace, void* stream);

/* ------------------------------------------------------------------------- *
 * Device-side data path — train.py:208-228 (DenoiseDataset), finetune.py:94-150
 * (DenoisePatchDataset): the training images stay resident in device memory as the
 * reference holds them (float32 H x W x C, 0..255); one launch cuts `batch` patches
 * out[b][c][y][x] = images[sel[b][0]][(sel[b][1]+y) * W + sel[b][2]+x][c] * scale.
 * images: device array of device pointers; dims_hw: int32 [nimg][2]; sel: int32
 * [batch][3] = (image, top, left) — the caller draws the coordinates (host RNG as
 * the reference's np.random.randint) and guarantees they are in range.
 * ------------------------------------------------------------------------- */
int n2n_crop_patches(const float* const* images, const int32_t* dims_hw, const int32_t* sel, int batch,
                     int channels, int patch, float scale, float* out, void* stream);

/* ------------------------------------------------------------------------- *
 * Probes (test / bring-up only): a bare tcgen05 GEMM used to validate the
 * shared-memory descriptor encodings the convolution kernels rely on.
 * ------------------------------------------------------------------------- */
int n2n_probe_umma(int variant, const void* a_bf16, const void* b_bf16, float* d,
                   int m, int n, int k, void* stream);
/* tcgen05.mma issue-rate probe: cycles for `iters` back-to-back M=128 x N x K=16 bf16 MMAs per CTA
 * on shared-memory operands in K-major layout 0 (SWIZZLE_32B), 1 (SWIZZLE_128B) or 2 (SWIZZLE_64B),
 * round-robin over `naccum` accumulators.  cycles_dev: int64[nblocks]. */
/* Diagnostic: device int64[8] that CTA 0 of every subsequent bf16 tap-GEMM launch fills with stall
 * cycle counters (NULL disables): producer-wait, producer-total, mma-wait-data, mma-wait-accumulator,
 * mma-total, tiles, epilogue-wait, epilogue-total. */
int n2n_debug_stall_buffer(long long* dev_counters);
int n2n_probe_mma_rate(int layout, int n, int iters, int naccum, long long* cycles_dev, int nblocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* N2N_B200_H */
