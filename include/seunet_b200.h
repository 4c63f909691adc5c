/*
 * seunet_b200 - C ABI of the B200-native (sm_100a) SE-UNet forward/backward hot path.
 *
 * The reference (Beryl2000/SE-UNet-AirSeg) has no FFI layer: its boundary for this path is the
 * Python nn.Module API of SE_UNet.py.  This header is the thin C ABI that the drop-in module
 * (se_unet_airseg_b200/SE_UNet.py) binds with ctypes.  Each entry point names the reference
 * interface it replaces.  Conventions:
 *   - plain pointers and sizes only, no torch types; all pointers are DEVICE pointers unless noted;
 *   - the caller owns all memory (parameters, workspace, inputs, outputs); nothing is allocated,
 *     synchronised or called back inside; work is enqueued on the given CUDA stream;
 *   - every function returns 0 on success, non-zero on error; seunet_last_error() then returns a
 *     thread-local message;
 *   - re-entrant per plan: one plan per (device, shape); no global mutable state.
 */
#ifndef SEUNET_B200_H
#define SEUNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct seunet_plan seunet_plan_t;
typedef void* seunet_stream_t; /* cudaStream_t */

/* Library identity. act_dtype: 0 = fp16 storage/operands (default build), 1 = bf16. */
int seunet_version(void);
int seunet_act_dtype(void);
const char* seunet_last_error(void);

/* ---- parameters ---------------------------------------------------------------------------
 * Parameters are ONE flat fp32 device buffer holding the 117 tensors of SE_UNet.state_dict()
 * in registration order (SE_UNet.py:108-153), each in its native
 * (Cout,Cin,kD,kH,kW) layout.  Gradients use the same flat layout. */
int64_t seunet_param_count(int in_channel, int n_classes);
/* Offset (in floats) of a named tensor, e.g. "dc5.conv1.weight"; -1 if unknown. */
int64_t seunet_param_offset(int in_channel, int n_classes, const char* name);
/* Number of tensors and, for index i, its name / element count (to cross-check against state_dict). */
int seunet_param_tensors(int in_channel, int n_classes);
const char* seunet_param_name(int in_channel, int n_classes, int i);
int64_t seunet_param_numel(int in_channel, int n_classes, int i);

/* ---- plan ---------------------------------------------------------------------------------
 * A plan fixes (batch, D, H, W, in_channel, mode) and owns the launch descriptors (TMA tensor
 * maps, tile geometry) for every layer of SE_UNet.forward (SE_UNet.py:181-238).
 * mode: 0 = inference (conv scratch shared per resolution), 1 = training (activations kept
 * for seunet_backward). */
int seunet_plan_create(seunet_plan_t** plan, int batch, int D, int H, int W, int in_channel, int n_classes,
                       int mode, int device);
void seunet_plan_destroy(seunet_plan_t* plan);
size_t seunet_plan_workspace_bytes(const seunet_plan_t* plan);
/* Bytes of the packed tensor-core weight image (depends only on in_channel). */
size_t seunet_plan_wimg_bytes(const seunet_plan_t* plan);
/* Bind the plan to a caller-owned workspace + weight image (both 256-byte aligned). */
int seunet_plan_bind(seunet_plan_t* plan, void* workspace, void* wimg, seunet_stream_t stream);

/* Re-pack the fp32 parameters into the tensor-core weight image (call after every parameter
 * update; replaces cuDNN's per-call filter transforms). */
int seunet_pack_weights(seunet_plan_t* plan, const float* params, seunet_stream_t stream);

/* SE_UNet.forward (SE_UNet.py:181-238).
 *   x        fp32, element strides x_strides[5] = (n, c, d, h, w)  (callers pass non-contiguous
 *            slices, prediction.py:102)
 *   x_offsets HOST array of `batch` element offsets of each sample relative to x (sliding windows of
 *            one resident CT volume batched into one forward), or NULL: sample n starts at n*x_strides[0]
 *   params   flat fp32 parameter buffer (see above)
 *   drop0/1  DropLayer scale factors r*C/(sum r + 0.01) of shape [batch][24] / [batch][12]
 *            (SE_UNet.py:89-97; all ones in eval mode) - drawn by the host module with the
 *            reference's CPU-generator semantics
 *   pred0/1  fp32 logits [batch][1][D][H][W], contiguous */
int seunet_forward(seunet_plan_t* plan, const float* x, const int64_t* x_strides, const int64_t* x_offsets,
                   const float* params, const float* drop0, const float* drop1, float* pred0, float* pred1,
                   seunet_stream_t stream);

/* One step of the sliding-window loop, prediction.py:102-106, for the `batch` windows of an inference (mode 0) plan:
 *   p0, p = model(x_input); pred[window] += sigmoid(p)
 * x / x_strides / x_offsets / params / drop0 / drop1 as in seunet_forward; starts = HOST [batch][3] window origins inside
 * the (X,Y,Z) accumulator volume `acc` (32-bit fixed point, units of 2^-acc_log2, see seunet_window_accumulate).
 * Equivalent to seunet_forward + seunet_window_accumulate(apply_sigmoid = 1) bit for bit, but p0 - which prediction.py
 * discards - is not computed (no head-0 side branches), and p never reaches memory. */
int seunet_forward_window(seunet_plan_t* plan, const float* x, const int64_t* x_strides, const int64_t* x_offsets,
                          const float* params, const float* drop0, const float* drop1, const int* starts, uint32_t* acc,
                          int X, int Y, int Z, int acc_log2, seunet_stream_t stream);

/* Backward of SE_UNet.forward (autograd reached from loss.backward(), train.py:246/300/439/490/602) for a mode-1 plan
 * whose last seunet_forward used the same x / params / drop factors.  dpred0/dpred1: fp32 gradients w.r.t. the two
 * logit tensors [batch][1][D][H][W]; grads: flat fp32 gradient buffer in the parameter layout (overwritten).
 * conv1.bias gradients are exactly 0 (bias cancels in the non-affine InstanceNorm); dc62 (dead code, SE_UNet.py:230)
 * gets 0 here and None in the host module. */
int seunet_backward(seunet_plan_t* plan, const float* x, const int64_t* x_strides, const int64_t* x_offsets,
                    const float* params, const float* drop0, const float* drop1, const float* dpred0,
                    const float* dpred1, float* grads, seunet_stream_t stream);

/* Optional per-launch timing with CUDA events recorded on the caller's stream (bench roofline).
 * After a forward and a stream synchronize: interval i covers the launches named by `label`
 * ("conv:dc5", "apply:dc5", "cat:ec33", "up:d1", "head", "prep"); flops = algorithmic FLOPs of a
 * conv interval (2*voxels*Cin*Cout*k^3), 0 otherwise. */
int seunet_plan_set_timing(seunet_plan_t* plan, int on);
int seunet_plan_timing_count(const seunet_plan_t* plan);
int seunet_plan_timing_get(const seunet_plan_t* plan, int i, const char** label, float* ms, double* flops);

/* Test hook: locate an intermediate tensor inside the bound workspace ("CAT1", "raw:dc5", "T0:1", ...). */
int seunet_plan_debug_buffer(const seunet_plan_t* plan, const char* name, void** ptr, int* chunks, int* level);

/* Test hook: fill every SM's shared memory with 0xFF so kernels that read stale shared memory are caught. */
int seunet_debug_poison_smem(seunet_stream_t stream);

/* ---- single-op entry points (parity tests and micro-benchmarks) --------------------------- */
/* nn.Conv3d(Cin,Cout,k,padding=dil,dilation=dil) forward on chunk-plane activations
 * (SE_UNet.py:15,42,57).  in: [N][in_chunks][D][H][W][8] 16-bit; w: fp32 (Cout,Cin,k,k,k);
 * out: conv output [N][ceil(Cout/8)][D][H][W][8]; stats: [N][COUT][2] fp64 (sum, sum sq; COUT = Cout rounded
 * up to 16/32/64), zeroed by the call, or NULL.  scratch must hold seunet_conv_scratch_bytes().
 * transpose_flip=1: the data-gradient operator (w is the FORWARD weight (Cin_op, Cout_op, k,k,k), taps mirrored).
 * bf16=1: operands (and storage-format output) are bf16 instead of the build's activation storage type.
 * grad_out=1: the output is written in the gradient format (fp32 chunk planes [N][ceil(Cout/8)][D][H][W][8]);
 * accum_out=1 (grad_out only): out += result (gradient accumulation at fan-out nodes). */
size_t seunet_conv_scratch_bytes(int Cin, int Cout, int ksize, int dil);
int seunet_conv_fprop(const void* in, int in_chunks, int in_chunk_off, const float* w, int N, int D, int H, int W,
                      int Cin, int Cout, int ksize, int dil, void* out, double* stats, void* scratch,
                      int transpose_flip, int bf16, int grad_out, int accum_out, seunet_stream_t stream);
/* Weight gradient of the same conv (autograd of SE_UNet.py:15/42/57): dw[co][ci][k][k][k] = sum x * dy.
 * x: activations (storage type) [N][x_chunks][D][H][W][8], slice of Cin channels at x_chunk_off;
 * dy: SAME storage type (tcgen05 kind::f16 needs equal operand formats) [N][dy_chunks][D][H][W][8] with at least
 * COUT/8 planes from dy_chunk_off; dw: fp32, overwritten. */
size_t seunet_wgrad_scratch_bytes(int Cin, int Cout, int ksize);
int seunet_conv_wgrad(const void* x, int x_chunks, int x_chunk_off, const void* dy, int dy_chunks, int dy_chunk_off,
                      int N, int D, int H, int W, int Cin, int Cout, int ksize, int dil, void* scratch, float* dw,
                      seunet_stream_t stream);
/* fp32 NCDHW <-> chunk-plane storage conversion helpers (tests, sliding-window driver). */
int seunet_to_chunks(const float* src, int N, int C, int D, int H, int W, void* dst, int dst_chunks, int dst_off,
                     seunet_stream_t stream);
int seunet_from_chunks(const void* src, int src_chunks, int src_off, int N, int C, int D, int H, int W, float* dst,
                       seunet_stream_t stream);

/* ---- losses and optimizer (train.py:51-76; stage combinations 597-599 / 432-435 / 238-243; AdamW 188/386/569) -------
 * The losses are ratios of BATCH-GLOBAL sums.  seunet_loss_sums reduces this rank's logits to partial sums
 *   sums[2 heads (en=pred0, de=pred1)][8] = {sum p t, sum p, sum t, sum w (p+1e-4)^0.7 t, sum w (0.2p+0.8t),
 *                                            sum w p s^2, sum w (p s + s), 0}          (p = sigmoid(logit), fp64, device)
 * (data-parallel ranks all-reduce the 16 doubles - the reference evaluates the loss on the gathered batch), and
 * seunet_loss_grad turns the global sums into the scalar stage loss and d loss / d logit.
 * stage 1: dice(de)+dice(en); 2: gul(de)+0.5 gul(en); 3: stage 2 + 0.5 (atr(en)+atr(de)).
 * per_sample (optional): [batch][2 heads][2] = per-sample (A, Bs) GUL sums used for hard-mining names (train.py:249-253). */
int seunet_loss_sums(int stage, const float* pred_en, const float* pred_de, const float* label, const float* weight,
                     const float* skel, int batch, int64_t voxels_per_sample, double* sums, double* per_sample,
                     seunet_stream_t stream);
int seunet_loss_grad(int stage, const float* pred_en, const float* pred_de, const float* label, const float* weight,
                     const float* skel, int64_t n, const double* sums, float* dpred_en, float* dpred_de, float* loss_out,
                     seunet_stream_t stream);
/* torch.optim.AdamW step (amsgrad=False) over the flat parameter / gradient buffers; grad_scale multiplies the gradient
 * first; [skip_off, skip_off+skip_len) is left untouched (dc62: grad None in the reference => never updated). */
int seunet_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int step, float grad_scale, int64_t skip_off,
                      int64_t skip_len, seunet_stream_t stream);

/* ---- sliding-window inference support (prediction.py:39-49, 69-111) ------- */
/* two_channel(img + offset): img int16 (dtype 0) or fp32 (dtype 1) of nvox voxels -> out[2][nvox] fp32,
 * HU windows [-1024,1024] and [-1000,500] scaled to [0,1]; evaluated in fp64 like the numpy reference. */
int seunet_hu_windows(const void* img, int dtype, int64_t nvox, double offset, float* out, seunet_stream_t stream);
/* Same on a slab of nvox voxels of a larger volume: channel 1 is written channel_stride floats after channel 0 (lets the
 * host copy of a volume be pipelined slab by slab with the windows that only need its first planes). */
int seunet_hu_windows_slab(const void* img, int dtype, int64_t nvox, int64_t channel_stride, double offset, float* out,
                           seunet_stream_t stream);
/* acc[window b] += sigmoid(logits[b]) for B windows of size (cd,ch,cw) starting at HOST starts[b][3]
 * inside the (X,Y,Z) accumulator volume (prediction.py:103-106).  The accumulator is 32-bit FIXED POINT in units of
 * 2^-acc_log2 (zero it before the first window; pick acc_log2 with max_overlap_count * 2^acc_log2 < 2^31, 26 for the
 * reference's 128/64 grid): integer sums are order-independent, so window batches on several streams and the partial
 * volumes of several ranks (one integer SUM all-reduce) give bit-identical results. */
int seunet_window_accumulate(const float* logits, const int* starts, int B, int cd, int ch, int cw, uint32_t* acc, int X,
                             int Y, int Z, int apply_sigmoid, int acc_log2, seunet_stream_t stream);
/* mean = acc / count (prediction.py:109), mask = mean >= threshold. counts_dev: DEVICE int[X+Y+Z] with the
 * per-axis window coverage counts. mask may be NULL; write_mean overwrites acc with the fp32 mean (same 4-byte slots). */
int seunet_window_finalize(uint32_t* acc, const int* counts_dev, int X, int Y, int Z, float threshold,
                           unsigned char* mask, int write_mean, int acc_log2, seunet_stream_t stream);

/* ---- post-processing of the mean-probability volume (prediction.py:13-37, 111-116; util.py:58-75) ----
 * scratch: caller-owned DEVICE buffer of seunet_postproc_scratch_bytes(); max_runs bounds the number of row runs (maximal
 * runs of set voxels along the last axis) the component labelling may see - D*H*ceil(W/2) is always enough. */
size_t seunet_postproc_scratch_bytes(int D, int H, int W, int64_t max_runs);
/* double_threshold_iteration(pred, h_thresh, l_thresh) - the reference's single in-place raster sweep, bit-exact - then
 * (border_frac >= 0) zeroing of the first/last int(border_frac*n) planes of axes 0 and 1 (prediction.py:112-115).
 * The bit-packed result stays in scratch for seunet_postproc_largest_component; mask_out (uint8 [D][H][W]) may be NULL. */
int seunet_postproc_dti(const float* prob, int D, int H, int W, double h_thresh, double l_thresh, double border_frac,
                        unsigned char* mask_out, void* scratch, int64_t max_runs, seunet_stream_t stream);
/* maximum_3d (util.py:58-75): largest 26-connected component (the second largest when the largest misses the slices
 * k = W/2, W/3, (W/3)*2 of the last axis), then binary_fill_holes (fill_holes != 0).  mask_in == NULL takes the bitmask left
 * in scratch by seunet_postproc_dti.  info_out (DEVICE int[16], may be NULL): [0] error (1 = more runs than max_runs),
 * [1] foreground runs, [2] voxels of the largest component, [3] of the second, [4] 1 if the second was chosen. */
int seunet_postproc_largest_component(const unsigned char* mask_in, int D, int H, int W, int fill_holes,
                                      unsigned char* mask_out, int* info_out, void* scratch, int64_t max_runs,
                                      seunet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SEUNET_B200_H */
