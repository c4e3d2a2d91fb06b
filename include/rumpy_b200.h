/* rumpy_b200 -- C ABI of the B200-native EDSR/RCAN trunk (sm_100a).
 *
 * This is the drop-in boundary.  The reference (um-dsrg/RUMpy) has no native code:
 * its trunk calls stock torch.nn modules.  Each entry point below replaces one of those call sites;
 * the Python mirror of the reference's block library (rumpy_b200/SISR/models/advanced/ *.py) binds them
 * through ctypes (see INTEGRATION.md for the stub a RUMpy maintainer would add).
 *
 * Conventions
 *   - plain C types only: device pointers as void* / float*, sizes as int, the CUDA stream as void*
 *     (a cudaStream_t); no torch / C++ types cross this boundary.
 *   - every function returns 0 on success, a negative code on failure; the message is available from
 *     rumpy_last_error() (thread-local).  Nothing throws, nothing calls exit().
 *   - the caller owns every buffer (activations, packed weights, workspaces).  All work is enqueued on
 *     the caller's stream: the calls compose with CUDA-graph capture, autograd streams and NCCL streams.
 *   - activations are NHWC: bf16 operand tensors (64 channels = 128 contiguous bytes) and fp32
 *     residual-stream tensors; channel counts on the tensor-core path are multiples of 64.
 *   - there is NO CPU fallback: on a machine without an sm_100 device every compute entry point fails.
 *
 * Reference citations are relative to /root/reference (um-dsrg/RUMpy v1.0).
 */
#ifndef RUMPY_B200_H_
#define RUMPY_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define RUMPY_B200_VERSION 100

/* error codes */
#define RUMPY_OK 0
#define RUMPY_ERR_ARG (-1)     /* bad argument (shape / alignment / null pointer) */
#define RUMPY_ERR_DEVICE (-2)  /* no sm_100 device, or the driver lacks cuTensorMapEncodeTiled */
#define RUMPY_ERR_CUDA (-3)    /* a CUDA runtime / driver call failed */

int rumpy_version(void);
const char* rumpy_last_error(void);
/* 0 when the current device is an sm_100 part and the TMA driver entry point resolves. */
int rumpy_device_check(void);

/* flags for rumpy_conv3x3 (epilogue fusions) */
#define RUMPY_CONV_RELU 1u     /* v = max(v,0)                         (nn.ReLU(True), architectures.py:150) */
#define RUMPY_CONV_POOL 32u    /* emit per-tile channel sums for CALayer's AdaptiveAvgPool2d(1) (:32)        */

/* OIHW fp32 conv weights -> packed bf16 tensor-core operand.
 *   dgrad == 0: P[tap][row][ci]; rows_padded >= cout rows (extra rows zero).  shuffle_r > 1 orders the rows
 *               so that output chunk q = i*r+j holds PixelShuffle sub-pixel (i,j)  (common.py:33,40).
 *   dgrad == 1: P[tap][ci][co] with taps rotated 180 degrees -- the operand of dX = conv_transpose(g, W).
 * Replaces: the implicit weight layout of nn.Conv2d (common.py:6-9).  w_packed holds 9*rows*k bf16. */
int rumpy_pack_conv3x3(const float* w_oihw, void* w_packed, int cout, int cin, int rows_padded, int shuffle_r,
                       int dgrad, void* stream);
int rumpy_pack_bias(const float* bias, float* bias_packed, int cout, int rows_padded, int shuffle_r, void* stream);

/* 3x3, stride 1, zero pad 1 convolution on tcgen05 tensor cores with a fused epilogue:
 *     v = alpha * act(conv(x, W) + bias);  v = mask > 0 ? v : 0;  v += residual
 *   x_bf16    [N,H,W,Cin] bf16, or when in_unshuffle_r = r > 1 the tensor [N,H*r,W*r,Cin/r^2] read through
 *             a pixel-unshuffle (gradient of PixelShuffle, SURVEY 8a')
 *   w_packed  from rumpy_pack_conv3x3 (rows = Cout, k = Cin)
 *   bias      [Cout] fp32 in packed-row order, or NULL
 *   residual  [N,H,W,Cout] fp32 or NULL          (`res += x`, architectures.py:83,123,174; common.py:72-73)
 *   mask      [N,H,W,Cout] bf16 or NULL           (ReLU backward: the saved post-ReLU activation)
 *   y_bf16    [N,H,W,Cout] bf16 or NULL; with out_shuffle_r = r > 1 the tensor [N,H*r,W*r,Cout/r^2] written
 *             in pixel-shuffled order (conv + nn.PixelShuffle of common.Upsampler, common.py:30-41)
 *   y_f32     [N,H,W,Cout] fp32 or NULL
 *   pool_partial [N][rumpy_pool_rows(H, W, Cout)][Cout] fp32, required with RUMPY_CONV_POOL (Cin == Cout)
 * Replaces: nn.Conv2d.forward at common.py:6-9 (+ the elementwise ops fused above); with dgrad-packed
 * weights it is also the conv's input-gradient (autograd's convolution_backward, base_architecture.py:432). */
int rumpy_conv3x3(const void* x_bf16, const void* w_packed, const float* bias, const float* residual,
                  const void* mask_bf16, void* y_bf16, float* y_f32, float* pool_partial, int N, int H, int W,
                  int Cin, int Cout, int in_unshuffle_r, int out_shuffle_r, unsigned flags, float alpha,
                  void* stream);

/* Weight gradient of the 3x3 conv on tensor cores:
 *     dW[co][ci][ky][kx] (=|+=) alpha * sum_{n,y,x} g[n,y,x,co] * x[n,y+ky-1,x+kx-1,ci]      (OIHW fp32)
 *   g_bf16 [N,H,W,Cout] bf16 (or [N,H*r,W*r,Cout/r^2] read through a pixel-unshuffle when g_unshuffle_r = r > 1,
 *   i.e. the gradient of conv + PixelShuffle), x_bf16 [N,H,W,Cin] bf16; Cin, Cout multiples of 64.
 *   workspace: rumpy_conv3x3_wgrad_workspace(...) bytes of device memory (job descriptors + K-split partials).
 * Replaces: the weight half of autograd's convolution_backward for common.default_conv (common.py:6-9;
 * base_architecture.py:432).  Deterministic: K splits are summed in a fixed order. */
long long rumpy_conv3x3_wgrad_workspace(int N, int H, int W, int Cin, int Cout);
int rumpy_conv3x3_wgrad(const void* g_bf16, const void* x_bf16, float* dw_oihw, void* workspace, int N, int H, int W,
                        int Cin, int Cout, int g_unshuffle_r, float alpha, int accumulate, void* stream);

/* Thin tail conv C -> cout_real (<=16) on tensor cores; w_packed has 16 (zero padded) rows; output is the
 * reference's fp32 NCHW tensor.  Replaces: tail.1 = default_conv(n_feats, out_feats, 3) (architectures.py:165). */
int rumpy_conv3x3_tail(const void* x_bf16, const void* w_packed, const float* bias16, float* y_nchw, int N, int H,
                       int W, int Cin, int cout_real, void* stream);

/* Head conv in_feats (<=4) -> C on CUDA cores in fp32: NCHW fp32 in, NHWC fp32 + bf16 out.
 * Replaces: head.0 = default_conv(in_feats, n_feats, 3) (architectures.py:153,172). */
int rumpy_head_conv(const float* x_nchw, const float* w_oihw, const float* bias, float* y_f32, void* y_bf16, int N,
                    int H, int W, int Cin, int C, void* stream);

/* Channel attention + RCAB skip, fused: y = sigmoid(W2 relu(W1 mean(u)+b1)+b2); x_out = x_in + u*y.
 * `pool_partial` comes from rumpy_conv3x3(..., RUMPY_CONV_POOL) on the same N,H,W.  u is fp32 (u_is_f32) or
 * bf16 NHWC.  save_* (N x C / N x Cr fp32) may be NULL; they keep mean / hidden / y for backward.
 * Replaces: CALayer.forward (architectures.py:41-44) and `res += x` (:83). */
int rumpy_ca_apply(const float* pool_partial, const void* u, int u_is_f32, const float* x_in, const float* w1,
                   const float* b1, const float* w2, const float* b2, float* x_out, void* x_out_bf16,
                   float* save_mean, float* save_hid, float* save_y, int N, int H, int W, int C, int Cr,
                   void* stream);

/* Layout plumbing between the reference's fp32 NCHW tensors and the kernels' NHWC tensors (block-level use;
 * whole networks convert inside the head / tail convs).  Either output of nchw_to_nhwc may be NULL. */
int rumpy_nchw_to_nhwc(const float* x_nchw, float* y_f32, void* y_bf16, int N, int C, int H, int W, void* stream);
int rumpy_nhwc_to_nchw(const void* x_nhwc, int x_is_bf16, float* y_nchw, int N, int C, int H, int W, void* stream);
/* Rows per image of a pool_partial buffer (two per conv tile; the tile geometry depends on C).  Pure host arithmetic. */
int rumpy_pool_rows(int H, int W, int C);

/* Per-image channel sums in pool_partial format, for a stand-alone CALayer.forward (architectures.py:41-42:
 * nn.AdaptiveAvgPool2d(1)); inside an RCAB the sums come from rumpy_conv3x3(..., RUMPY_CONV_POOL). */
int rumpy_pool_sum(const float* x_nhwc, float* pool_partial, int N, int H, int W, int C, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Whole-network executor: one call enqueues every kernel of RCAN.forward / EDSR.forward
 * (architectures.py:171-176, 236-241).  `params` are the module's fp32 parameters as device pointers in
 * state_dict order; `packed` (rumpy_net_packed_bytes) and `workspace` (rumpy_net_workspace_bytes)
 * are caller-owned device buffers.  The handle caches the per-layer launch plan (TMA descriptors) for the last
 * (packed, workspace, N, H, W, training) it saw; a steady-state forward only launches kernels and is
 * CUDA-graph capturable.  arch: 0 = RCAN (n_groups x n_blocks RCABs), 1 = EDSR (n_blocks ResBlocks),
 * 2 = HAN (HAN.forward architectures.py:368-392 with LAM_Module / CSAM_Module HAN_blocks.py:8-76:
 * RCAN's groups in the trunk kernels, then layer attention over the 11 stacked group / body outputs -> last_conv,
 * channel-spatial attention of the body output, last(cat) + head skip; n_feats 64, n_groups 10 as the reference fixes
 * them; params in the reference's order head | body | csa.gamma, csa.conv | la.gamma | last_conv | last | tail;
 * trains too: backward through the per-layer kernels with the layer-attention gradients injected between the groups).
 * x: fp32 NCHW [N,in_feats,H,W] -> y: fp32 NCHW [N,out_feats,H*scale,W*scale], the reference's tensors.
 * training != 0 keeps every activation backward needs in the workspace.
 * ------------------------------------------------------------------------------------------------- */
int rumpy_net_create(void** net, int arch, int n_feats, int n_groups, int n_blocks, int reduction, int scale,
                     float res_scale, int in_feats, int out_feats, int u_f32);
int rumpy_net_destroy(void* net);
int rumpy_net_num_params(void* net);
int rumpy_net_num_launches(void* net);          /* kernels per forward of the cached plan */
int rumpy_net_num_launches_backward(void* net); /* kernels per backward of the cached plan */
/* How the 64-channel body (every default_conv 64->64, CALayer, RCAB / group / global skips:
 * architectures.py:81-84, 121-124, 172-174; common.py:71-75) of the cached plan is executed:
 *   0  one kernel per layer (any shape);
 *   1  ONE persistent tile-stationary dataflow kernel for the whole body (<= 4 tiles per SM): neighbour tiles
 *      hand over through per-tile epochs in global memory, fp32 residual stream in tensor memory;
 *   2  one thread-block CLUSTER per image: activations stay in (distributed) shared memory across all layers,
 *      halos and the channel-attention pool travel by st.async between the CTAs of the cluster;
 *   3  (option "band", experiment) the cluster layout with the MMA roles swapped: weights in tensor memory as the
 *      A operand, up to 144 pixels as N (trunk_band.cuh).
 * The mode is picked per (N,H,W) when the plan is built; all of them compute the same layer program. */
int rumpy_net_trunk_mode(void* net);

/* Per-handle execution options.  The library has NO process-global switches: every knob lives in the net handle and
 * applies to the calls made with that handle only (re-entrant; two handles with different options can run side by
 * side).  An option change takes effect with the next forward (the cached launch plan is rebuilt when it depends on
 * the option).  Names (value): "trunk" (1; 0 = one kernel per layer), "cluster" (1; 0 = dataflow kernel only),
 * "cluster_groups" (2 | 4 epilogue groups of the cluster kernel), "cluster_split" (1; 0 = one hand-over barrier per
 * layer), "band" (0; 1 = role-swapped band kernel,
 * experiment), "trunk_bwd" (1; 0 = per-layer backward), "fused_ca" (0), "wgrad_chunks" (4; 1..8), "wgrad_tiles_per_split"
 * (64), "pdl" (1), "trunk_sync_mode" (8), "infer_u_bf16" (1; 0 = the per-layer inference path keeps the
 * pre-attention activation in fp32), "cluster_dbg" / "conv_dbg" (0; timing experiments that switch parts of the cluster /
 * per-layer conv kernel off -- results are garbage, tools/gpu_conv_bound_probe.py).  Unknown names return
 * RUMPY_ERR_ARG / -1. */
int rumpy_net_set_option(void* net, const char* name, long long value);
long long rumpy_net_get_option(void* net, const char* name);
/* Measurement hook (bench.py's roofline timing): two caller-owned CUDA events (cudaEvent_t) recorded on the launching
 * stream right before / after the trunk kernel of every forward of this handle; NULL, NULL = off. */
int rumpy_net_set_trunk_events(void* net, void* ev_start, void* ev_stop);
/* Diagnostics: caller-owned device buffer of int64 that this handle's kernels fill with clock64 stamps; layers > 0
 * additionally requests a per-layer timeline of that many layers from the trunk kernels.  NULL = off. */
int rumpy_net_set_timeline(void* net, void* buf, int layers);

/* Chunked weight gradients for data-parallel training: rumpy_net_backward computes the conv weight gradients last,
 * in chunks ordered from the LAST parameters to the first.  After chunk k every gradient with parameter index
 * >= first_param[k] is final (first_param is decreasing and ends with 0); if events were registered, event k is
 * recorded on the backward's stream at that point, so the caller can all-reduce that range of its flat gradient
 * buffer on another stream while the remaining chunks run (replaces nn.DataParallel's gather,
 * base_architecture.py:70-77, with an overlapped NCCL all-reduce).  rumpy_net_backward_chunks is valid once the
 * training plan exists (after rumpy_net_forward(training=1) for the current shape) and returns the chunk count. */
int rumpy_net_backward_chunks(void* net, int* first_param, int max_chunks);
int rumpy_net_set_backward_events(void* net, void* const* events /* cudaEvent_t[n_events], n_events <= 8 */, int n_events);

/* Meta-attention networks: the same trunks modulated by per-image metadata.
 *   arch 0  Q-RCAN: replaces QRCAN.__init__/forward (reference SISR/models/attention_manipulators/architectures.py
 *           :313-462) for QCALayer style 'standard' or 'modulate' (:113-116, :127-128) with optional q-nodes:
 *               RCAB(x) = x + conv2(relu(conv1(x))) * CA(.) * [attributes] * [q]
 *           params: final_body | head | per group: final_body, per block: final_body.conv_du, [q_node], body.0,
 *           body.2 | tail   (the reference's registration order);
 *   arch 1  Q-EDSR: replaces QEDSR (:496-556) / ParamResBlock (:463-493):
 *               ResBlock(x) = x + res_scale * conv2(relu(conv1(x))) * [q]
 *           params: head | final_body | per block: body.0, body.2, [attention_layer] | tail;  n_groups ignored.
 *   arch 2  Q-HAN: replaces QHAN (:643-760): HAN (rumpy_net_create arch 2) whose residual groups are Q-RCAN's;
 *           params: head | per group: final_body, per block: ... | body conv | csa | la | last_conv | last | tail.
 * q = sigmoid(FC2 act(FC1 metadata)) is the reference's 2-layer ParaCALayer (q_layer.py:5-45): num_metadata ->
 * q_hidden -> n_feats, act = ReLU iff q_relu.  block_has_q[i] != 0: block i owns one.  rumpy_net_forward(training)
 * / rumpy_net_backward also return the q-layer gradients ('modulate' combined with q-layers: inference only).  All multipliers are evaluated
 * by ONE small kernel per forward and applied inside the trunk kernels (channel-attention step / residual epilogue)
 * or, for shapes outside the trunk kernels' envelope, in the per-layer kernels' epilogues. */
int rumpy_net_create_q(void** net, int arch, int n_feats, int n_groups, int n_blocks, int reduction, int scale,
                       float res_scale, int in_feats, int out_feats, int num_metadata, int q_hidden,
                       const unsigned char* block_has_q, int modulate, int q_relu);
/* metadata: device fp32 [N][M] (the `metadata` tensor of QRCAN.forward, [N,M,1,1] squeezed); it is read by every
 * following rumpy_net_forward on that call's stream, so it must stay allocated. */
int rumpy_net_set_metadata(void* net, const float* metadata, int N, int M);
long long rumpy_net_packed_bytes(void* net, int training);
long long rumpy_net_workspace_bytes(void* net, int N, int H, int W, int training);
int rumpy_net_pack(void* net, const float* const* params, void* packed, int training, void* stream);
int rumpy_net_forward(void* net, const float* const* params, const void* packed, const float* x_nchw,
                      float* y_nchw, void* workspace, int N, int H, int W, int training, void* stream);
/* Backward of the last training forward on the same (packed, workspace, N, H, W): every gradient the reference
 * gets from `loss.backward()` (base_architecture.py:432).  dy: upstream gradient, fp32 NCHW like the output;
 * grads[i] (device pointers, state_dict order, fp32, same shapes as params[i]) are overwritten.
 * dgrad and wgrad of every 3x3 conv run on tensor cores; K-split / per-block partial sums are reduced in a fixed
 * order, so gradients are bit-reproducible run to run. */
int rumpy_net_backward(void* net, const float* const* params, const void* packed, const float* x_nchw,
                       const float* dy_nchw, float* const* grads, void* workspace, int N, int H, int W, void* stream);

/* ---- train-step kernels (base_architecture.py:40, 93-95, 425-440) -------------------------------------- */
/* loss = mean|out - y| (nn.L1Loss) into *loss_out (device float) and dy = gscale * sign(out - y) / numel
 * (dy may be NULL).  ws: rumpy_l1_workspace_floats() device floats. */
long long rumpy_l1_workspace_floats(void);
int rumpy_l1_loss_grad(const float* out, const float* y, float* dy, float* loss_out, float* ws, long long numel,
                       float gscale, void* stream);
/* nn.utils.clip_grad_norm_ coefficient over a flat gradient buffer: coef_out[0] = min(1, max_norm/(norm+1e-6)),
 * coef_out[1] = norm.  ws: 1024 device floats. */
int rumpy_grad_clip_coef(const float* grad_flat, long long n, float max_norm, float* coef_out, float* ws,
                         void* stream);
/* Fused Adam (torch.optim.Adam defaults: no weight decay / amsgrad) over flat fp32 buffers; step is 1-based;
 * the gradient is multiplied by grad_scale and, when non-NULL, by *grad_scale_dev (clip coefficient). */
int rumpy_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, int step, const float* grad_scale_dev, float grad_scale, void* stream);

/* ---- eval glue and training-patch pipeline on the device (csrc/glue.cu) ------------------
 * rumpy_psnr_y: per-image PSNR on the Y channel of jpg-style YCbCr after clipping both images to [0,1]
 *   (replaces the numpy chain ImageModelInterface.colorspace_convert base_interface.py:208-222 -> ycbcr_convert
 *   image_functions.py:72-88 -> psnr sr_tools/metrics.py:33-44).  sr, hr: device fp32 NCHW [N][3][H][W];
 *   psnr: device fp32 [N]; workspace: rumpy_psnr_y_workspace(N) bytes.  Identical images give 100 (metrics.py:41-42).
 * rumpy_quantize_u8: dst[n][y][x][c] = uint8(clip(src[n][c][y][x] * 255, 0, 255)) with numpy's truncation
 *   (the uint8 image the reference saves, sr_tools/visualization.py:31-61); bit-exact.
 * rumpy_patch_batch: one training batch of LR/HR patches from uint8 HWC images resident on the device: crop at
 *   (y, x) [LR coordinates; HR = * scale], hflip, vflip, transpose (in that order), ToTensor (u8 / 255 -> fp32 CHW)
 *   (replaces image_functions.py:287-362 random_matched_crop / random_flip_rotate + torchvision ToTensor per sample
 *   in data_handler.py:570-645).  lr_imgs / hr_imgs: DEVICE arrays of device pointers; geom: device int32 [N][6] =
 *   {image index, y, x, flags (1 hflip | 2 vflip | 4 transpose), lr_h, lr_w}; outputs fp32 [N][3][crop][crop] and
 *   [N][3][crop*scale][crop*scale]; bit-exact.
 * rumpy_bicubic_upsample: the evaluation's bicubic baseline ("LR" row of the metrics, the image saved under
 *   `bicubic/`), replaces EvalHub._low_res_prep (shared_framework/evaluation/standard_eval.py:240-275; also
 *   image_functions.py:38-41 `upsample`): per image torchvision ToPILImage (x * 255, truncated to uint8) ->
 *   PIL Image.resize((W*scale, H*scale), BICUBIC) (Pillow's 8-bit two-pass fixed-point resampler) -> ToTensor
 *   (u8 / 255).  lr_nchw: device fp32 [N][C][H][W] in [0,1] (values outside are clamped; the reference's `.byte()` is
 *   undefined there); out_nchw: device fp32 [N][C][H*scale][W*scale]; workspace: rumpy_bicubic_workspace(H, W, scale)
 *   bytes, 16-byte aligned (tap tables: one entry per output column and per group of four output rows, written by
 *   the call itself); scale 2..8, N*C <= 65535; bit-exact with Pillow.
 * rumpy_lanczos_upsample: the same baseline with Pillow's Lanczos-3 filter (`--lanczos_upsample`,
 *   standard_eval.py:252-253: `Image.resize(..., LANCZOS)`); same arguments and layout, workspace
 *   rumpy_lanczos_workspace(H, W, scale) bytes.  The tap tables are evaluated on the HOST by the call (Pillow uses libm's
 *   sin in double precision; the device's is not correctly rounded) and uploaded on the caller's stream; bit-exact
 *   with Pillow. */
long long rumpy_psnr_y_workspace(int N);
int rumpy_psnr_y(const float* sr, const float* hr, float* psnr, void* workspace, int N, int H, int W, float max_value,
                 void* stream);
int rumpy_quantize_u8(const float* src_nchw, unsigned char* dst_nhwc, int N, int C, int H, int W, void* stream);
int rumpy_patch_batch(const unsigned char* const* lr_imgs, const unsigned char* const* hr_imgs, const int* geom,
                      float* lr_out, float* hr_out, int N, int crop, int scale, void* stream);
long long rumpy_bicubic_workspace(int H, int W, int scale);
long long rumpy_lanczos_workspace(int H, int W, int scale);
int rumpy_lanczos_upsample(const float* lr_nchw, float* out_nchw, void* workspace, int N, int C, int H, int W, int scale,
                           void* stream);
int rumpy_bicubic_upsample(const float* lr_nchw, float* out_nchw, void* workspace, int N, int C, int H, int W, int scale,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RUMPY_B200_H_ */
