/* skyeye_b200.h -- C ABI of libskyeye_b200.so: the B200 (sm_100a) kernels under SkyEye's batched
 * detector forward path (backbone -> neck -> CLA -> transformer heads -> decode -> NMS).
 *
 * The reference (pure Python/PyTorch) has no FFI; its seam for this path is nn.Module.forward on
 * torch tensors plus the free function non_max_suppression.  Each entry point below names the
 * reference code it replaces (file:line under /root/reference).  The Python host package
 * (skyeye/...) binds these with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns 0 (SKB_OK) or a negative error code; skb_last_error() has the message
 *     (thread-local).  No CPU fallback: on a non-sm_100 device compute calls return SKB_ERR_ARCH.
 *   - the library never allocates or frees device memory: inputs, outputs and workspaces are
 *     caller-owned (the Python host allocates them with torch.empty and passes data_ptr()).
 *   - every launch goes to the explicit `stream` (a cudaStream_t passed as void*); no internal
 *     synchronisation, no host-visible results => CUDA-graph capturable.
 *   - activations are NHWC; a view addresses a channel slice of a wider buffer (concat fusion).
 */
#ifndef SKYEYE_B200_H_
#define SKYEYE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKB_VERSION 100 /* 0.1.0 */

#define SKB_OK 0
#define SKB_ERR_ARG (-1)         /* bad shape / alignment / null pointer */
#define SKB_ERR_ARCH (-2)        /* device is not sm_100 */
#define SKB_ERR_CUDA (-3)        /* CUDA runtime / driver error */
#define SKB_ERR_UNSUPPORTED (-4) /* valid request the kernels do not cover (message says why) */
#define SKB_ERR_WORKSPACE (-5)   /* workspace too small (see skb_*_workspace_bytes) */

#define SKB_BF16 0
#define SKB_F32 1

#define SKB_ACT_NONE 0
#define SKB_ACT_SILU 1 /* ConvolutionBlock, blocks.py:34 */
#define SKB_ACT_RELU 2 /* TransformerLayer FFN, attention.py:274 */

/* NHWC view: element (n,y,x,ch) lives at ptr + ((n*h + y)*w + x)*pitch + ch, ch < c <= pitch.
 * ptr already includes the channel offset of the slice; ptr must be 16-byte aligned. */
typedef struct skb_view {
    void* ptr;
    int32_t n, h, w, c;
    int32_t pitch; /* elements between consecutive pixels */
    int32_t dtype; /* SKB_BF16 or SKB_F32 */
} skb_view;

int skb_version(void);
const char* skb_last_error(void);
/* SKB_OK iff the current CUDA device is compute capability 10.x */
int skb_device_check(void);

/* ---- conv-BN-SiLU implicit GEMM on tcgen05 (TMA im2col tiles, TMEM accumulators) --------------
 * Replaces ConvolutionBlock.forward (skyeye/core/models/blocks.py:36-38), BottleneckBlock's residual
 * (blocks.py:88-90), the concat writes of CSPBlock/SPPBlock/FeatureNeck (blocks.py:121-123,149;
 * detector.py:215,219,224,228), nn.Conv2d 1x1 projections (attention.py:167-170; detector.py:56-59)
 * and the nn.Linear layers of TransformerLayer (attention.py:265,272-278).
 *   y = [residual +] act(conv(x, w) + bias)        (fp32 accumulate; bf16 or fp32 store)
 * x: bf16 view [N,H,W,Cin], Cin % 16 == 0.   w: bf16 [cout_pad][k*k][Cin] (K-major, BN folded),
 * cout_pad % 32 == 0 (% 64 if > 32), rows >= y->c are zero.   bias: fp32 [cout_pad].
 * k in {1,3}, stride in {1,2} (pad k/2; stride 2 needs even H, W).  y: view [N,Ho,Wo,Cout] (or
 * [N,2Ho,2Wo,Cout] when upsample2x != 0: each result is replicated 2x2 = nearest upsampling fused
 * into the store, detector.py:214,218).  residual (nullable): bf16 view shaped like y, added after
 * the activation; may alias y (in-place bottleneck update). */
int skb_conv2d_bf16(const skb_view* x, const void* w_packed, const float* bias, const skb_view* residual,
                    const skb_view* y, int32_t cout_pad, int32_t ksize, int32_t stride, int32_t act,
                    int32_t upsample2x, void* stream);

/* Debug aid, not part of the drop-in surface: CTA 0 of every following conv launch records
 * (event, clock64) pairs into this device buffer of 3*8192 int64; NULL switches tracing off. */
int skb_debug_conv_trace(void* device_buffer);

/* ---- layout / pooling / attention-gate kernels (HBM-bound) ------------------------------------ */
/* FocusBlock space-to-depth (blocks.py:170-181) fused with the NCHW fp32 -> NHWC bf16 conversion:
 * y[n, y, x, p*3 + c] = img[n, c, 2y + dy(p), 2x + dx(p)], patches TL, BL, TR, BR; channels
 * 12..y->c-1 are zero-filled.  img: fp32 [N,3,H,W] contiguous. */
int skb_focus_nchw_f32(const float* img, int32_t n, int32_t h, int32_t w, const skb_view* y, void* stream);
/* Same for a uint8 image [N,3,H,W]: the caller's `img.float() / 255` (validate.py:237-238,
 * detect.py:133-134) is fused into the load, so the host ships 1 byte per sample over PCIe. */
int skb_focus_nchw_u8(const uint8_t* img, int32_t n, int32_t h, int32_t w, const skb_view* y, void* stream);
/* Whole FocusBlock.forward (blocks.py:170-182: four strided slices + cat + Conv-BN-SiLU k3 s1) in two
 * launches.  (1) space-to-depth into a 16-channel NHWC scratch image whose rows carry one zero pixel on
 * the left and three on the right; (2) the 3x3 conv as an implicit GEMM with K = 3 row taps x 64: the
 * three horizontal taps of an output pixel are 96 CONTIGUOUS bytes of that scratch row, so the TMA
 * tensor map uses a 128-byte window sliding by one 32-byte pixel (element stride < window) and each
 * row tap is ONE full-width SWIZZLE_128B tile instead of three 32-byte-row tiles.
 * img: fp32 or uint8 [N,3,H,W] (img_dtype SKB_F32, or SKB_U8 = scaled by 1/255, validate.py:237-238).
 * w_rowtap: bf16 [cout_pad][3][4][16] (ky, kx slot 0..3 -- slot 3 zero --, focus channel 0..15 -- 12..15
 * zero); see skyeye.engine.PackedFocusConv.  y: bf16 view [N,H/2,W/2,Cout].
 * workspace: skb_focus_conv_workspace_bytes(n,h,w) bytes of scratch. */
#define SKB_U8 2
size_t skb_focus_conv_workspace_bytes(int32_t n, int32_t h, int32_t w);
int skb_focus_conv_bf16(const void* img, int32_t img_dtype, int32_t n, int32_t h, int32_t w, const void* w_rowtap,
                        const float* bias, const skb_view* y, int32_t cout_pad, int32_t act, void* workspace,
                        size_t workspace_bytes, void* stream);
/* Same, reading image n in place as the h x w window at (y0, x0) of frame f of `frames` [F,3,frame_h,frame_w]:
 * tiles_dev = device int32 [n][3] rows (f, y0, x0).  Tiled inference of 4K drone frames (BASELINE config 4; the reference
 * has no tile code -- nearest call site validate.py:234-256 -- SURVEY.md D8): the tile batch is never materialised. */
int skb_focus_conv_tiles_bf16(const void* frames, int32_t img_dtype, int32_t frame_h, int32_t frame_w, const int32_t* tiles_dev,
                              int32_t n, int32_t h, int32_t w, const void* w_rowtap, const float* bias, const skb_view* y,
                              int32_t cout_pad, int32_t act, void* workspace, size_t workspace_bytes, void* stream);
/* Inference-time pre-processing in one pass: `letterbox` (skyeye/core/data/augmentation.py:442-496:
 * cv2.resize INTER_LINEAR to new_h x new_w, constant border `pad`) + BGR->RGB + HWC->CHW (detect.py:131-132).
 * src: uint8 [h0][w0][3] BGR (device memory, row pitch src_pitch bytes); dst: uint8 [3][H][W] RGB planes (device) --
 * the layout skb_focus_conv_bf16 / skb_focus_nchw_u8 consume.  The resized image occupies rows [top, top+new_h) and
 * columns [left, left+new_w).  The bilinear resize is OpenCV's 8-bit fixed-point scheme (11-bit coefficients,
 * (b0*(S0>>4)>>16 + b1*(S1>>4)>>16 + 2) >> 2): bit-exact with cv2.resize when down-scaling or copying; when
 * up-scaling OpenCV's dispatched path differs by 1 LSB on < 0.5 % of the samples. */
int skb_letterbox_u8(const uint8_t* src, int32_t h0, int32_t w0, int32_t src_pitch, uint8_t* dst, int32_t H, int32_t W,
                     int32_t new_h, int32_t new_w, int32_t top, int32_t left, int32_t pad, void* stream);
/* nn.MaxPool2d(5, stride 1, pad 2) (blocks.py:143-144); SPP's 9 and 13 pools are cascades of it. */
int skb_maxpool5_bf16(const skb_view* x, const skb_view* y, void* stream);
/* SPPBlock's three pools (kernel 5, 9, 13; blocks.py:143-149) of the same map in one pass: x is read once, y5 / y9 / y13 are the
 * concat slices behind it (same shape as x, a common pixel pitch).  Maps up to 110 KB / 64 B pixels (41 x 41); larger ones
 * return SKB_ERR_UNSUPPORTED (three skb_maxpool5_bf16 calls do the same). */
int skb_spp_pools_bf16(const skb_view* x, const skb_view* y5, const skb_view* y9, const skb_view* y13, void* stream);
/* CombinedAttention = ChannelAttention + SpatialAttention (attention.py:37-60, 80-98, 118-130).
 * w0: fp32 [C/r][C], w1: fp32 [C][C/r], w7: fp32 [2][7][7].  workspace: fp32, see size query. */
size_t skb_cbam_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t c);
int skb_cbam_bf16(const skb_view* x, const float* w0, const float* w1, int32_t reduced, const float* w7,
                  const skb_view* y, void* workspace, size_t workspace_bytes, void* stream);
/* CrossLayerAttention core in closed form (attention.py:196-238 with R4):
 * s = scale * sum_{c in head} q * bilinear(k);  a = softmax over image rows;  o = r2 * a * bilinear(v).
 * q: [N,H,W,Cq]; k: [N,H/2,W/2,Cq]; v: [N,H/2,W/2,Cv]; o: [N,H,W,Cv]; all bf16 views. */
size_t skb_cla_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t heads);
int skb_cla_core_bf16(const skb_view* q, const skb_view* k, const skb_view* v, const skb_view* o, int32_t heads,
                      float scale, float r2, void* workspace, size_t workspace_bytes, void* stream);
/* nn.LayerNorm over channels (attention.py:268-269, 297, 302), eps 1e-5; x, y bf16 views. */
int skb_layernorm_bf16(const skb_view* x, const float* gamma, const float* beta, float eps, const skb_view* y,
                       void* stream);
/* nn.MultiheadAttention core (attention.py:298): softmax(q k^T * scale) v per (image, head), flash
 * style on tcgen05 (never materialises N x N).  qkv: bf16 view [B,H,W,3C] = in_proj output (q | k | v
 * on the channel axis); o: bf16 view [B,H,W,C]; head_dim must be 64. */
int skb_flash_attn_bf16(const skb_view* qkv, const skb_view* o, int32_t heads, float scale, void* stream);
/* Diagnostics only (no reference counterpart): with SKB_ATT_PROF=1 in the environment skb_flash_attn_bf16 runs an
 * instrumented kernel that accumulates per-role phase cycle counts; this copies (and optionally clears) the 16 counters
 * (layout: csrc/attention.cu, printed by scripts/attn_prof.py). */
int skb_debug_attn_prof(unsigned long long* out16, int32_t reset);

/* WindowedSelfAttention core (attention.py:372-395; the class is unwired in the reference):
 * per (window, head)  o = softmax(q * scale @ k^T + bias[head] (+ mask[window % n_mask])) @ v  for windows of
 * n_tok <= 64 tokens.  qkv: bf16 view [n_windows, 1, n_tok, 3C] = the qkv Linear's output (q | k | v on the channel
 * axis, head-major inside each); bias: fp32 [heads][n_tok][n_tok] = relative_position_bias_table gathered by
 * relative_position_index; mask: fp32 [n_mask][n_tok][n_tok] or NULL; o: bf16 view [n_windows, 1, n_tok, C];
 * head_dim = C / heads must be a multiple of 8 and <= 64. */
int skb_window_attn_bf16(const skb_view* qkv, const float* bias, const float* mask, int32_t n_mask, const skb_view* o,
                         int32_t heads, float scale, void* stream);
/* The same core with window partition and reverse folded into its addressing (NOT IN REFERENCE: the class is never wired,
 * SURVEY.md X5 / §8f N3; Swin-style partition of attention.py:358's input layout): qkv bf16 view [B,H,W,3C] and o bf16 view
 * [B,H,W,C] are unpartitioned feature maps, H and W multiples of `window` (<= 8); window w of an image (row-major over the
 * window grid) covers pixels (wy*window .., wx*window ..), token t of it is pixel (t / window, t % window) of that square;
 * mask [n_mask, w*w, w*w] is indexed by (window index within the image) % n_mask. */
int skb_window_attn2d_bf16(const skb_view* qkv, const float* bias, const float* mask, int32_t n_mask, const skb_view* o,
                           int32_t heads, int32_t window, float scale, void* stream);

/* ---- decode (DetectionHead.process_detections, detector.py:88-145) -----------------------------
 * raw[l]: fp32 view [B,h_l,w_l,>=na*no] (channel = a*no + o, the 1x1 head conv output); 16-byte aligned,
 *   pitch a multiple of 4 and >= na*no rounded up to 4 (rows are staged with 16-byte loads).
 * det: fp32 [B, sum_l na*h_l*w_l, no], rows ordered (level, anchor, y, x).
 * raw_out[l] (nullable): fp32 [B,na,h_l,w_l,no] = the reference's raw_outputs (detector.py:81-82).
 * anchors: host fp32 [levels][na][2] in pixels (multiplied by stride again, quirk X16). */
int skb_decode_f32(const skb_view* raw, int32_t levels, int32_t na, int32_t no, const float* anchors_host,
                   int32_t in_h, int32_t in_w, float* det, float* const* raw_out, void* stream);

/* ---- NMS ---------------------------------------------------------------------------------------
 * skb_nms_f32: torchvision.ops.nms semantics (call site skyeye/utils/metrics.py:442), bit-exact with
 * the CPU op: stable descending sort, fp32 IoU with true division and no FMA, suppress iff iou > thr.
 * boxes [n,4] xyxy, scores [n] (device).  keep: int64 [n] (device), n_keep: int32 (device). */
size_t skb_nms_workspace_bytes(int32_t n);
int skb_nms_f32(const float* boxes, const float* scores, int32_t n, float iou_thr, int64_t* keep, int32_t* n_keep,
                void* workspace, size_t workspace_bytes, void* stream);

/* skb_nms_batched_f32: the whole reference wrapper non_max_suppression (metrics.py:361-457) for a
 * batch, no host synchronisation: confidence filter, best-class / multi-label expansion, optional
 * class filter, top-30000 cap, class-offset boxes, greedy NMS, max_det.
 * pred: fp32 [B,N,5+nc] (device).  out: fp32 [B,max_det,7] rows [cx,cy,w,h,obj,cls_prob,cls_id]
 * (compat 0 = reference quirks X8; nc==1 rows use 6 columns) or [x1,y1,x2,y2,conf,cls,0] (compat 1).
 * out_count: int32 [B].  classes: host int32 [n_classes] or NULL. */
size_t skb_nms_batched_workspace_bytes(int32_t b, int32_t n, int32_t nc, int32_t multi_label);
int skb_nms_batched_f32(const float* pred, int32_t b, int32_t n, int32_t nc, float conf_thr, float iou_thr,
                        const int32_t* classes_host, int32_t n_classes, int32_t agnostic, int32_t multi_label,
                        int32_t max_det, int32_t compat, float* out, int32_t* out_count, void* workspace,
                        size_t workspace_bytes, void* stream);
/* Tiled form of the wrapper (BASELINE config 4): image b is a tile whose origin tile_xy_dev[b] = (x0, y0)
 * (device int32 [B][2]) is added to the kept rows (columns 0,1 of reference rows; 0..3 of fixed rows) in the kernel's
 * epilogue, and the rows go, zero padded, with the count in an extra trailing row, straight into the all_gather send buffer:
 * out_packed fp32 [B][max_det + 1][7], row max_det = [count, 0, ...].  No class filter. */
int skb_nms_batched_tiles_f32(const float* pred, int32_t b, int32_t n, int32_t nc, float conf_thr, float iou_thr, int32_t agnostic,
                              int32_t multi_label, int32_t max_det, int32_t compat, const int32_t* tile_xy_dev, float* out_packed,
                              int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);
/* Correct-prediction matrix of validate()'s process_batch (skyeye/cli/validate.py:71-108, IoU as box_iou, skyeye/utils/metrics.py:17-44)
 * for ONE image, on the device: labels fp32 [n_labels][5] = (cls, x1, y1, x2, y2), dets fp32 [n_dets][6] = (x1, y1, x2, y2, conf, cls),
 * iouv fp32 [n_iou] thresholds (device), correct uint8 [n_dets][n_iou] (1 = the detection is its best same-class label's
 * highest-IoU detection and that IoU >= the threshold).  Workspace: skb_match_workspace_bytes. */
size_t skb_match_workspace_bytes(int32_t n_dets, int32_t n_labels);
int skb_match_detections_f32(const float* labels, int32_t n_labels, const float* dets, int32_t n_dets, const float* iouv, int32_t n_iou,
                             uint8_t* correct, void* workspace, size_t workspace_bytes, void* stream);
/* Diagnostics only (no reference counterpart): counter_dev = device pointer of an unsigned 64-bit counter that receives the
 * number of IoU pair tests of every following skb_nms_batched*_f32 call (SURVEY.md §8d asks for pairs/s on config 5;
 * scripts/bench_nms_stress.py); NULL switches the counting kernel variant off. */
int skb_debug_nms_pair_counter(unsigned long long* counter_dev);
/* Gathered per-tile rows -> prediction tensor of the per-frame merge NMS.  gathered: fp32 [world][tiles_per_rank][max_det+1][7]
 * (rank r holds global tiles r, r + world, ...; global tile = frame * tiles_per_frame + k); pred: fp32
 * [n_frames][tiles_per_frame * max_det][5 + nc], fed to skb_nms_batched_f32 with the same compat. */
int skb_tile_merge_pred_f32(const float* gathered, int32_t world, int32_t tiles_per_rank, int32_t n_frames, int32_t tiles_per_frame,
                            int32_t max_det, int32_t nc, int32_t compat, float* pred, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SKYEYE_B200_H_ */
