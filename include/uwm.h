/*
 * uwm.h — C ABI of libuwm_b200.so: the B200 (sm_100a) implementation of the
 * UNet watermark-mask inference hot path of Dave-he/unet-watermark.
 *
 * The reference has no FFI of its own: its seam is the Python nn.Module returned by
 * `create_model_from_config(cfg)` (reference src/models/unet_model.py:93-120) and called as
 * `self.model(input_tensor)` (reference src/predict.py:339, :611).  This header is the
 * C-ABI a host binds *underneath* that seam; `unet_watermark_b200/unet_model.py` is the
 * ctypes host that keeps the reference's Python surface.  See INTEGRATION.md.
 *
 * Conventions
 *   - plain C: pointers, ints, sizes.  No torch / C++ types.
 *   - every pointer named `d_*` is a DEVICE pointer owned by the caller.
 *   - activations are NHWC bf16 with an explicit pixel pitch (elements between
 *     consecutive pixels, >= channels) so that a tensor can live inside a wider
 *     concat buffer.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - every function returns 0 on success, a negative UWM_E* code otherwise, and
 *     `uwm_last_error()` returns the message for the calling thread.
 *   - there is NO CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef UWM_H_
#define UWM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UWM_OK            0
#define UWM_EINVAL       -1   /* bad argument / unsupported shape            */
#define UWM_ECUDA        -2   /* CUDA runtime / driver error                 */
#define UWM_ENOMEM       -3   /* workspace allocation failed                 */
#define UWM_ESTATE       -4   /* weights not set, wrong batch, ...           */

/* input formats of uwm_model_forward / uwm_prep_input */
#define UWM_IN_F32_NCHW   0   /* float32 [B,3,H,W], already ImageNet-normalised (what
                                 get_val_transform feeds the net, reference src/utils/dataset.py:389-395) */
#define UWM_IN_U8_NHWC    1   /* uint8  [B,H,W,3] RGB; (x/255-mean)/std fused on the GPU        */

/* weight packing kinds reported by uwm_model_layer_desc */
#define UWM_PACK_TAPS     0   /* [cout_pad][kh*kw][cin] bf16, K index = (kh*kw_idx)*cin + c      */
#define UWM_PACK_UP2X_SHUFFLE 2 /* conv3x3 over a nearest-2x upsampled input as a sub-pixel conv on the source grid:
                                 [4*cout][3*3][cin] bf16, row (ph*2+pw)*cout + co, tap (a,b) holds the SUM (in fp32,
                                 rounded once) of w[co,:,dy,dx] over the taps with floor((ph+dy-1)/2) == a-1 and
                                 floor((pw+dx-1)/2) == b-1; bias = 4 copies (packing.pack_up2x_shuffle)          */
#define UWM_PACK_UPCAT_SUBPIXEL 3 /* conv3x3 over concat(nearest-2x(x), skip) as a sub-pixel conv on x's grid, weight
                                 slices of 64 input channels in issue order, rows (qh*2+qw)*cout + co:
                                 per 64-channel chunk of x the 9 UWM_PACK_UP2X_SHUFFLE taps, centre tap first, then
                                 row-major; then per parity plane
                                 (ph,pw) of skip, per 64-channel chunk, the taps r in {1-ph,2-ph} x c in {1-pw,2-pw}
                                 holding w[co, c_x + ci, 2(r-1)+ph-qh+1, 2(c-1)+pw-qw+1] (zero when out of the
                                 3x3 kernel); bias = 4 copies (packing.pack_upcat_subpixel).  Tap row 0 / 2 of the
                                 block neighbourhood only reaches qh = 0 / 1 (same for columns): the kernel loads
                                 and multiplies only those rows of such a slice                                   */
#define UWM_PACK_S2D_CONV 4    /* conv3x3 on a [.,2h,2w,cin] tensor stored space-to-depth [.,h,w,4*cin] (channel =
                                 (ph*2+pw)*cin + ci), producing the same layout: [4*cout][3*3][4*cin] bf16 in
                                 UWM_PACK_TAPS order over BLOCKS, row (qh*2+qw)*cout + co, entry = w[co, ci,
                                 2(r-1)+ph-qh+1, 2(c-1)+pw-qw+1] for block tap (r,c), zero outside the 3x3 kernel;
                                 bias = 4 copies.  The 1-channel head: rows 0..3 = the 4 output parities, padded
                                 to 16 rows (packing.pack_s2d_conv3x3)                                           */
#define UWM_PACK_S2_PLANES 5   /* stride-2 3x3 conv (pad 1) over the input's four parity planes: [cout][9*cin] bf16 as
                                 64-channel slices in the kernel's issue order - per plane (ph,pw), per 64-channel
                                 chunk, per halo tap (r,c) with r in {1} (ph = 0) or {0,1} (ph = 1), c likewise:
                                 w[co, ci, kr, kc], kr = 1 if ph == 0 else (0 if r == 0 else 2), kc likewise
                                 (packing.pack_s2_planes); the same 9*cin values as UWM_PACK_TAPS, reordered      */
#define UWM_PACK_TAPS_SKIP_PART 6       /* UWM_PACK_TAPS of the LAST cin_skip input channels of the conv (the skip half of a
                                          decoder conv1 run as two launches); bias = the conv's folded bias            */
#define UWM_PACK_UP2X_SHUFFLE_X_PART 7 /* UWM_PACK_UP2X_SHUFFLE of the FIRST cin - cin_skip input channels (the upsampled
                                          half); bias = zeros (the skip-half launch carried it)                         */
#define UWM_PACK_STEM_S2D 1   /* 7x7/s2 stem as 4x4/s1 over the 2x2 space-to-depth input:
                                 [64][4*4][16] bf16, channel = (ph*2+pw)*3 + c, 12..15 zero      */

const char* uwm_last_error(void);
int         uwm_abi_version(void);
/* number of CUDA kernels launched by this library in this process so far */
uint64_t    uwm_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Single operators (the kernels of the path, callable on their own; the parity tests drive
 * every distinct layer shape through these).
 * ---------------------------------------------------------------------------------------- */

/* Implicit-GEMM convolution on tcgen05 tensor cores.
 * Replaces nn.Conv2d(+folded BatchNorm2d)(+residual add)(+ReLU) as used by torchvision
 * BasicBlock/Bottleneck and smp Conv2dReLU.
 *   x   : [n,h,w,cin]  bf16, pixel pitch x_pitch
 *   wgt : [cout_pad][kh*kw*cin] bf16 (UWM_PACK_TAPS), cout_pad = cout rounded up to 16
 *   bias: [cout_pad] fp32
 *   res : optional [n,ho,wo,cout] bf16 added before ReLU (pitch res_pitch), or NULL
 *   y   : [n,ho,wo,cout] bf16, pitch y_pitch;  ho = (h+2*pad-kh)/stride+1
 * cin must be a multiple of 16, cout a multiple of 16, stride 1 or 2. */
int uwm_conv2d_nhwc_bf16(const void* d_x, int n, int h, int w, int cin, int x_pitch,
                         const void* d_wgt, const float* d_bias, int cout,
                         int kh, int kw, int stride, int pad,
                         const void* d_res, int res_pitch, int relu,
                         void* d_y, int y_pitch, void* stream);

/* smp DecoderBlock.forward's `F.interpolate(x, scale_factor=2, mode="nearest")`, `torch.cat([x, skip], 1)`
 * and conv1 (Conv2d + folded BN + ReLU) as ONE kernel: the loader of the
 * halo-resident conv reads pixel (i>>1, j>>1) of x and pixel (i, j) of skip, so neither the upsampled
 * tensor nor the concat is written to memory.
 *   x    : [n,h,w,c_x] bf16 (pitch x_pitch); upsample != 0 -> the conv runs at (2h, 2w)
 *   skip : [n,H,W,c_skip] bf16 at the conv's resolution, or NULL; weights' K order is [x channels | skip channels]
 *   wgt  : [cout][kh*kw*(c_x+c_skip)] bf16 (UWM_PACK_TAPS); 'same' padding (2*pad == k-1), stride 1
 *   y    : [n,H,W,cout] bf16.  c_x, c_skip, cout multiples of 16. */
int uwm_conv2d_upcat_nhwc_bf16(const void* d_x, int n, int h, int w, int c_x, int x_pitch, int upsample,
                               const void* d_skip, int c_skip, int skip_pitch,
                               const void* d_wgt, const float* d_bias, int cout,
                               int kh, int kw, int pad, int relu,
                               void* d_y, int y_pitch, void* stream);

/* conv3x3(pad 1) over F.interpolate(x, scale_factor=2, mode="nearest") WITHOUT a skip tensor (smp DecoderBlock 4 of the
 * default Unet: interpolate + conv1), computed on the source grid: every output pixel (2i+ph, 2j+pw) sees only a 2x2
 * neighbourhood of x, so the 9 taps collapse to 4 per parity.  One 3x3 conv x[n,h,w,cin] -> 4*cout channels with
 * UWM_PACK_UP2X_SHUFFLE weights, whose epilogue scatters channel group (ph,pw) to pixel (2i+ph, 2j+pw) of
 * y[n,2h,2w,cout] (pixel shuffle).  2.6x fewer tensor-pipe cycles than the N=cout conv at the upsampled resolution.
 * cout multiple of 16, <= 64. */
int uwm_conv2d_up2x_shuffle_nhwc_bf16(const void* d_x, int n, int h, int w, int cin, int x_pitch,
                                      const void* d_wgt /*[4*cout][9*cin]*/, const float* d_bias /*[4*cout]*/,
                                      int cout, int relu, void* d_y, int y_pitch, void* stream);

/* Sub-pixel conv with a residual added at the shuffled output position before the ReLU; cout up to 64 as above, or 128 /
 * 256 (cin a multiple of 64): then every output parity is its own N tile and streams 4 of the 9 taps.  d_res (may be
 * NULL) is [n,2h,2w,cout].  With uwm_conv2d_nhwc_bf16 on the skip source this is DecoderBlock conv1 in two launches. */
int uwm_conv2d_up2x_shuffle_res_nhwc_bf16(const void* d_x, int n, int h, int w, int cin, int x_pitch,
                                          const void* d_wgt, const float* d_bias, int cout,
                                          const void* d_res, int res_pitch, int relu, void* d_y, int y_pitch,
                                          void* stream);

/* The same sub-pixel form for the decoder conv1 WITH a skip: y = act(conv3x3(concat(nearest-2x(x), skip)) + bias).
 * x[n,h,w,c_x]; skip[n,2h,2w,c_skip]; y[n,2h,2w,cout]; c_x and c_skip multiples of 64, cout a multiple of 16, <= 64.
 * wgt: UWM_PACK_UPCAT_SUBPIXEL.  The skip source is read as four parity planes (TMA boxes with element stride 2);
 * plane (ph,pw) meets only the 2x2 taps that can reach it.  Replaces segmentation_models_pytorch
 * decoders/unet/decoder.py DecoderBlock.forward (interpolate + cat + conv1) for blocks with a skip. */
int uwm_conv2d_upcat_subpixel_nhwc_bf16(const void* d_x, int n, int h, int w, int c_x, int x_pitch,
                                        const void* d_skip, int c_skip, int skip_pitch, const void* d_wgt,
                                        const float* d_bias, int cout, int relu, void* d_y, int y_pitch,
                                        void* stream);

/* Stride-2 conv3x3 (pad 1) (+bias)(+ReLU): x[n,h,w,cin] (h, w even; cin, cout multiples of 64) -> y[n,h/2,w/2,cout].
 * wgt: UWM_PACK_S2_PLANES.  The four parity planes of x arrive as dense TMA boxes (element stride 2); each kernel tap
 * is a stride-1 tap on one plane, so the halo kernel's MMA loop runs unchanged.  Same products and K as
 * uwm_conv2d_nhwc_bf16(stride 2), different fp32 summation order.  Replaces torchvision resnet.py BasicBlock.conv1 /
 * Bottleneck.conv2 of the first block of layer2-4. */
int uwm_conv2d_s2_planes_nhwc_bf16(const void* d_x, int n, int h, int w, int cin, int x_pitch,
                                   const void* d_wgt, const float* d_bias, int cout, int relu, void* d_y,
                                   int y_pitch, void* stream);

/* conv3x3 'same' (+bias)(+ReLU) on a 16-channel tensor kept space-to-depth: x, y = [n,h,w,4*16] standing for
 * [n,2h,2w,16] (channel = (ph*2+pw)*16 + c, what uwm_conv2d_nhwc_bf16 with UWM_PACK_UP2X_SHUFFLE weights writes).
 * wgt: UWM_PACK_S2D_CONV [64][9*64].  Parity plane (ph,pw) can only meet block taps {1-ph,2-ph} x {1-pw,2-pw}: 16 of
 * the 36 (tap, plane) MMAs are issued, each with N = 64 instead of 16.  Same arithmetic as the conv at full
 * resolution (every product of the reference conv appears exactly once; fp32 accumulation order differs). */
int uwm_conv2d_s2d_nhwc_bf16(const void* d_x, int n, int h, int w, int x_pitch, const void* d_wgt,
                             const float* d_bias, int relu, void* d_y, int y_pitch, void* stream);

/* The segmentation head on such a space-to-depth tensor: x[n,h,w,4*16] -> logits fp32 [n,2h,2w] / mask uint8
 * [n,2h,2w] (either may be NULL).  wgt: UWM_PACK_S2D_CONV of the [1,16,3,3] head conv, [16][9*64]. */
int uwm_head_s2d_nhwc_bf16(const void* d_x, int n, int h, int w, int x_pitch, const void* d_wgt,
                           const float* d_bias, float* d_logits, int apply_sigmoid, uint8_t* d_mask,
                           float thr_logit, void* stream);

/* Segmentation head: conv3x3(cin -> 1, bias) + optional sigmoid + threshold + uint8 mask.
 * Replaces smp SegmentationHead and `(mask > thr)*255`
 * (reference src/predict.py:624-625).  d_logits (fp32 [n,h,w]) and d_mask (uint8 [n,h,w],
 * values 0/255) may each be NULL.  thr_logit is compared against the raw logit. */
int uwm_head_nhwc_bf16(const void* d_x, int n, int h, int w, int cin, int x_pitch,
                       const void* d_wgt /*[16][9*cin]*/, const float* d_bias /*[16]*/,
                       float* d_logits, int apply_sigmoid, uint8_t* d_mask, float thr_logit,
                       void* stream);

/* MaxPool2d(3, stride 2, padding 1) — torchvision ResNet stem pool. */
int uwm_maxpool3x3s2_nhwc_bf16(const void* d_x, int n, int h, int w, int c, int x_pitch,
                               void* d_y, int y_pitch, void* stream);

/* F.interpolate(scale_factor=2, mode="nearest") written into channels [0,c) of a (possibly
 * wider, pitch y_pitch) buffer — smp DecoderBlock's upsample fused with the concat
 * (the skip half is written in place by its producer). */
int uwm_upsample2x_nhwc_bf16(const void* d_x, int n, int h, int w, int c, int x_pitch,
                             void* d_y, int y_pitch, void* stream);

/* Input preparation: normalise (u8) / cast (f32) + 2x2 space-to-depth to bf16
 * [n,h/2,w/2,16].  Replaces the Normalize+ToTensorV2 half of get_val_transform. */
int uwm_prep_input(const void* d_in, int in_fmt, int n, int h, int w, void* d_y, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-model plan: smp.Unet(resnet34|resnet50, depth 5) forward.
 * ---------------------------------------------------------------------------------------- */
typedef struct uwm_model uwm_model;

typedef struct uwm_layer_desc {
  char    conv_key[96];   /* state-dict prefix of the conv, e.g. "encoder.layer1.0.conv1"     */
  char    bn_key[96];     /* prefix of the BatchNorm folded into it, "" if the conv has bias  */
  int32_t cin, cout, cout_pad, kh, kw, stride, pad;
  int32_t pack;           /* UWM_PACK_*                                                      */
  int32_t relu, has_residual;
  int64_t w_elems;        /* bf16 elements expected by uwm_model_set_layer                   */
  int64_t b_elems;        /* fp32 elements expected                                           */
  double  flops_per_image;/* 2*MACs of the reference conv (algorithmic, un-padded)            */
  int32_t cin_skip;       /* channels of the skip source of the reference conv (its last cin_skip input channels):
                             UWM_PACK_UPCAT_SUBPIXEL, and the two _PART packings, whose `cin` is the part's own */
  int32_t reserved;
} uwm_layer_desc;

/* encoder: 34 or 50.  decoder_channels: 5 ints (smp default 256,128,64,32,16).
 * H,W: network input size, both divisible by 32.  Allocates the activation workspace for
 * max_batch images on the current device. */
int    uwm_model_create(int encoder, const int* decoder_channels, int h, int w, int max_batch,
                        uwm_model** out);
int    uwm_model_destroy(uwm_model* m);
int    uwm_model_num_layers(const uwm_model* m);
int    uwm_model_layer_desc(const uwm_model* m, int i, uwm_layer_desc* d);
/* copies host-or-device packed weights into library-owned device memory */
int    uwm_model_set_layer(uwm_model* m, int i, const void* wgt_bf16, int64_t w_elems,
                           const float* bias, int64_t b_elems);
size_t uwm_model_workspace_bytes(const uwm_model* m);
int    uwm_model_num_kernels(const uwm_model* m);       /* launches per forward             */
double uwm_model_flops_per_image(const uwm_model* m);   /* algorithmic conv FLOPs           */

/* One forward pass for `batch` (<= max_batch) images.
 *   d_in     : UWM_IN_F32_NCHW or UWM_IN_U8_NHWC
 *   d_logits : fp32 [batch,1,H,W] (raw logits, or probabilities if apply_sigmoid) or NULL
 *   d_mask   : uint8 [batch,H,W] 0/255 where logit > thr_logit, or NULL
 * use_graph != 0 replays a cached CUDA graph (captured on first use per argument set). */
int    uwm_model_forward(uwm_model* m, const void* d_in, int in_fmt, int batch,
                         float* d_logits, int apply_sigmoid, uint8_t* d_mask, float thr_logit,
                         int use_graph, void* stream);

/* Debug/parity tap: copy the bf16 NHWC output of plan tensor `name` (e.g. "encoder.layer1",
 * "decoder.blocks.0") for the last forward into d_dst as dense [batch,h,w,c]; returns dims. */
int    uwm_model_read_tensor(uwm_model* m, const char* name, int batch, void* d_dst,
                             int64_t dst_bytes, int* h, int* w, int* c, void* stream);

/* Per-kernel device timing of one eager forward (CUDA events around every launch).
 * names: caller buffer of n_max*64 chars; ms: n_max floats.  Returns number of kernels. */
int    uwm_model_profile(uwm_model* m, const void* d_in, int in_fmt, int batch,
                         float* d_logits, uint8_t* d_mask, float thr_logit,
                         char* names, float* ms, double* flops, double* bytes, int n_max,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * Image-side kernels of the reference's mask path: ragged batches of images of
 * different sizes travel in one packed buffer described by a table of uwm_image_desc.  Every entry takes the
 * table twice: h_desc (host copy, used for launch geometry and validation) and d_desc (device copy, read by the
 * kernels).
 * ---------------------------------------------------------------------------------------- */
typedef struct uwm_image_desc {
  int64_t offset;     /* start of the image inside the packed buffer, in elements of that buffer */
  int32_t width;      /* pixels */
  int32_t height;
  int32_t pitch;      /* elements between consecutive rows (>= width * channels) */
  int32_t reserved;
} uwm_image_desc;

/* cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) on uint8 3-channel images, bit-exact with OpenCV's
 * 11-bit fixed-point path (incl. its INTER_AREA shortcut for an exact 2x reduction) - the A.Resize of
 * get_val_transform (reference src/utils/dataset.py:389-395, src/predict.py:598-602).
 *   d_src  : packed uint8 HxWx3 images (descriptor i: offset, width, height, row pitch in bytes)
 *   swap_rb: != 0 reads BGR (cv2.imread order) and writes RGB (reference src/predict.py:595 cvtColor folded in)
 *   d_dst  : uint8 [n, dst_h, dst_w, 3] - the UWM_IN_U8_NHWC input of uwm_model_forward */
int uwm_resize_bilinear_u8(const uint8_t* d_src, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc, int n,
                           int dst_w, int dst_h, int swap_rb, uint8_t* d_dst, void* stream);

/* (cv2.resize(map, (W0, H0)) > threshold) * 255 per image (reference src/predict.py:620-625): bilinear resize of the
 * network's float output to each image's original size with OpenCV's float operation order, then binarisation.
 *   d_maps : fp32 [n, src_h, src_w] (logits, or probabilities for the sigmoid convention)
 *   desc i : where mask i goes in d_masks (offset, width = W0, height = H0, row pitch in bytes)
 *   d_masks: packed uint8 masks {0,255} (or NULL);  d_resized_f32: optional parity tap, same layout in floats */
int uwm_mask_upscale_threshold(const float* d_maps, int n, int src_w, int src_h, const uwm_image_desc* h_desc,
                               const uwm_image_desc* d_desc, float threshold, uint8_t* d_masks, float* d_resized_f32,
                               void* stream);

/* reference WatermarkPredictor._optimize_mask (src/predict.py:161-301) on a ragged batch of uint8 masks, in place:
 * threshold 127, elliptical open / close / dilate sequence of the chosen strategy, 8-connected components and the
 * area rules (largest component | area > 200 | > 50 | > 100).  Bit-exact with the OpenCV calls of the reference.
 *   mode: 0 'watermark' (:232-273), 1 'text' (:192-230), 2 'mixed' (:275-301)
 * d_workspace: uwm_mask_postprocess_workspace(h_desc, n) bytes of device scratch. */
size_t uwm_mask_postprocess_workspace(const uwm_image_desc* h_desc, int n);
int uwm_mask_postprocess(uint8_t* d_masks, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc, int n, int mode,
                         void* d_workspace, size_t workspace_bytes, void* stream);

/* reference _analyze_text_features (src/predict.py:443-508): d_out[i] = {components scoring > 0.5, components} of
 * mask i.  score_mask: bit (3*ia + ib)*3 + ic set when partial scores (aspect class ia, density class ib, area class
 * ic; 0 = best class) sum to more than 0.5 in the reference's float arithmetic (tabulated by the host). */
int uwm_mask_text_features(const uint8_t* d_masks, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc, int n,
                           unsigned score_mask, int32_t* d_out, void* d_workspace, size_t workspace_bytes, void* stream);

/* Single operations for the parity tests (same kernels as uwm_mask_postprocess).
 * uwm_mask_morphology: cv2.erode / dilate / morphologyEx(OPEN | CLOSE) with cv2.getStructuringElement(shape, (w, h)),
 *   default anchor and border, `iterations` times, in place.  op 0 erode, 1 dilate, 2 open, 3 close.
 * uwm_mask_components: cv2.connectedComponentsWithStats(mask, connectivity=8) as per-pixel root index (-1 =
 *   background) plus, AT each root pixel, area, OpenCV's label-order key and (optional) bbox {x0,y0,x1,y1}. */
int uwm_mask_morphology(uint8_t* d_masks, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc, int n, int op,
                        int shape, int ksize_w, int ksize_h, int iterations, void* d_workspace, size_t workspace_bytes,
                        void* stream);
int uwm_mask_components(const uint8_t* d_masks, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc, int n,
                        int32_t* d_labels, int32_t* d_area, int32_t* d_order, int32_t* d_bbox, void* d_workspace,
                        size_t workspace_bytes, void* stream);

/* d_out[i] = {foreground pixels, 8-connected components, area of the largest component} of mask i - the statistics
 * reference src/scripts/model_selector.py:171-197 takes from cv2.connectedComponentsWithStats. */
int uwm_mask_component_summary(const uint8_t* d_masks, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc, int n,
                               int32_t* d_out, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- training-step glue (BASELINE.json configs[4]; reference src/train.py:68-127 runs the smp Unet under
 * model.train(): every Conv2dReLU is conv -> BatchNorm2d with BATCH statistics -> ReLU, every residual block ends in
 * BatchNorm -> add -> ReLU).  NHWC bf16 rows of c channels (c a multiple of 8, <= 2048), dense (pixel pitch c). ---- */

/* y = [relu](BatchNorm2d_train(x) [+ residual]).  d_save[4][c] receives {batch mean, 1/sqrt(biased var + eps),
 * scale = gamma * rstd, shift = beta - mean * scale} for the backward; running_mean / running_var (both or neither)
 * are updated in place like torch.nn.BatchNorm2d (momentum, unbiased variance).  d_ws: UWM_BN_WS_SLOTS*2*c doubles that
 * must be ZERO on entry and are left zero on exit (cross-block fp64 sums, spread over slots to keep atomics apart). */
#define UWM_BN_WS_SLOTS 32
int uwm_bn_train_forward_nhwc_bf16(const void* d_x, long long pixels, int c, const float* d_gamma, const float* d_beta,
                                   float* d_running_mean, float* d_running_var, float momentum, float eps,
                                   const void* d_residual, int relu, void* d_y, float* d_save, double* d_ws, void* stream);

/* Backward of the above: g = dy * [y > 0] (when relu), dbeta = sum g, dgamma = sum g * xhat,
 * dx = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat)), d_residual = g (when has_residual; d_y is then the
 * forward's output, otherwise the ReLU mask is recomputed from x and d_y may be NULL).  d_coef: 2*c floats scratch;
 * d_ws as above. */
int uwm_bn_train_backward_nhwc_bf16(const void* d_dy, const void* d_x, const void* d_y, long long pixels, int c,
                                    const float* d_save, int relu, int has_residual, void* d_dx, void* d_dres,
                                    float* d_dgamma, float* d_dbeta, float* d_coef, double* d_ws, void* stream);

/* Backward of uwm_upsample2x_nhwc_bf16: dx[n,h,w,0:c) = sum of the 2x2 block of dy[n,2h,2w,0:c) (dy may be the leading
 * channel slice of the concat's gradient: pixel pitch dy_pitch). */
int uwm_upsample2x_backward_nhwc_bf16(const void* d_dy, int n, int h, int w, int c, int dy_pitch, void* d_dx, int dx_pitch,
                                      void* stream);

/* Backward of uwm_maxpool3x3s2_nhwc_bf16 / torch MaxPool2d(3, 2, 1): d_x is the pool's INPUT [n,h,w,c] (dense), d_dy the
 * gradient of its output [n,(h-1)/2+1,(w-1)/2+1,c]; every window's gradient goes to its first maximum in row-major scan
 * order (torch's arg-max rule, recomputed instead of saved); overlapping windows add in fp32. */
int uwm_maxpool3x3s2_backward_nhwc_bf16(const void* d_dy, const void* d_x, int n, int h, int w, int c, void* d_dx,
                                        void* stream);

/* Filters of one training-step conv: fp32 [cout][cin][kh][kw] (the nn.Conv2d parameter) -> bf16 UWM_PACK_TAPS operands
 * d_fwd [cout][kh*kw][cin] (for uwm_conv2d_nhwc_bf16 on x) and, unless NULL, d_dgrad [cin][kh*kw][cout] with the taps
 * flipped: uwm_conv2d_nhwc_bf16 on dy with d_dgrad is the data gradient of a stride-1 'same' conv. */
int uwm_pack_train_weights(const float* d_w, int cout, int cin, int kh, int kw, void* d_fwd, void* d_dgrad, void* stream);

/* dst[p, 0:c) = src[p, 0:c) for p < pixels; rows of src / dst are src_pitch / dst_pitch bf16 apart (the skip half of
 * smp DecoderBlock's torch.cat([x, skip], 1), and the skip's slice of the concat's gradient); 16-byte aligned bases. */
int uwm_copy_channels_nhwc_bf16(const void* d_src, long long pixels, int c, int src_pitch, void* d_dst, int dst_pitch,
                                void* stream);

/* ---- bench tools: exported only by the tools build of the library (-DUWM_BENCH_TOOLS; python -m
 * unet_watermark_b200.build --tools -> lib/libuwm_b200_tools.so).  The product library has none of these, nor the
 * UWM_DBG pipeline-isolation switches. ---- */
#ifdef UWM_BENCH_TOOLS
/* Sizing micro-benchmark (tools/gpu_microbench.py), not on the product path: each of `blocks` CTAs issues
 * `iters` x 4 back-to-back tcgen05.mma (M=128, N=n, K=16) and writes its clock64() delta to d_cycles. */
int    uwm_debug_mma_rate(int n, int iters, int distinct_stages, int mode, int blocks, long long* d_cycles,
                          void* stream);

/* Bench-only: when non-NULL, CTA 0 of every halo conv launched afterwards writes clock64() stamps of its pipeline
 * events into d_trace[0..600) (tools/gpu_trace.py decodes them).  NULL switches tracing off. */
int    uwm_debug_set_trace(long long* d_trace);

/* Sizing micro-benchmark: cycles of one synchronisation primitive (tools/gpu_microbench3.py); d_out: 16 int64. */
int    uwm_debug_prim_cost(int which, int iters, long long* d_out, void* stream);

/* Sizing micro-benchmark: mbarrier ping-pong between two warps, cycles for `iters` round trips. */
int    uwm_debug_handshake(int iters, int variant, int blocks, long long* d_cycles, void* stream);

#endif /* UWM_BENCH_TOOLS */

#ifdef __cplusplus
}
#endif
#endif /* UWM_H_ */
