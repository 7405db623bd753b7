/* densehead.h -- C ABI of libdensehead.so: the B200 (sm_100a) dense-head path of
 * WD-Leong/CV-Lite-Object-Detection (target encoding, losses, decode + NMS).
 *
 * This is the drop-in boundary.  The reference is pure Python with no FFI of its own, so each
 * entry point below names the reference routine it replaces (file:line relative to the reference
 * repository) and `INTEGRATION.md` shows the ctypes binding a maintainer adds on the reference
 * side.  Conventions:
 *
 *   - Plain C: pointers, sizes, scalars.  No framework types.  Pointers marked [dev] are CUDA device
 *     pointers (obtained zero-copy from the framework tensor, e.g. via DLPack); pointers marked
 *     [host] are small host arrays read synchronously during the call.
 *   - The caller owns every buffer.  The library keeps only an opaque handle with its scratch.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream) and the call returns without synchronising, unless stated otherwise.
 *   - Every function returns DH_OK (0) or a negative DH_ERR_* code; dh_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 *   - GT rows are float32 (cy, cx, h, w, class) normalised by the image's unpadded size
 *     `img_dim[b] = (H, W)`, padded to [B, max_boxes, 5] with a per-image count `nbox[b]`
 *     (max_boxes <= DH_MAX_BOXES_PER_IMAGE).  Target maps are float32, channel-last, batch-major.
 */
#ifndef DENSEHEAD_H_
#define DENSEHEAD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DH_OK 0
#define DH_ERR_BAD_ARG (-1)
#define DH_ERR_SHAPE (-2)
#define DH_ERR_CUDA (-3)
#define DH_ERR_CAPACITY (-4)
#define DH_ERR_NCCL (-5)

#define DH_MAX_BOXES_PER_IMAGE 256
#define DH_MAX_PYRAMID_LEVELS 8

typedef struct dh_handle_s* dh_handle_t;

/* ---- library / handle ------------------------------------------------------------------- */
const char* dh_version(void);
const char* dh_last_error(void);
/* Binds the handle to CUDA device `device` (its SM count sizes the persistent grids). */
int dh_create(dh_handle_t* out, int device);
int dh_destroy(dh_handle_t h);
/* Options: see DH_OPT_*.  Returns DH_ERR_BAD_ARG for an unknown option. */
#define DH_OPT_TMA_STORE 1   /* 1 (default): tiles leave shared memory as TMA bulk stores; 0: st.global.v4 */
#define DH_OPT_TILE_BYTES 2  /* upper bound on the shared-memory tile size in bytes (default 49152; small problems use smaller tiles) */
#define DH_OPT_CTAS_PER_SM 3 /* upper bound on persistent CTAs per SM (default 4; shared memory may allow fewer) */
#define DH_OPT_PHASE_TIMING 4 /* profiling aid: 1 = CTA 0 of each encode kernel accumulates per-phase clock64 totals; (re)sets them */
#define DH_OPT_FUSED_LOSS_KERNEL 5 /* 0 (default): stream + correct kernel when num_classes <= 128; 1: always the shared-memory target-tile kernel */
#define DH_OPT_NMS_KERNEL 6 /* 0 (default): pick per call; 1: lazy one-CTA-per-image NMS whenever the output cap is <= 1024; 2: always mask matrix on all SMs + block sweep */
#define DH_OPT_FCOS_SELECT 7 /* dh_fcos_detect candidate selection (every choice is exact and bit-identical). 0 (default): in logit space when it can -- estimate + one streaming pass + per-level finish for pre_nms_topk <= 1024, else a thread-block cluster per long level for small batches; 1: always score every pair, one CTA per (image, level); 2: logit space, one CTA per (image, level); 3: logit space, always clusters; 4: logit space, the streaming pre-select whenever it fits */
#define DH_OPT_FUSED_CHUNKS_PER_CTA 8 /* fused loss scheduler: aim at this many image-aligned chunks per persistent CTA (default 12; a chunk is always 4..8 tiles of 256 rows) */
#define DH_OPT_ENCODE_MIN_CHUNK 9 /* encoders: smallest scheduler chunk in tiles (default 2) */
#define DH_OPT_NMS_SORT 10 /* score sort ahead of the NMS. 0 (default): bucket sort in shared memory, bitonic network when the scores pile onto few buckets; 1: always the bitonic network (A/B checks) */
#define DH_OPT_LOSS_ALLREDUCE 11 /* 1: the out_total of every dh_*_encode_loss(_grad) call is summed over the ranks of the handle's communicator inside the same kernel (its last CTA exchanges the scalars through the NVLink peer mailboxes, see dh_comm_peer_import); 0 (default): out_total is this rank's sum */
#define DH_OPT_ALLREDUCE 12 /* transport of dh_allreduce_loss. 0 (default): peer mailboxes when imported, else NCCL; 1: NCCL; 2: peer mailboxes */
#define DH_OPT_FUSED_TAIL 13 /* fused loss scheduler: 1 (default) cuts the last images of a launch into finer chunks so that the tail is short; 0: uniform chunks.  Tuning aid: + 2 * f, f = 1..5, cuts a chunk's last tile into up to 2^f spans for the warps to take (default f = 2) */
#define DH_OPT_ENCODE_KERNEL 14 /* target encoders. 0 (default): pick per problem -- direct-store kernel for small outputs, shared-memory tile streamer with TMA bulk stores for large ones; 1: always the tile streamer; 2: always the direct-store kernel */
#define DH_OPT_FUSED_MAX_CHUNK 15 /* fused loss scheduler with the tiered tail: upper bound on the coarse tier's chunk in tiles of 256 rows (default 16; the coarse chunk is about half of a CTA's share of the work) */
#define DH_OPT_NMS_FILTER 16 /* mask-matrix NMS: 1 (default) pairs that provably do not overlap (disjoint slab masks) skip the exact predicate; 0: every pair takes it (A/B checks; the bits are identical); 2..64: on, and a warp whose rows keep more than this many of a block's 64 columns walks the columns in step (tuning aid; default 64 = never).  Images where more than 1 candidate in 8 is not a proper box (inverted corners, NaN) skip the filter by themselves */
#define DH_OPT_NMS_CHAIN 17 /* block sweep of the mask-matrix NMS: 0 (default) the greedy chain of a 64-box block is resolved in parallel rounds when there are no per-class caps; 1: always the serial walk (A/B checks; identical keeps) */
int dh_set_option(dh_handle_t h, int option, int value);
/* Synchronous read of the DH_OPT_PHASE_TIMING counters: out8[0..4] = cycles CTA 0 spent in
 * {stage GT + records, candidates + buffer recycle, emit rows, hand-off to TMA, drain}, out8[5] = tiles. */
int dh_read_phase_timing(dh_handle_t h, long long* out8 /*[host] [8]*/);
/* Sticky input-validation bits accumulated by the kernels launched through this handle (synchronous read; `reset`
 * clears them).  DH_STATUS_BAD_SCALE: a CenterNet box was not below the largest box scale (the reference raises
 * ValueError, CenterNet/tf_centernet_resnet_s8.py:306-307); DH_STATUS_BAD_CLASS: a GT class was outside
 * [0, num_classes) -- the reference raises IndexError (FCOS/fcos.py:281-283), the kernels drop the box;
 * DH_STATUS_COMM_TIMEOUT: a peer rank did not arrive at a loss all-reduce within 10 s (the sums are NaN). */
#define DH_STATUS_BAD_SCALE 1
#define DH_STATUS_BAD_CLASS 2
#define DH_STATUS_COMM_TIMEOUT 4
#define DH_STATUS_TRUNCATED 8 /* dh_prepare_labels met an image with more boxes than max_boxes and kept the first max_boxes (the reference keeps all: FCOS/train_fcos.py:131-135) */
int dh_get_status(dh_handle_t h, int32_t* out /*[host] [1]*/, int reset);
/* Profiling aid: while `buf` ([dev], `bytes` long; NULL switches it off) is set, every fused encode+loss launch whose
 * grid fits writes 12 values per CTA b: buf[12b+0] = start, [1] = first chunk staged, [2] = chunk loop done (globaltimer
 * nanoseconds), [3] = chunks processed, [4..9] = time thread 0 spent in {staging GT rows, marking candidate rows, the
 * streaming pass, the barrier behind it, visiting marked rows, the chunk-end reduction}; buf[12*grid] = end of the
 * in-kernel reduction. */
int dh_set_trace(dh_handle_t h, long long* buf /*[dev]*/, long long bytes);
/* Number of kernels this handle has launched since creation (bench.py's `gpu_launches`). */
long long dh_launch_count(dh_handle_t h);
/* Host-only (no device needed): the work split dh_fcos_detect's candidate selection would use for per-level head sizes
 * `values[l]` = rows * (num_classes + 5) floats.  Thread-block-cluster selector: returns the number of clusters per
 * image (<= n_levels); level[t*8 + r] is the level rank r of cluster t works on (-1: idle), lead[t*8 + r] the rank of
 * that group's leader CTA, members[t*8 + r] the group's size.  Streaming pre-select: chunk_first[l] (n_levels + 1
 * entries) is the first CTA of level l when `batch` images are processed with float4 loads (ids run level > image >
 * chunk).  Any output may be NULL.  Returns < 0 on a bad argument. */
int dh_plan_fcos_select(const long long* values, int n_levels, int batch, signed char* level /*[n_levels*8]*/,
                        unsigned char* lead /*[n_levels*8]*/, unsigned char* members /*[n_levels*8]*/,
                        int* chunk_first /*[n_levels+1]*/);

/* Host-only (no device needed): the chunk plan dh_*_encode_loss(_grad) would use for `batch` images of
 * `tiles_per_image` tiles (256-row tiles of `ch` floats per row) on `grid` resident CTAs, with DH_OPT_FUSED_TAIL = tail and
 * DH_OPT_FUSED_MAX_CHUNK = max_chunk.  out[0] = tiers, out[1] = chunks in total, then per tier k (at most 8)
 * out[2 + 4k ..] = {first image, tiles per chunk, chunks per image, id of the tier's first chunk}.  A tier ends where the
 * next begins (the last at `batch`).  Returns the number of tiers, < 0 on a bad argument. */
int dh_plan_fused_chunks(int batch, int tiles_per_image, int rows_per_tile, int ch, int grid, int with_grad, int tail, int max_chunk,
                         int32_t* out /*[host] [34]*/);

/* ---- target encoders ---------------------------------------------------------------------- */

/* FCOS family.  Replaces format_data in FCOS/fcos.py:136-378 (mode 0), FCOS/fcos_center.py:149-279
 * (mode 1: 3x3 window, mode 2: center_only=True) and FCOS/fcos_center_v1.py:149-258 (mode 3).
 * out_levels[l] is [B, Hl, Wl, C+5] with Hl = int(pad_h / strides[l]); fully overwritten.
 * num_targets (optional) is [B, n_levels]: GT boxes assigned to each level.               */
#define DH_FCOS_FOOTPRINT 0
#define DH_FCOS_CENTER3X3 1
#define DH_FCOS_CENTER_ONLY 2
#define DH_FCOS_CENTER_V1 3
#define DH_FCOS_MIN_AREA 4 /* extension: mode 0 with the tie-break the FCOS paper and the reference's own comment (FCOS/fcos.py:185-188) ask for -- where footprints overlap the SMALLEST box supplies channels 0..4 (the reference's ascending-area paint order lets the largest win, :202-209); equal areas: the higher index.  Specified by oracle.fcos_format_data(order="min_area") */
int dh_fcos_encode(dh_handle_t h,
                   const float* boxes /*[dev] [B,max_boxes,5]*/, const int32_t* nbox /*[dev] [B]*/,
                   const float* img_dim /*[dev] [B,2]*/,
                   int batch, int max_boxes, int pad_h, int pad_w,
                   int n_levels, const int32_t* strides /*[host] [n_levels]*/,
                   const float* b_dim /*[host] [n_levels-1]*/,
                   int num_classes, int mode,
                   float* const* out_levels /*[host] n_levels [dev] pointers*/,
                   int32_t* num_targets /*[dev] [B,n_levels] or NULL*/,
                   void* stream);

/* RetinaNet matcher + box encoder.  Replaces RetinaNet.format_data (RetinaNet/retinanet_module.py:
 * 251-365) together with get_anchors (:221-246) and compute_iou (RetinaNet/utils.py:42-83); anchors
 * are generated analytically from anchor_hw.  out_levels[l] is [B, A, Hl, Wl, C+4] (slice [:, a] is
 * the reference's all_outputs[l][a]).  num_pairs (optional) is [B]: positive (gt, anchor) pairs. */
int dh_retina_encode(dh_handle_t h,
                     const float* boxes, const int32_t* nbox, const float* img_dim,
                     int batch, int max_boxes, int pad_h, int pad_w,
                     int n_levels, const int32_t* strides /*[host]*/,
                     int n_anchors, const float* anchor_hw /*[host] [n_levels,n_anchors,2] (h,w)*/,
                     float iou_thresh, int num_classes,
                     float* const* out_levels /*[host] n_levels [dev] pointers*/,
                     int32_t* num_pairs /*[dev] [B] or NULL*/,
                     void* stream);

/* CenterNet encoders.  mode 0: tf_centernet_resnet_s8.format_data (CenterNet/tf_centernet_resnet_s8.py:
 * 243-330), out [B,H,W,S,C+4]; mode 1: tf_centernet_hourglass.format_data (CenterNet/
 * tf_centernet_hourglass.py:379-456), out [B,H,W,C+4]; mode 2: tf_centernet.format_data
 * (CenterNet/tf_centernet.py:152-342, inverse-power fall-off heat), out [B,H,W,C+5]; mode 3: the 4-scale encoder that
 * CenterNet/train_hourglass_voc.py:99-153 keeps inline in train(): out [B,H,W,4,C+5] = (h_off, w_off, h_reg, w_reg,
 * objectness, classes), scales = pad0 / (8, 4, 2, 1), image offset int((pad - img_dim) / 2) (box_scales is ignored).
 * pad0/pad1 are img_pad[0]/img_pad[1] exactly as the reference indexes them (modes 0 and 1 swap
 * them, tf_centernet_resnet_s8.py:259-262).  status (optional, [dev] int32, zeroed by the call)
 * gets bit 0 set when a box is not below the largest box scale (the reference raises ValueError). */
#define DH_CENTERNET_ONEHOT_SCALES 0
#define DH_CENTERNET_HOURGLASS 1
#define DH_CENTERNET_POWER_FALLOFF 2
#define DH_CENTERNET_HOURGLASS4 3
#define DH_CENTERNET_GAUSSIAN 4 /* extension: mode 2 with a Gaussian heat channel -- the reference's commented-out gaussian_dist_2d (CenterNet/tf_centernet.py:30-40) with std = max(1, sqrt(box area in cells)) (:203-205, before its override to 8.0), normalised by its maximum over the footprint, centre cell 1; where footprints overlap the heat is the MAXIMUM over the boxes (canonical CenterNet splat, order-free), channels 0..3 and the classes as in mode 2.  Specified by oracle.centernet_gaussian_format_data */
int dh_centernet_encode(dh_handle_t h,
                        const float* boxes, const int32_t* nbox, const float* img_dim,
                        int batch, int max_boxes, int pad0, int pad1, int stride,
                        int n_scales, const float* box_scales /*[host] [n_scales], mode 0 only*/,
                        float sigma, int num_classes, int mode,
                        float* out /*[dev]*/, int32_t* status /*[dev] or NULL*/,
                        void* stream);

/* ---- label preparation / result formatting (the steps either side of the path) ------------------------------- */

/* Element-wise box transforms on [n, 4] float32 rows: swap_xy, convert_to_xywh, convert_to_corners
 * (FCOS/utils.py:6-40, identical in RetinaNet/ and CenterNet/) and the box half of random_flip_horizontal
 * (FCOS/data_preprocess.py:36-39: (xmin, ymin, xmax, ymax) -> (1 - xmax, ymin, 1 - xmin, ymax)).  In place is allowed. */
#define DH_BOX_SWAP_XY 0
#define DH_BOX_TO_XYWH 1
#define DH_BOX_TO_CORNERS 2
#define DH_BOX_FLIP_HORIZONTAL 3
int dh_box_convert(dh_handle_t h, const float* boxes /*[dev] [n,4]*/, long long n, int mode, float* out /*[dev] [n,4]*/, void* stream);

/* Dataset boxes (xmin, ymin, xmax, ymax, normalised; FCOS/format_VOC_fcos.py:60-68) + class ids -> the padded
 * [B, max_boxes, 5] (cy, cx, h, w, class) + nbox layout the encoders take: per-image optional horizontal flip,
 * swap_xy, convert_to_xywh, concat with the class (FCOS/data_preprocess.py:121-131, FCOS/train_fcos.py:131-135).
 * Input is ragged when box_offsets ([B+1], rows of image b are [off[b], off[b+1])) is given, else padded
 * [B, in_max_boxes, 4] with optional nbox.  Images with more than max_boxes rows are truncated to the first max_boxes:
 * out_nbox holds the kept count and DH_STATUS_TRUNCATED is set (dh_get_status). */
int dh_prepare_labels(dh_handle_t h, const float* raw_boxes /*[dev]*/, const float* classes /*[dev] float32, indexed like raw_boxes*/,
                      const int32_t* box_offsets /*[dev] [B+1] or NULL*/, const int32_t* nbox /*[dev] [B] or NULL*/,
                      const int32_t* flip /*[dev] [B] or NULL*/, int batch, int in_max_boxes, int max_boxes,
                      float* out_labels /*[dev] [B,max_boxes,5]*/, int32_t* out_nbox /*[dev] [B] or NULL*/, void* stream);

/* RetinaNet.detect_bboxes after the NMS (RetinaNet/retinanet_module.py:559-569): kept rows (y1, x1, y2, x2, score,
 * label) -> boxes (x1, y1, x2, y2) rescaled by ratios[b] = (w_ratio, h_ratio) exactly as the reference multiplies
 * them (columns 0 and 2 by w_ratio, 1 and 3 by h_ratio, in float64), scores, integer labels (the caller maps them to
 * names).  Slots >= n_keep[b] are zero / label -1. */
int dh_format_detections(dh_handle_t h, const float* rows /*[dev] [B,n,6]*/, const int32_t* n_keep /*[dev] [B]*/,
                         const float* ratios /*[dev] [B,2]*/, int batch, int n, float* out_boxes /*[dev] [B,n,4]*/,
                         float* out_scores /*[dev] [B,n]*/, int32_t* out_labels /*[dev] [B,n]*/, void* stream);

/* The offline COCO -> sparse FCOS target formatter, /format_COCO_annotations_fcos.py:66-183 (one call = a batch of
 * images of the script's per-image loop).  boxes rows are (x_lower, y_lower, box_width, box_height, label) in SOURCE
 * pixels, float64 like the script's table (the truncations decide indices), label = the script's 1-based class index;
 * src_dims[b] = the source image's (img_width, img_height); (img_width, img_height) = the canvas (the script: 448, 448),
 * scale limits int(min(canvas) / 2^x) (:49-56).  Per object, per footprint cell (objects in input order, cells x-major
 * as np.nonzero walks them) seven COO entries: indices [y, x, scale, k] with values (b, t, l, r, centerness, 1) for
 * k = 0..5 and [y, x, scale, label + 4] = 1 (the script writes this one five long, :171 -- the duplicated scale is
 * dropped).  Values are float32 (centerness computed in float64, :8-11).  Overlapping objects all emit (`tmp_mask` is
 * never written, :91).
 * out_offsets [B+1] (int64): entries of image b are [out_offsets[b], out_offsets[b+1]).  With out_indices ==
 * out_values == NULL only the offsets are computed (size the buffers from out_offsets[B], then call again); entries
 * at or beyond `capacity` are not written. */
int dh_fcos_sparse_encode(dh_handle_t h, const double* boxes /*[dev] [B,max_boxes,5]*/, const int32_t* nbox /*[dev] [B] or NULL*/,
                          const double* src_dims /*[dev] [B,2]*/, int batch, int max_boxes, int img_width, int img_height,
                          int num_scale, long long capacity, int32_t* out_indices /*[dev] [capacity,4] or NULL*/,
                          float* out_values /*[dev] [capacity] or NULL*/, long long* out_offsets /*[dev] [B+1]*/, void* stream);

/* The tail of show_heatmap (FCOS/train_fcos_center_voc.py:92-121): combined-NMS boxes [B,T,4] (y1, x1, y2, x2) of
 * target maps decoded by dh_fcos_detect(center = DH_FCOS_SCORE_MAP[_CEN]) -> the rectangles it draws, (x1, y1, w, h):
 * the columns are multiplied by ratios[b] = (w_ratio, h_ratio) in the reference's order (0 and 2 by w_ratio, 1 and 3
 * by h_ratio, float32 like the TF multiply), swapped to (x1, y1, x2, y2), x1 / y1 <= 0 become 0, w = x2 - x1,
 * h = y2 - y1.  Slots >= valid[b] are zero. */
int dh_fcos_rectangles(dh_handle_t h, const float* boxes /*[dev] [B,T,4]*/, const int32_t* valid /*[dev] [B]*/,
                       const float* ratios /*[dev] [B,2]*/, int batch, int t, float* out_rect /*[dev] [B,T,4]*/, void* stream);

/* ---- losses --------------------------------------------------------------------------------- */

/* Channel layout of a row: [0, reg_ch) box regression, then one centerness channel if cen_mode != 0,
 * then the class channels.  All sums are over every element (the reference never normalises).
 *   cen_mode  DH_CEN_NONE      no centerness channel (RetinaNet, CenterNet s8 / hourglass)
 *             DH_CEN_SMOOTH_L1 smooth-L1(target, sigmoid(pred)) over ALL rows (FCOS/fcos.py:483-486)
 *             DH_CEN_FOCAL     focal(target, pred) (FCOS/fcos_center.py:387-389, fcos_center_v1.py:308-310)
 *             DH_CEN_IGNORE    channel present, contributes 0 (fcos.py model_loss with cen_type != "l1")
 *   reg_mode  DH_REG_SMOOTH_L1 where(|d| < delta, d^2/2, |d|) over positive rows (FCOS/fcos.py:380-391)
 *             DH_REG_IOU       -log(IoU + 1e-12) on the integer grid over positive rows (FCOS/fcos.py:393-441)
 *             DH_REG_GIOU      1 - GIoU on the same box construction (extension: the reference has no GIoU, SURVEY.md
 *                              section 0; oracle.giou_loss states it), GIoU = IoU - (C - union) / (C + 1e-12)
 *   cls_mode  DH_CLS_FOCAL (alpha, gamma) or DH_CLS_SIGMOID_BCE (CenterNet/tf_hourglass_net.py:347-349, :381-382)
 *   pos_rule  DH_POS_GE1 max(class) >= 1 (fcos.py:475-477), DH_POS_GT0 max(class) > 0
 *             (retinanet_module.py:416-418), DH_POS_MASK caller-supplied per-row float mask.
 * Outputs are float32 {cls, reg, cen, n_pos}: out_per_image [B,4] and/or out_total [4] (either may be
 * NULL, not both).  Summation order is fixed, so results are run-to-run deterministic.            */
#define DH_CLS_FOCAL 0
#define DH_CLS_SIGMOID_BCE 1 /* tf.nn.sigmoid_cross_entropy_with_logits summed (CenterNet/tf_hourglass_net.py:347-349) */
#define DH_CEN_NONE 0
#define DH_CEN_SMOOTH_L1 1
#define DH_CEN_FOCAL 2
#define DH_CEN_IGNORE 3
#define DH_REG_SMOOTH_L1 0
#define DH_REG_IOU 1
#define DH_REG_GIOU 2
#define DH_POS_GE1 0
#define DH_POS_GT0 1
#define DH_POS_MASK 2

/* Loss over materialised targets: replaces focal_loss / smooth_l1_loss / iou_loss / model_loss
 * (FCOS/fcos.py:380-496 and the identical copies in fcos_center*.py, retinanet_module.py:367-426,
 * tf_centernet*.py).  Map m holds [B, H_m*W_m*sub_m, ch] rows, targets and predictions alike.     */
int dh_dense_loss(dh_handle_t h, int n_maps,
                  const float* const* target_maps /*[host] n_maps [dev] ptrs*/,
                  const float* const* pred_maps /*[host] n_maps [dev] ptrs*/,
                  const float* const* mask_maps /*[host] n_maps [dev] ptrs [B,rows], DH_POS_MASK only, else NULL*/,
                  const int32_t* map_height /*[host]*/, const int32_t* map_width /*[host]*/,
                  const int32_t* map_sub /*[host] rows per cell, or NULL (=1)*/,
                  int batch, int ch, int reg_ch, int cen_mode, int reg_mode, int pos_rule, int cls_mode,
                  float alpha, float gamma, float delta,
                  float* out_per_image /*[dev] [B,4] or NULL*/, float* out_total /*[dev] [4] or NULL*/,
                  void* stream);

/* Fused encode + loss: targets are generated in shared memory and consumed there; HBM traffic is one
 * read of the predictions.  Same target semantics as the matching dh_*_encode call, same loss
 * semantics as dh_dense_loss.  pred_levels[l] has the layout of the corresponding out_levels[l].   */
int dh_fcos_encode_loss(dh_handle_t h,
                        const float* boxes, const int32_t* nbox, const float* img_dim,
                        int batch, int max_boxes, int pad_h, int pad_w,
                        int n_levels, const int32_t* strides, const float* b_dim,
                        int num_classes, int mode,
                        const float* const* pred_levels /*[host] n_levels [dev] ptrs [B,Hl,Wl,C+5]*/,
                        int reg_mode, int cen_mode, float alpha, float gamma, float delta,
                        float* out_per_image, float* out_total, int32_t* num_targets, void* stream);
int dh_retina_encode_loss(dh_handle_t h,
                          const float* boxes, const int32_t* nbox, const float* img_dim,
                          int batch, int max_boxes, int pad_h, int pad_w,
                          int n_levels, const int32_t* strides, int n_anchors, const float* anchor_hw,
                          float iou_thresh, int num_classes,
                          const float* const* pred_levels /*[host] n_levels [dev] ptrs [B,A,Hl,Wl,C+4]*/,
                          float alpha, float gamma, float delta,
                          float* out_per_image, float* out_total, int32_t* num_pairs, void* stream);
int dh_centernet_encode_loss(dh_handle_t h,
                             const float* boxes, const int32_t* nbox, const float* img_dim,
                             int batch, int max_boxes, int pad0, int pad1, int stride,
                             int n_scales, const float* box_scales, float sigma, int num_classes, int mode,
                             const float* pred /*[dev] same layout as dh_centernet_encode's out*/,
                             int reg_mode /*mode 2 only*/, int cls_mode /*DH_CLS_*; mode 3 counts objectness + classes as the class channels, boxes = smooth-L1 with delta (0 = plain L1, CenterNet/tf_hourglass_net.py:386-387)*/,
                             float alpha, float gamma, float delta,
                             float* out_per_image, float* out_total, int32_t* status, void* stream);

/* ---- losses with gradients ------------------------------------------------------------
 * The reference differentiates through model_loss with tf.GradientTape (FCOS/train_fcos.py:152-176,
 * RetinaNet/train_retinanet_coco.py:207-218, CenterNet/tf_centernet_resnet_s8.py:417-432).  Each *_grad entry point
 * computes the same loss sums as its forward twin AND, in the same pass over the predictions, writes
 *     grad = d (w_cls * cls + w_reg * reg + w_cen * cen) / d pred
 * into caller-owned maps with the layout of the predictions (every element is written; one read of the predictions,
 * one write of the gradient).  out_per_image / out_total may both be NULL when only the gradient is wanted.     */
int dh_dense_loss_grad(dh_handle_t h, int n_maps, const float* const* target_maps, const float* const* pred_maps,
                       const float* const* mask_maps, const int32_t* map_height, const int32_t* map_width,
                       const int32_t* map_sub, int batch, int ch, int reg_ch, int cen_mode, int reg_mode, int pos_rule, int cls_mode,
                       float alpha, float gamma, float delta, float w_cls, float w_reg, float w_cen,
                       float* const* grad_maps /*[host] n_maps [dev] ptrs*/, float* out_per_image, float* out_total,
                       void* stream);
int dh_fcos_encode_loss_grad(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                             int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, const float* b_dim,
                             int num_classes, int mode, const float* const* pred_levels, int reg_mode, int cen_mode,
                             float alpha, float gamma, float delta, float w_cls, float w_reg, float w_cen,
                             float* const* grad_levels /*[host] n_levels [dev] ptrs*/, float* out_per_image,
                             float* out_total, int32_t* num_targets, void* stream);
int dh_retina_encode_loss_grad(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                               int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, int n_anchors,
                               const float* anchor_hw, float iou_thresh, int num_classes, const float* const* pred_levels,
                               float alpha, float gamma, float delta, float w_cls, float w_reg,
                               float* const* grad_levels /*[host] n_levels [dev] ptrs*/, float* out_per_image,
                               float* out_total, int32_t* num_pairs, void* stream);
int dh_centernet_encode_loss_grad(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                                  int max_boxes, int pad0, int pad1, int stride, int n_scales, const float* box_scales,
                                  float sigma, int num_classes, int mode, const float* pred, int reg_mode, int cls_mode, float alpha,
                                  float gamma, float delta, float w_cls, float w_reg, float w_cen, float* grad /*[dev]*/,
                                  float* out_per_image, float* out_total, int32_t* status, void* stream);

/* ---- multi-GPU: the loss-scalar exchange -----------------------------------------
 * The reference has no distribution strategy at all; the batch is sharded by image, one rank
 * per GPU, and the only collective of the path is the sum of {cls, reg, cen, n_pos} -- what the reference's training
 * loop accumulates over the images of a batch (FCOS/train_fcos.py:167-194).  Two transports:
 *   peer mailboxes  every rank's 2 KB mailbox is mapped into its peers (cudaIpc between processes, peer access
 *                   inside one process); a rank stores its values into every peer's mailbox over NVLink and sums what
 *                   arrives in its own, in rank order (bit-identical on all ranks, deterministic).  ~2 us; can run
 *                   inside the fused loss kernel itself (DH_OPT_LOSS_ALLREDUCE).
 *   NCCL            ncclAllReduce on the caller's stream (libnccl.so.2 is dlopen'ed on first use).
 * Both are stream-ordered and capturable in a CUDA graph.  Setup, one process per GPU: rank 0 calls
 * dh_comm_get_unique_id and every rank dh_comm_init_rank with those 128 bytes (NCCL); every rank calls
 * dh_comm_peer_export, the 64-byte blobs are all-gathered out of band (MPI, a file, torch.distributed ...) and handed
 * to dh_comm_peer_import (mailboxes).  One process driving several GPUs: dh_comm_init_all sets up both.
 * Every rank must issue the same sequence of all-reduces.  dh_comm_destroy (also run by dh_destroy) must only be
 * called once no peer can still be inside an exchange (barrier first). */
#define DH_UNIQUE_ID_BYTES 128
#define DH_IPC_HANDLE_BYTES 64
#define DH_COMM_MAX_RANKS 16
int dh_comm_get_unique_id(void* out128 /*[host] [DH_UNIQUE_ID_BYTES]*/);
int dh_comm_init_rank(dh_handle_t h, int world, int rank, const void* unique_id128 /*[host]*/);
int dh_comm_init_all(dh_handle_t* handles /*[host] [ndev]*/, int ndev);
int dh_comm_peer_export(dh_handle_t h, void* out64 /*[host] [DH_IPC_HANDLE_BYTES]*/);
int dh_comm_peer_import(dh_handle_t h, int world, int rank, const void* handles /*[host] [world][DH_IPC_HANDLE_BYTES], rank-major*/);
/* world / rank of the handle's communicator (1 / 0 without one); transports: bit 0 NCCL, bit 1 peer mailboxes. */
int dh_comm_info(dh_handle_t h, int* world, int* rank, int* transports);
/* In-place sum of scalars[0..count) (count <= 12) over all ranks; no-op for a single rank. */
int dh_allreduce_loss(dh_handle_t h, float* scalars /*[dev] [count]*/, int count, void* stream);
int dh_comm_destroy(dh_handle_t h);

/* ---- inference: decode, candidate selection, NMS ------------------------------------------------- */

/* prediction_to_corners: head regression values -> pixel corner boxes (y1, x1, y2, x2), float32.
 * pred holds rows of ch_in floats (the first 4 are used) laid out [B, H, W, sub]; out is [B, H, W, sub, 4].
 *   mode 0  FCOS tblr, cell centres at i+.5, times stride         (FCOS/fcos.py:112-134; same in
 *           fcos_center.py:125-147, tf_centernet.py:128-150, tf_centernet_hourglass.py:355-377)
 *   mode 1  RetinaNet: centre = i*stride - p*anchor, size = p*anchor; d0, d1 = anchor (h, w)
 *           (RetinaNet/retinanet_module.py:428-451)
 *   mode 2  fcos_center_v1: centre = (i + p)*stride, size = p*box_sc; d0 = box_sc (FCOS/fcos_center_v1.py:125-147)
 *   mode 3  CenterNet s8: like mode 2 with one scale per `sub` index (CenterNet/tf_centernet_resnet_s8.py:210-241) */
int dh_prediction_to_corners(dh_handle_t h, const float* pred /*[dev]*/, int batch, int height, int width, int sub,
                             int ch_in, int mode, float stride, float d0, float d1,
                             const float* scales /*[host] [sub], mode 3*/, float* out /*[dev]*/, void* stream);

/* How dh_fcos_decode / dh_fcos_detect score a (location, class) pair (their `center` argument).
 *   DH_FCOS_SCORE_CLS      sigmoid(class logit)                                      FCOS/infer_fcos.py:50-51
 *   DH_FCOS_SCORE_CLS_CEN  sigmoid(centerness logit) * sigmoid(class logit)          FCOS/infer_fcos.py:44-48
 *   DH_FCOS_SCORE_MAP      the class channel as it is: the input is a TARGET map ([B,Hl,Wl,C+5] as dh_fcos_encode writes
 *                          it) sent back through the detector -- show_heatmap(center=False),
 *                          FCOS/train_fcos_center_voc.py:63-66
 *   DH_FCOS_SCORE_MAP_CEN  sqrt(class channel * centerness channel), product and root in float64 like NumPy, rounded to
 *                          float32 where combined NMS takes it -- show_heatmap(center=True), :58-62 */
#define DH_FCOS_SCORE_CLS 0
#define DH_FCOS_SCORE_CLS_CEN 1
#define DH_FCOS_SCORE_MAP 2
#define DH_FCOS_SCORE_MAP_CEN 3

/* FCOS decode front end (FCOS/infer_fcos.py:35-57): per-level heads [B,Hl,Wl,C+5] -> boxes [B,N,4] and
 * scores [B,N,C] (DH_FCOS_SCORE_* mode `center`), levels concatenated. */
int dh_fcos_decode(dh_handle_t h, const float* const* pred_levels /*[host] n_levels [dev] ptrs*/, int batch,
                   int pad_h, int pad_w, int n_levels, const int32_t* strides /*[host]*/, int num_classes, int center,
                   float* boxes /*[dev] [B,N,4]*/, float* scores /*[dev] [B,N,C]*/, void* stream);

/* RetinaNet decode front end (RetinaNet/retinanet_module.py:487-520): per-level heads [B,A,Hl,Wl,C+4] ->
 * dets [B,N,6] = (y1, x1, y2, x2, max_c sigmoid, first argmax), order level > anchor > row-major cell. */
int dh_retina_decode(dh_handle_t h, const float* const* pred_levels /*[host] n_levels [dev] ptrs*/, int batch,
                     int pad_h, int pad_w, int n_levels, const int32_t* strides /*[host]*/, int n_anchors,
                     const float* anchor_hw /*[dev] [n_levels,n_anchors,2]*/, int num_classes,
                     float* dets /*[dev] [B,N,6]*/, void* stream);

/* Pre-NMS candidate selection: for every image and every segment [seg_off[s], seg_off[s+1]) of its
 * n_total rows (a segment = a pyramid level), keep the rows whose score (column score_col) passes min_score and is
 * among the k highest of the segment (exact; ties go to the lower index), in index order.
 * out is [B, n_seg*k, row_floats]; unused slots get score = -inf.  out_src (optional) [B, n_seg*k]: source row.
 * The reference has no top-k; with k >= segment length this is its plain threshold. */
int dh_select_topk(dh_handle_t h, const float* dets /*[dev] [B,n_total,row_floats]*/, int batch, long long n_total,
                   int row_floats, int score_col, const int32_t* seg_off /*[dev] [n_seg+1]*/, int n_seg, int k, float min_score,
                   int score_inclusive, float* out /*[dev]*/, int32_t* out_src /*[dev] or NULL*/, void* stream);

/* Greedy NMS.  dets is [B, n_max, row_floats] float32 rows (c0, c1, c2, c3, score[, class]); n_valid (optional)
 * [B] limits the rows per image; rows failing the score test (>= min_score when score_inclusive, else >) are dropped.
 *   DH_NMS_AGNOSTIC   RetinaNet.cpu_nms (RetinaNet/retinanet_module.py:453-481): class-agnostic,
 *                     ovr = inter / (area_i + area_j - inter + 1e-8), a box survives while ovr <= iou_thr.
 *   DH_NMS_PER_CLASS  the per-class greedy NMS with caps that FCOS/infer_fcos.py:58-61 gets from TensorFlow's
 *                     combined_non_max_suppression: same-class boxes with IoU > iou_thr are suppressed, at most
 *                     max_per_class kept per class and max_total overall (0 = no cap).  Parity unpinned (the op's
 *                     source is not part of the reference); restated in oracle/dense_head_ref.py.
 * keep is [B, max_out] original row indices in descending score order (ties: lower index first, where the
 * reference's unstable argsort leaves the order unspecified); n_keep is [B].  n_max <= 16384. */
#define DH_NMS_AGNOSTIC 0
#define DH_NMS_PER_CLASS 1
int dh_nms(dh_handle_t h, const float* dets /*[dev]*/, const int32_t* n_valid /*[dev] [B] or NULL*/, int batch,
           int n_max, int row_floats, int mode, float iou_thr, float min_score, int score_inclusive,
           int num_classes, int max_per_class, int max_total,
           int32_t* keep /*[dev] [B,max_out]*/, int max_out, int32_t* n_keep /*[dev] [B]*/, void* stream);

/* Whole-image pipelines: head outputs in, final detections out, in one call (all intermediates live in the handle's
 * scratch; nothing synchronises with the host).
 *
 * dh_fcos_detect = image_detections of FCOS/infer_fcos.py:27-62.  Scores follow the DH_FCOS_SCORE_* mode `center`
 * (with DH_FCOS_SCORE_MAP[_CEN], iou_thr = cls_thr = 0.75 and target maps as input it is the round trip of show_heatmap,
 * FCOS/train_fcos_center_voc.py:54-90); per level the pre_nms_topk best (location, class) pairs with score > cls_thr
 * go to the per-class NMS with caps (see DH_NMS_PER_CLASS).  Outputs have the layout of
 * tf.image.combined_non_max_suppression: boxes [B,T,4], scores [B,T], classes [B,T] zero padded, valid [B], with
 * T = max_total.  pre_nms_topk >= the number of passing pairs reproduces the reference (which has no top-k).
 * out_cand (optional) receives the candidate rows [B, n_levels*k, 6], k = min(pre_nms_topk, longest level). */
int dh_fcos_detect(dh_handle_t h, const float* const* pred_levels /*[host] n_levels [dev] ptrs [B,Hl,Wl,C+5]*/, int batch,
                   int pad_h, int pad_w, int n_levels, const int32_t* strides /*[host]*/, int num_classes, int center,
                   float iou_thr, float cls_thr, int max_per_class, int max_total, int pre_nms_topk,
                   float* out_boxes /*[dev] [B,T,4]*/, float* out_scores /*[dev] [B,T]*/, float* out_classes /*[dev] [B,T]*/,
                   int32_t* out_valid /*[dev] [B]*/, float* out_cand /*[dev] or NULL*/, void* stream);

/* dh_retina_detect = RetinaNet.image_detections of RetinaNet/retinanet_module.py:483-530 for a batch: rows
 * (y1, x1, y2, x2, score, label) in kept (score-descending) order, out_rows [B, max_out, 6] zero padded, out_n [B].
 * pre_nms_topk > 0 keeps that many candidates per level; 0 = threshold only like the reference, limited to
 * 16384 / n_levels candidates per level -- out_overflow [B] (optional) gets 1 where an image had more.
 * out_cand [B, n_levels*k, 6] / out_keep [B, max_out] (optional) receive the candidates and the kept indices. */
int dh_retina_detect(dh_handle_t h, const float* const* pred_levels /*[host] n_levels [dev] ptrs [B,A,Hl,Wl,C+4]*/, int batch,
                     int pad_h, int pad_w, int n_levels, const int32_t* strides /*[host]*/, int n_anchors,
                     const float* anchor_hw /*[dev] [n_levels,n_anchors,2]*/, int num_classes, float iou_thr, float cls_thr,
                     int pre_nms_topk, float* out_rows /*[dev]*/, int max_out, int32_t* out_n /*[dev] [B]*/,
                     int32_t* out_overflow /*[dev] [B] or NULL*/, float* out_cand /*[dev] or NULL*/, int32_t* out_keep /*[dev] or NULL*/,
                     void* stream);

/* Pairwise IoU of centre-size boxes (c0, c1, size0, size1), float32: compute_iou of RetinaNet/utils.py:42-83
 * (union floored at 1e-8, result clipped to [0, 1]).  out is [n, m]. */
int dh_compute_iou(dh_handle_t h, const float* boxes1 /*[dev] [n,4]*/, int n, const float* boxes2 /*[dev] [m,4]*/, int m,
                   float* out /*[dev] [n,m]*/, void* stream);

/* float64 IoU of corner boxes floored at float32 eps: bboxes_iou of CenterNet/tf_centernet_resnet_s8.py:22-42
 * (also tf_centernet_hourglass.py).  n1 and n2 must be equal, or one of them 1 (broadcast); out is [max(n1,n2)]. */
int dh_bboxes_iou(dh_handle_t h, const double* boxes1 /*[dev] [n1,4]*/, int n1, const double* boxes2 /*[dev] [n2,4]*/,
                  int n2, double* out /*[dev]*/, void* stream);

/* CenterNet per-class (soft-)NMS: nms of CenterNet/tf_centernet_resnet_s8.py:44-85.  rows [n,6] float64
 * (xmin, ymin, w, h, score, class); classes [n_classes] float64 = the distinct class values in the order to
 * visit them (the reference iterates a Python set; ascending order is the canonical choice).  Kept rows come
 * back as (x1, y1, x2, y2, score, class) in out_rows [n,6] with their source row in out_src [n]; n_out [1].
 * soft != 0 selects method='soft-nms' with `sigma`.  The input is not modified (the reference mutates it). */
int dh_centernet_nms(dh_handle_t h, const double* rows /*[dev]*/, int n, const double* classes /*[dev]*/, int n_classes,
                     double iou_threshold, double sigma, int soft, double* out_rows /*[dev] [n,6]*/,
                     int32_t* out_src /*[dev] [n]*/, int32_t* n_out /*[dev] [1]*/, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DENSEHEAD_H_ */
