// Internal launch helpers shared by decode.cu (which owns the kernels) and detect.cu (the pipelines).
#pragma once
#include "dh_host.h"

namespace dh {
// Fused FCOS decode + per-level top-k: candidate rows (y1, x1, y2, x2, score, class) -> cand [B, n_levels*k, 6].
int launch_fcos_select(dh_handle_s* h, const float* const* pred_levels, int batch, int pad_h, int pad_w, int n_levels,
                       const int32_t* strides, int num_classes, int center, int k, float min_score, int inclusive, float* cand,
                       cudaStream_t st);
// dh_select_topk with the segment offsets taken from the host; overflow[b] (optional) gets bit 0 set when a
// segment of image b had more than k rows above the threshold.
int launch_select_segs(dh_handle_s* h, const float* dets, int batch, long long n_total, int row_floats, int score_col, const int* seg_off_host,
                       int n_seg, int k, float min_score, int inclusive, float* out, int* overflow, cudaStream_t st,
                       const float* scores = nullptr /*[B, n_total]: the score column on its own, when the producer wrote one*/);
// dh_retina_decode that can also write the score column alone ([B, N]; *wrote_scores = 1 when it did)
int retina_decode_scores(dh_handle_s* h, const float* const* pred_levels, int batch, int pad_h, int pad_w, int n_levels,
                         const int32_t* strides, int n_anchors, const float* anchor_hw_dev, int num_classes, float* dets, float* scores,
                         int* wrote_scores, void* stream);
}  // namespace dh
