// Host helpers shared by the launchers in encode.cu / loss.cu.
#pragma once
#include "dh_policies.cuh"

namespace dh {
// Fill tile_begin / n_tiles / fast divisors for tt.maps[0..n_maps) and pick rows_per_tile so that one
// tile is about `tile_bytes`.  Returns the shared-memory bytes one stage buffer needs.
int finish_table(TileTable& tt, int ch, int batch, int tile_bytes);
// Tile size for a problem: `max_tile_bytes` for large outputs, smaller when the output would not fill the SMs.
int auto_tile_bytes(const TileTable& tt, int ch, int batch, int max_tile_bytes, int sm_count);

// Validate the detector configuration and fill the policy parameters + map geometry.  `out_*` are the
// target tensors (encode), `pred_*` the prediction tensors (fused loss); either may be null.
int fill_fcos(FcosPolicy::Params& p, TileTable& tt, int pad_h, int pad_w, int n_levels, const int32_t* strides,
              const float* b_dim, int num_classes, int mode, float* const* out_levels,
              const float* const* pred_levels, int32_t* num_targets, const char* who);
int fill_retina(RetinaPolicy::Params& p, TileTable& tt, int pad_h, int pad_w, int n_levels, const int32_t* strides,
                int n_anchors, const float* anchor_hw, float iou_thresh, int num_classes, float* const* out_levels,
                const float* const* pred_levels, int32_t* num_pairs, const char* who);
int fill_centernet(CenterNetPolicy::Params& p, TileTable& tt, int pad0, int pad1, int stride, int n_scales,
                   const float* box_scales, float sigma, int num_classes, int mode, float* out, const float* pred,
                   int32_t* status, const char* who);
}  // namespace dh
