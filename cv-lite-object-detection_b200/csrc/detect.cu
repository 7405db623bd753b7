// Whole-image detection pipelines behind one C call each: head outputs in, final detections out, every
// intermediate in the handle's scratch and no host synchronisation.
//
//   dh_fcos_detect    FCOS/infer_fcos.py:27-62  image_detections: decode + sigmoid [* centerness] ->
//                     per-level top-k of the (location, class) scores above the threshold -> per-class greedy NMS
//                     with the combined-NMS caps -> zero-padded (boxes, scores, classes, valid_detections)
//   dh_retina_detect  RetinaNet/retinanet_module.py:483-530  image_detections: decode, max / first-argmax over
//                     classes -> score >= threshold [-> per-level top-k] -> class-agnostic greedy NMS (cpu_nms)
//                     -> rows (y1, x1, y2, x2, score, label) in kept order
#include <cstring>

#include "dh_common.cuh"
#include "dh_infer.h"

namespace dh {

// keep [B, max_out] indices into cand [B, n_cand, 6] -> rows [B, max_out, 6]; slots >= n_keep[b] are zeroed
__global__ void gather_kept_rows_kernel(const float* __restrict__ cand, int n_cand, const int* __restrict__ keep, const int* __restrict__ n_keep,
                                        int max_out, float* __restrict__ rows) {
    const int b = blockIdx.y;
    const int nk = min(n_keep[b], max_out);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < max_out * 6; e += gridDim.x * blockDim.x) {
        const int r = e / 6, c = e - r * 6;
        float v = 0.f;
        if (r < nk) v = cand[(static_cast<long long>(b) * n_cand + keep[static_cast<long long>(b) * max_out + r]) * 6 + c];
        rows[(static_cast<long long>(b) * max_out + r) * 6 + c] = v;
    }
}
// the combined-NMS output layout: boxes [B,T,4], scores [B,T], classes [B,T], zero padded
__global__ void gather_fcos_outputs_kernel(const float* __restrict__ cand, int n_cand, const int* __restrict__ keep, const int* __restrict__ n_keep,
                                           int max_out, float* __restrict__ boxes, float* __restrict__ scores, float* __restrict__ classes) {
    const int b = blockIdx.y;
    const int nk = min(n_keep[b], max_out);
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < max_out; r += gridDim.x * blockDim.x) {
        float row[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < nk) {
            const float* src = cand + (static_cast<long long>(b) * n_cand + keep[static_cast<long long>(b) * max_out + r]) * 6;
#pragma unroll
            for (int c = 0; c < 6; ++c) row[c] = src[c];
        }
        const long long o = static_cast<long long>(b) * max_out + r;
        boxes[o * 4 + 0] = row[0], boxes[o * 4 + 1] = row[1], boxes[o * 4 + 2] = row[2], boxes[o * 4 + 3] = row[3];
        scores[o] = row[4], classes[o] = row[5];
    }
}

}  // namespace dh

using namespace dh;

extern "C" {

int dh_fcos_detect(dh_handle_t h, const float* const* pred_levels, int batch, int pad_h, int pad_w, int n_levels,
                   const int32_t* strides, int num_classes, int center, float iou_thr, float cls_thr, int max_per_class,
                   int max_total, int pre_nms_topk, float* out_boxes, float* out_scores, float* out_classes, int32_t* out_valid,
                   float* out_cand, void* stream) {
    DH_CHECK_ARG(h && pred_levels && strides && out_boxes && out_scores && out_classes && out_valid, "dh_fcos_detect: NULL argument");
    DH_CHECK_ARG(n_levels >= 1 && n_levels <= DH_MAX_PYRAMID_LEVELS && num_classes >= 1, "dh_fcos_detect: bad configuration");
    DH_CHECK_ARG(batch >= 0 && max_total >= 1 && pre_nms_topk >= 1, "dh_fcos_detect: bad sizes");
    DH_CHECK_ARG(center >= DH_FCOS_SCORE_CLS && center <= DH_FCOS_SCORE_MAP_CEN, "dh_fcos_detect: score mode %d", center);
    if (batch == 0) return DH_OK;
    DeviceGuard guard(h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long longest = 1;
    for (int l = 0; l < n_levels; ++l) {
        DH_CHECK_ARG(pred_levels[l] && strides[l] > 0, "dh_fcos_detect: level %d", l);
        const long long n = static_cast<long long>(static_cast<int>(static_cast<double>(pad_h) / strides[l])) *
                            static_cast<int>(static_cast<double>(pad_w) / strides[l]) * num_classes;
        DH_CHECK_ARG(n < (1ll << 31), "dh_fcos_detect: level %d has too many (location, class) pairs", l);
        if (n > longest) longest = n;
    }
    const int k = static_cast<int>(pre_nms_topk < longest ? pre_nms_topk : longest);
    const int n_cand = n_levels * k;
    if (n_cand > 16384)
        return set_error(DH_ERR_CAPACITY, "dh_fcos_detect: pre_nms_topk * levels = %d exceeds the NMS capacity of 16384", n_cand);
    const size_t cand_bytes = (static_cast<size_t>(batch) * n_cand * 6 * 4 + 255) & ~size_t(255);
    char* sc = static_cast<char*>(scratch_b(h, cand_bytes + static_cast<size_t>(batch) * max_total * 4 + 256));
    if (!sc) return DH_ERR_CUDA;
    float* cand = out_cand ? out_cand : reinterpret_cast<float*>(sc);
    int32_t* keep = reinterpret_cast<int32_t*>(sc + cand_bytes);
    int rc = launch_fcos_select(h, pred_levels, batch, pad_h, pad_w, n_levels, strides, num_classes, center, k, cls_thr, 0, cand, st);
    if (rc) return rc;
    rc = dh_nms(h, cand, nullptr, batch, n_cand, 6, DH_NMS_PER_CLASS, iou_thr, cls_thr, 0, num_classes, max_per_class, max_total, keep,
                max_total, out_valid, stream);
    if (rc) return rc;
    dim3 grid((max_total + 127) / 128, batch);
    gather_fcos_outputs_kernel<<<grid, 128, 0, st>>>(cand, n_cand, keep, out_valid, max_total, out_boxes, out_scores, out_classes);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

int dh_retina_detect(dh_handle_t h, const float* const* pred_levels, int batch, int pad_h, int pad_w, int n_levels,
                     const int32_t* strides, int n_anchors, const float* anchor_hw_dev, int num_classes, float iou_thr,
                     float cls_thr, int pre_nms_topk, float* out_rows, int max_out, int32_t* out_n, int32_t* out_overflow,
                     float* out_cand, int32_t* out_keep, void* stream) {
    DH_CHECK_ARG(h && pred_levels && strides && anchor_hw_dev && out_rows && out_n, "dh_retina_detect: NULL argument");
    DH_CHECK_ARG(n_levels >= 1 && n_levels <= DH_MAX_PYRAMID_LEVELS && n_anchors >= 1 && num_classes >= 1, "dh_retina_detect: bad configuration");
    DH_CHECK_ARG(batch >= 0 && max_out >= 1 && pre_nms_topk >= 0, "dh_retina_detect: bad sizes");
    if (batch == 0) return DH_OK;
    DeviceGuard guard(h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int seg[DH_MAX_PYRAMID_LEVELS + 1];
    seg[0] = 0;
    long long n_total = 0, longest = 1;
    for (int l = 0; l < n_levels; ++l) {
        DH_CHECK_ARG(pred_levels[l] && strides[l] > 0, "dh_retina_detect: level %d", l);
        const long long n = static_cast<long long>(n_anchors) * static_cast<int>(static_cast<double>(pad_h) / strides[l]) *
                            static_cast<int>(static_cast<double>(pad_w) / strides[l]);
        n_total += n;
        DH_CHECK_ARG(n_total < (1ll << 31), "dh_retina_detect: too many anchors");
        seg[l + 1] = static_cast<int>(n_total);
        if (n > longest) longest = n;
    }
    // candidates per level: the caller's top-k, or (reference semantics, threshold only) as many as the NMS can take
    int k = pre_nms_topk > 0 ? pre_nms_topk : 16384 / n_levels;
    if (k > longest) k = static_cast<int>(longest);
    const int n_cand = n_levels * k;
    if (n_cand > 16384)
        return set_error(DH_ERR_CAPACITY, "dh_retina_detect: pre_nms_topk * levels = %d exceeds the NMS capacity of 16384", n_cand);
    const size_t dets_bytes = (static_cast<size_t>(batch) * n_total * 6 * 4 + 255) & ~size_t(255);
    const size_t cand_bytes = (static_cast<size_t>(batch) * n_cand * 6 * 4 + 255) & ~size_t(255);
    const size_t keep_bytes = (static_cast<size_t>(batch) * max_out * 4 + 255) & ~size_t(255);
    const size_t score_bytes = (static_cast<size_t>(batch) * n_total * 4 + 255) & ~size_t(255);
    char* sc = static_cast<char*>(scratch_b(h, dets_bytes + cand_bytes + keep_bytes + score_bytes + 256));
    if (!sc) return DH_ERR_CUDA;
    float* dets = reinterpret_cast<float*>(sc);
    float* cand = out_cand ? out_cand : reinterpret_cast<float*>(sc + dets_bytes);
    int32_t* keep = out_keep ? out_keep : reinterpret_cast<int32_t*>(sc + dets_bytes + cand_bytes);
    // the decode also writes the score column on its own: the selector's three passes over an image's 76 725 rows then
    // read 4 bytes per row instead of one 32-byte sector of the 24-byte rows
    float* scores = reinterpret_cast<float*>(sc + dets_bytes + cand_bytes + keep_bytes);
    int have_scores = 0;
    int rc = retina_decode_scores(h, pred_levels, batch, pad_h, pad_w, n_levels, strides, n_anchors, anchor_hw_dev, num_classes, dets, scores,
                                  &have_scores, stream);
    if (rc) return rc;
    if (out_overflow) DH_CUDA(cudaMemsetAsync(out_overflow, 0, sizeof(int32_t) * batch, st));
    rc = launch_select_segs(h, dets, batch, n_total, 6, 4, seg, n_levels, k, cls_thr, 1, cand, pre_nms_topk > 0 ? nullptr : out_overflow, st,
                            have_scores ? scores : nullptr);
    if (rc) return rc;
    rc = dh_nms(h, cand, nullptr, batch, n_cand, 6, DH_NMS_AGNOSTIC, iou_thr, cls_thr, 1, 0, 0, 0, keep, max_out, out_n, stream);
    if (rc) return rc;
    dim3 grid((max_out * 6 + 255) / 256 > 64 ? 64 : (max_out * 6 + 255) / 256, batch);
    gather_kept_rows_kernel<<<grid, 256, 0, st>>>(cand, n_cand, keep, out_n, max_out, out_rows);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

}  // extern "C"
