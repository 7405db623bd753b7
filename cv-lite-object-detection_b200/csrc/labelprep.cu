// The steps immediately before and after the dense-head path, on the device so
// that a training / inference step never has to round-trip small box arrays through the host:
//
//   dh_box_convert        swap_xy / convert_to_xywh / convert_to_corners (FCOS/utils.py:6-40, same in the other
//                         folders) and the box half of random_flip_horizontal (FCOS/data_preprocess.py:24-41)
//   dh_prepare_labels     dataset boxes (xmin, ymin, xmax, ymax normalised, FCOS/format_VOC_fcos.py:60-68) + class ids
//                         -> the padded [B, max_boxes, 5] (cy, cx, h, w, class) layout every encoder takes:
//                         optional horizontal flip, swap_xy, convert_to_xywh (FCOS/data_preprocess.py:121-131,
//                         FCOS/train_fcos.py:131-135), from either a padded or a ragged (offsets) input
//   dh_format_detections  RetinaNet.detect_bboxes after the NMS (RetinaNet/retinanet_module.py:559-569): rescale the
//                         kept rows to the source image and swap to (x1, y1, x2, y2); scores and integer labels
//
// All float32 in the reference's operation order (TF float32 ops); detect_bboxes multiplies in float64 (a NumPy
// float64 ratio array), which is reproduced and then rounded to float32.
#include "dh_common.cuh"
#include "dh_host.h"

namespace dh {

__device__ __forceinline__ float4 convert_box(float4 b, int mode) {
    switch (mode) {
        case DH_BOX_SWAP_XY:
            return make_float4(b.y, b.x, b.w, b.z);
        case DH_BOX_TO_XYWH:  // [(lo + hi) / 2, hi - lo]
            return make_float4(fdiv(fadd(b.x, b.z), 2.0f), fdiv(fadd(b.y, b.w), 2.0f), fsub(b.z, b.x), fsub(b.w, b.y));
        case DH_BOX_TO_CORNERS: {  // [c - size / 2, c + size / 2]
            const float hx = fdiv(b.z, 2.0f), hy = fdiv(b.w, 2.0f);
            return make_float4(fsub(b.x, hx), fsub(b.y, hy), fadd(b.x, hx), fadd(b.y, hy));
        }
        default:  // DH_BOX_FLIP_HORIZONTAL on (xmin, ymin, xmax, ymax) normalised
            return make_float4(fsub(1.0f, b.z), b.y, fsub(1.0f, b.x), b.w);
    }
}

__global__ void box_convert_kernel(const float4* __restrict__ in, long long n, int mode, float4* __restrict__ out) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = convert_box(in[i], mode);
}

__global__ void prepare_labels_kernel(const float* __restrict__ raw, const float* __restrict__ classes, const int* __restrict__ offsets,
                                      const int* __restrict__ nbox, const int* __restrict__ flip, int batch, int max_boxes, int in_stride,
                                      float* __restrict__ out, int* __restrict__ out_nbox, int* __restrict__ status) {
    const int total = batch * max_boxes;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int b = e / max_boxes, k = e - b * max_boxes;
        const int n_in = offsets ? offsets[b + 1] - offsets[b] : (nbox ? nbox[b] : max_boxes);
        const int n = max(0, min(n_in, max_boxes));
        float* o = out + static_cast<long long>(e) * 5;
        if (k == 0 && out_nbox) out_nbox[b] = n;
        if (k == 0 && n_in > max_boxes && status) atomicOr(status, DH_STATUS_TRUNCATED);
        if (k >= n) {
            o[0] = o[1] = o[2] = o[3] = o[4] = 0.f;
            continue;
        }
        const long long src = offsets ? static_cast<long long>(offsets[b]) + k : static_cast<long long>(b) * in_stride + k;
        float4 bx = make_float4(raw[src * 4], raw[src * 4 + 1], raw[src * 4 + 2], raw[src * 4 + 3]);
        if (flip && flip[b]) bx = convert_box(bx, DH_BOX_FLIP_HORIZONTAL);
        bx = convert_box(convert_box(bx, DH_BOX_SWAP_XY), DH_BOX_TO_XYWH);
        o[0] = bx.x, o[1] = bx.y, o[2] = bx.z, o[3] = bx.w, o[4] = classes[src];
    }
}

__global__ void format_detections_kernel(const float* __restrict__ rows, const int* __restrict__ n_keep, const float* __restrict__ ratios,
                                         int batch, int n, float* __restrict__ boxes, float* __restrict__ scores, int* __restrict__ labels) {
    const int total = batch * n;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int b = e / n, r = e - b * n;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        float sc = 0.f;
        int lb = -1;
        if (r < n_keep[b]) {
            const float* q = rows + static_cast<long long>(e) * 6;
            const double wr = static_cast<double>(ratios[2 * b]), hr = static_cast<double>(ratios[2 * b + 1]);
            // tmp_detect[:, :4] * [w_ratio, h_ratio, w_ratio, h_ratio], then swap_xy (retinanet_module.py:562-567)
            o = make_float4(static_cast<float>(dmul(static_cast<double>(q[1]), hr)), static_cast<float>(dmul(static_cast<double>(q[0]), wr)),
                            static_cast<float>(dmul(static_cast<double>(q[3]), hr)), static_cast<float>(dmul(static_cast<double>(q[2]), wr)));
            sc = q[4], lb = __float2int_rz(q[5]);
        }
        reinterpret_cast<float4*>(boxes)[e] = o;
        scores[e] = sc, labels[e] = lb;
    }
}

// show_heatmap's rectangles (FCOS/train_fcos_center_voc.py:92-121): float32 throughout (a TF tensor times a ratio array)
__global__ void fcos_rectangles_kernel(const float4* __restrict__ boxes, const int* __restrict__ valid, const float* __restrict__ ratios,
                                       int batch, int t, float4* __restrict__ rect) {
    const int total = batch * t;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int b = e / t, r = e - b * t;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < valid[b]) {
            const float4 q = boxes[e];  // (y1, x1, y2, x2)
            const float wr = ratios[2 * b], hr = ratios[2 * b + 1];
            float x1 = fmul(q.y, hr), y1 = fmul(q.x, wr);
            const float x2 = fmul(q.w, hr), y2 = fmul(q.z, wr);
            if (x1 <= 0.f) x1 = 0.f;
            if (y1 <= 0.f) y1 = 0.f;
            o = make_float4(x1, y1, fsub(x2, x1), fsub(y2, y1));
        }
        rect[e] = o;
    }
}

static int grid1d(long long n, int block, int sm_count) {
    long long g = (n + block - 1) / block;
    const long long cap = static_cast<long long>(sm_count) * 8;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dh

using namespace dh;

extern "C" {

int dh_box_convert(dh_handle_t h, const float* boxes, long long n, int mode, float* out, void* stream) {
    DH_CHECK_ARG(n >= 0 && mode >= DH_BOX_SWAP_XY && mode <= DH_BOX_FLIP_HORIZONTAL, "dh_box_convert: bad arguments");
    if (n == 0) return DH_OK;
    DH_CHECK_ARG(h && boxes && out, "dh_box_convert: NULL argument");
    DH_CHECK_ARG(((reinterpret_cast<uintptr_t>(boxes) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0, "dh_box_convert: buffers must be 16-byte aligned");
    if (n == 0) return DH_OK;
    DeviceGuard guard(h);
    box_convert_kernel<<<grid1d(n, 256, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(boxes), n, mode, reinterpret_cast<float4*>(out));
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

int dh_prepare_labels(dh_handle_t h, const float* raw_boxes, const float* classes, const int32_t* box_offsets, const int32_t* nbox,
                      const int32_t* flip, int batch, int in_max_boxes, int max_boxes, float* out_labels, int32_t* out_nbox,
                      void* stream) {
    DH_CHECK_ARG(h && raw_boxes && classes && out_labels, "dh_prepare_labels: NULL argument");
    DH_CHECK_ARG(batch >= 0 && max_boxes >= 1 && (box_offsets || in_max_boxes >= 0), "dh_prepare_labels: bad sizes");
    if (batch == 0) return DH_OK;
    DeviceGuard guard(h);
    prepare_labels_kernel<<<grid1d(static_cast<long long>(batch) * max_boxes, 256, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        raw_boxes, classes, box_offsets, nbox, flip, batch, max_boxes, in_max_boxes, out_labels, out_nbox, h->dev_status);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

int dh_format_detections(dh_handle_t h, const float* rows, const int32_t* n_keep, const float* ratios, int batch, int n,
                         float* out_boxes, float* out_scores, int32_t* out_labels, void* stream) {
    DH_CHECK_ARG(h && rows && n_keep && ratios && out_boxes && out_scores && out_labels, "dh_format_detections: NULL argument");
    DH_CHECK_ARG(batch >= 0 && n >= 0, "dh_format_detections: bad sizes");
    DH_CHECK_ARG((reinterpret_cast<uintptr_t>(out_boxes) & 15u) == 0, "dh_format_detections: out_boxes must be 16-byte aligned");
    if (batch == 0 || n == 0) return DH_OK;
    DeviceGuard guard(h);
    format_detections_kernel<<<grid1d(static_cast<long long>(batch) * n, 256, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rows, n_keep, ratios, batch, n, out_boxes, out_scores, out_labels);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

int dh_fcos_rectangles(dh_handle_t h, const float* boxes, const int32_t* valid, const float* ratios, int batch, int t, float* out_rect,
                       void* stream) {
    DH_CHECK_ARG(h && boxes && valid && ratios && out_rect, "dh_fcos_rectangles: NULL argument");
    DH_CHECK_ARG(batch >= 0 && t >= 0, "dh_fcos_rectangles: bad sizes");
    DH_CHECK_ARG(((reinterpret_cast<uintptr_t>(boxes) | reinterpret_cast<uintptr_t>(out_rect)) & 15u) == 0,
                 "dh_fcos_rectangles: boxes and out_rect must be 16-byte aligned");
    if (batch == 0 || t == 0) return DH_OK;
    DeviceGuard guard(h);
    fcos_rectangles_kernel<<<grid1d(static_cast<long long>(batch) * t, 256, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(boxes), valid, ratios, batch, t, reinterpret_cast<float4*>(out_rect));
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

}  // extern "C"
