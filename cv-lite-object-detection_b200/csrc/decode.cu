// Inference-time decode: head outputs -> pixel boxes + scores, and the pre-NMS candidate selection.
//
//   dh_prediction_to_corners   the four prediction_to_corners variants of the reference
//                              (FCOS/fcos.py:112-134, fcos_center_v1.py:125-147,
//                               retinanet_module.py:428-451, tf_centernet_resnet_s8.py:210-241)
//   dh_fcos_decode             what FCOS/infer_fcos.py:35-57 hands to the NMS op: boxes [B,N,4], scores [B,N,C]
//   dh_retina_decode           RetinaNet.image_detections up to the threshold (retinanet_module.py:487-520):
//                              dets [B,N,6] = (y1, x1, y2, x2, max score, argmax label), level > anchor > row-major
//   dh_select_topk             per-segment (= per pyramid level) score threshold + exact top-k, stable in index
//                              order (radix select: 12+12+8 bit histograms in shared memory)
//
// These kernels are HBM-bound on one read of the head output; thread mapping is one thread per output
// row with the class loop vectorised where the row is 16-byte aligned.
#include <cstring>

#include "dh_common.cuh"
#include "dh_host.h"

namespace dh {

__device__ __forceinline__ float sigmoid_acc(float x) { return fdiv(1.0f, fadd(1.0f, expf(-x))); }

struct CornerParams {
    int mode, height, width, sub, ch_in;  // rows = height*width*sub, input row stride ch_in (first 4 used)
    float stride, d0, d1;                 // retina: anchor (h, w); v1: box_sc in d0
    float scales[8];                      // s8
};

// mode 0: FCOS tblr, centres at i+.5, float32 differences then `stride *` in float64 (the reference's container)
// mode 1: RetinaNet   c = i*s - p*a, size = p*a
// mode 2: fcos_center_v1  c = (i + p)*s, size = p*box_sc
// mode 3: CenterNet s8    c = (i + p)*s, size = p*scale[sub index]
__device__ __forceinline__ float4 corners_of(const CornerParams& p, int i, int j, int sub, float p0, float p1, float p2, float p3) {
    float4 o;
    if (p.mode == 0) {
        const float gy = static_cast<float>(i) + 0.5f, gx = static_cast<float>(j) + 0.5f;
        const double s = static_cast<double>(p.stride);
        o.x = static_cast<float>(dmul(s, static_cast<double>(fsub(gy, p0))));
        o.y = static_cast<float>(dmul(s, static_cast<double>(fsub(gx, p2))));
        o.z = static_cast<float>(dmul(s, static_cast<double>(fadd(gy, p1))));
        o.w = static_cast<float>(dmul(s, static_cast<double>(fadd(gx, p3))));
        return o;
    }
    float yc, xc, bh, bw;
    const float gy = static_cast<float>(i), gx = static_cast<float>(j);
    if (p.mode == 1) {
        yc = fsub(fmul(gy, p.stride), fmul(p0, p.d0)), xc = fsub(fmul(gx, p.stride), fmul(p1, p.d1));
        bh = fmul(p2, p.d0), bw = fmul(p3, p.d1);
    } else {
        const float sc = p.mode == 2 ? p.d0 : p.scales[sub];
        yc = fmul(fadd(gy, p0), p.stride), xc = fmul(fadd(gx, p1), p.stride);
        bh = fmul(p2, sc), bw = fmul(p3, sc);
    }
    const float hh = fdiv(bh, 2.0f), hw = fdiv(bw, 2.0f);
    o.x = fsub(yc, hh), o.y = fsub(xc, hw), o.z = fadd(yc, hh), o.w = fadd(xc, hw);
    return o;
}

__global__ void corners_kernel(const float* __restrict__ pred, long long rows_total, CornerParams p, float* __restrict__ out) {
    const long long rows_per_image = static_cast<long long>(p.height) * p.width * p.sub;
    for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < rows_total;
         r += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long local = r % rows_per_image;
        const int sub = static_cast<int>(local % p.sub);
        const long long cell = local / p.sub;
        const int i = static_cast<int>(cell / p.width), j = static_cast<int>(cell % p.width);
        const float* q = pred + r * p.ch_in;
        const float4 o = corners_of(p, i, j, sub, q[0], q[1], q[2], q[3]);
        reinterpret_cast<float4*>(out)[r] = o;
    }
}

// FCOS decode of one level: one warp per location; lanes stride over the classes.
__global__ void fcos_decode_kernel(const float* __restrict__ pred, int batch, int hl, int wl, int num_classes, float stride,
                                   int center, long long n_total, long long level_off, float* __restrict__ boxes,
                                   float* __restrict__ scores) {
    const int ch = num_classes + 5;
    const long long rows = static_cast<long long>(batch) * hl * wl;
    const int lane = threadIdx.x & 31;
    CornerParams cp;
    cp.mode = 0, cp.stride = stride;
    for (long long r = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; r < rows;
         r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
        const int b = static_cast<int>(r / (static_cast<long long>(hl) * wl));
        const int loc = static_cast<int>(r - static_cast<long long>(b) * hl * wl);
        const float* q = pred + r * ch;
        const long long orow = static_cast<long long>(b) * n_total + level_off + loc;
        if (lane == 0) {
            const float4 o = corners_of(cp, loc / wl, loc % wl, 0, q[0], q[1], q[2], q[3]);
            reinterpret_cast<float4*>(boxes)[orow] = o;
        }
        const float cen = center ? sigmoid_acc(q[4]) : 1.0f;
        for (int c = lane; c < num_classes; c += 32) {
            const float s = sigmoid_acc(q[5 + c]);
            scores[orow * num_classes + c] = center ? fmul(cen, s) : s;
        }
    }
}

// RetinaNet decode of one level: one warp per anchor row; max / first-argmax over classes by shuffle.
__global__ void retina_decode_kernel(const float* __restrict__ pred, int batch, int n_anchors, int hl, int wl, int num_classes,
                                     float stride, const float* __restrict__ anchor_hw /*[n_anchors,2] device*/,
                                     long long n_total, long long level_off, float* __restrict__ dets) {
    const int ch = num_classes + 4;
    const long long per_img = static_cast<long long>(n_anchors) * hl * wl;
    const long long rows = static_cast<long long>(batch) * per_img;
    const int lane = threadIdx.x & 31;
    for (long long r = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; r < rows;
         r += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
        const int b = static_cast<int>(r / per_img);
        const long long local = r - static_cast<long long>(b) * per_img;  // anchor-major, then row-major cells
        const int an = static_cast<int>(local / (static_cast<long long>(hl) * wl));
        const int loc = static_cast<int>(local - static_cast<long long>(an) * hl * wl);
        const float* q = pred + r * ch;
        float best = -1.0f;
        int best_c = 0x7fffffff;
        for (int c = lane; c < num_classes; c += 32) {
            const float s = sigmoid_acc(q[4 + c]);
            if (s > best) best = s, best_c = c;  // ascending c per lane: first max within the lane
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
            if (ob > best || (ob == best && oc < best_c)) best = ob, best_c = oc;  // np.argmax: first maximum
        }
        if (lane == 0) {
            CornerParams cp;
            cp.mode = 1, cp.stride = stride, cp.d0 = anchor_hw[2 * an], cp.d1 = anchor_hw[2 * an + 1];
            const float4 o = corners_of(cp, loc / wl, loc % wl, 0, q[0], q[1], q[2], q[3]);
            float* d = dets + (static_cast<long long>(b) * n_total + level_off + local) * 6;
            d[0] = o.x, d[1] = o.y, d[2] = o.z, d[3] = o.w, d[4] = best, d[5] = static_cast<float>(best_c);
        }
    }
}

// ---- exact per-segment top-k (radix select), stable in index order ------------------------------------
__device__ __forceinline__ unsigned score_key(float s) {
    const unsigned u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone float -> uint
}

constexpr int kSelThreads = 1024;

__global__ void __launch_bounds__(kSelThreads) select_topk_kernel(const float* __restrict__ dets, long long n_total, int row_floats, int score_col,
                                                                  const int* __restrict__ seg_off /*[n_seg+1] device*/, int k_slots,
                                                                  float min_score, int inclusive, float* __restrict__ out,
                                                                  int* __restrict__ out_src, int out_rows) {
    __shared__ unsigned hist[4096];
    __shared__ unsigned s_prefix, s_need, s_scan[kSelThreads / 32], s_carry_gt, s_carry_eq;
    const int b = blockIdx.y, seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lo = seg_off[seg], hi = seg_off[seg + 1];
    const float* d = dets + static_cast<long long>(b) * n_total * row_floats;
    float* o = out + (static_cast<long long>(b) * out_rows + static_cast<long long>(seg) * k_slots) * row_floats;
    int* osrc = out_src ? out_src + static_cast<long long>(b) * out_rows + static_cast<long long>(seg) * k_slots : nullptr;
    int k = k_slots;
    if (k > hi - lo) k = max(hi - lo, 0);  // a segment shorter than k: the slots beyond its length stay padded
    auto passes = [&](float s) { return inclusive ? (s >= min_score) : (s > min_score); };

    // three histogram rounds narrow the k-th largest key: bits [31:20], [19:8], [7:0]
    unsigned prefix = 0, need = static_cast<unsigned>(k);  // keys matching `prefix` on the bits decided so far
    bool take_all = false;
    for (int round = 0; round < 3 && !take_all; ++round) {
        const int shift = round == 0 ? 20 : (round == 1 ? 8 : 0);
        const int bits = round == 2 ? 8 : 12;
        const unsigned decided_mask = round == 0 ? 0u : (round == 1 ? 0xFFF00000u : 0xFFFFFF00u);
        for (int i = tid; i < (1 << bits); i += kSelThreads) hist[i] = 0;
        __syncthreads();
        for (int i = lo + tid; i < hi; i += kSelThreads) {
            const float s = d[static_cast<long long>(i) * row_floats + score_col];
            if (!passes(s)) continue;
            const unsigned key = score_key(s);
            if ((key & decided_mask) == (prefix & decided_mask)) atomicAdd(&hist[(key >> shift) & ((1u << bits) - 1u)], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned cum = 0;
            int bin = (1 << bits) - 1;
            for (; bin >= 0; --bin) {
                if (cum + hist[bin] >= need) break;
                cum += hist[bin];
            }
            if (bin < 0) {
                s_prefix = 0xFFFFFFFFu;  // fewer than `need` candidates: take everything that passes
                s_need = 0;
            } else {
                s_prefix = prefix | (static_cast<unsigned>(bin) << shift);
                s_need = need - cum;  // still to take among keys equal to the new prefix
            }
        }
        __syncthreads();
        if (s_prefix == 0xFFFFFFFFu && s_need == 0 && round == 0) take_all = true;
        if (!take_all) prefix = s_prefix, need = s_need;
        __syncthreads();
    }
    // stable compaction: key > T always, key == T for the first `need` in index order
    const unsigned T = prefix;
    if (tid == 0) s_carry_gt = 0, s_carry_eq = 0;
    __syncthreads();
    for (int base = lo; base < hi; base += kSelThreads) {
        const int i = base + tid;
        bool gt = false, eq = false;
        if (i < hi) {
            const float s = d[static_cast<long long>(i) * row_floats + score_col];
            if (passes(s)) {
                const unsigned key = score_key(s);
                if (take_all) gt = true;
                else gt = key > T, eq = key == T;
            }
        }
        // block-wide exclusive scans of gt and eq (packed: eq in the high 16 bits; chunk <= 1024 elements)
        const unsigned v = (gt ? 1u : 0u) | (eq ? (1u << 16) : 0u);
        unsigned incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned w = s_scan[lane];
            unsigned wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += t;
            }
            s_scan[lane] = wi - w;  // exclusive prefix of warp totals
        }
        __syncthreads();
        const unsigned excl = incl - v + s_scan[warp];
        const unsigned gt_before = (excl & 0xFFFFu) + s_carry_gt, eq_before = (excl >> 16) + s_carry_eq;
        bool take = gt;
        if (eq && eq_before < need) take = true;
        // rank among taken elements, in index order
        const unsigned eq_taken_before = eq_before < need ? eq_before : need;
        const unsigned rank = gt_before + eq_taken_before;
        if (take && rank < static_cast<unsigned>(k)) {
            const float* src = d + static_cast<long long>(i) * row_floats;
            float* dst = o + static_cast<long long>(rank) * row_floats;
            for (int c = 0; c < row_floats; ++c) dst[c] = src[c];
            if (osrc) osrc[rank] = i;
        }
        __syncthreads();
        if (tid == kSelThreads - 1) {
            s_carry_gt += (excl & 0xFFFFu) + (gt ? 1u : 0u);
            s_carry_eq += (excl >> 16) + (eq ? 1u : 0u);
        }
        __syncthreads();
    }
    // pad the unused slots with score = -inf so that any threshold drops them
    const unsigned taken_eq = s_carry_eq < need ? s_carry_eq : need;
    const unsigned filled = min(static_cast<unsigned>(k), s_carry_gt + (take_all ? 0u : taken_eq));
    for (int r = filled + tid; r < k_slots; r += kSelThreads) {
        float* dst = o + static_cast<long long>(r) * row_floats;
        for (int c = 0; c < row_floats; ++c) dst[c] = 0.f;
        dst[score_col] = -INFINITY;
        if (osrc) osrc[r] = -1;
    }
}

static int grid_for(long long threads_needed, int block, int sm_count) {
    long long g = (threads_needed + block - 1) / block;
    const long long cap = static_cast<long long>(sm_count) * 16;
    if (g > cap) g = cap;
    return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace dh

using namespace dh;

extern "C" {

int dh_prediction_to_corners(dh_handle_t h, const float* pred, int batch, int height, int width, int sub, int ch_in, int mode,
                             float stride, float d0, float d1, const float* scales, float* out, void* stream) {
    DH_CHECK_ARG(h && pred && out, "dh_prediction_to_corners: NULL argument");
    DH_CHECK_ARG(mode >= 0 && mode <= 3, "dh_prediction_to_corners: mode %d", mode);
    DH_CHECK_ARG(batch >= 0 && height >= 0 && width >= 0 && sub >= 1 && sub <= 8 && ch_in >= 4, "dh_prediction_to_corners: bad sizes");
    DH_CHECK_ARG(mode != 3 || scales, "dh_prediction_to_corners: mode 3 needs scales");
    DH_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "dh_prediction_to_corners: out must be 16-byte aligned");
    const long long rows = static_cast<long long>(batch) * height * width * sub;
    if (rows == 0) return DH_OK;
    DeviceGuard guard(h->device);
    CornerParams p;
    memset(&p, 0, sizeof(p));
    p.mode = mode, p.height = height, p.width = width, p.sub = sub, p.ch_in = ch_in, p.stride = stride, p.d0 = d0, p.d1 = d1;
    for (int k = 0; k < sub && mode == 3; ++k) p.scales[k] = scales[k];
    corners_kernel<<<grid_for(rows, 256, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, rows, p, out);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

int dh_fcos_decode(dh_handle_t h, const float* const* pred_levels, int batch, int pad_h, int pad_w, int n_levels,
                   const int32_t* strides, int num_classes, int center, float* boxes, float* scores, void* stream) {
    DH_CHECK_ARG(h && pred_levels && strides && boxes && scores, "dh_fcos_decode: NULL argument");
    DH_CHECK_ARG(n_levels >= 1 && n_levels <= DH_MAX_LEVELS && num_classes >= 1, "dh_fcos_decode: bad configuration");
    DH_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15u) == 0, "dh_fcos_decode: boxes must be 16-byte aligned");
    DeviceGuard guard(h->device);
    long long n_total = 0;
    for (int l = 0; l < n_levels; ++l)
        n_total += static_cast<long long>(static_cast<int>(static_cast<double>(pad_h) / strides[l])) *
                   static_cast<int>(static_cast<double>(pad_w) / strides[l]);
    long long off = 0;
    for (int l = 0; l < n_levels; ++l) {
        DH_CHECK_ARG(pred_levels[l] && strides[l] > 0, "dh_fcos_decode: level %d", l);
        const int hl = static_cast<int>(static_cast<double>(pad_h) / strides[l]), wl = static_cast<int>(static_cast<double>(pad_w) / strides[l]);
        const long long rows = static_cast<long long>(batch) * hl * wl;
        if (rows > 0) {
            fcos_decode_kernel<<<grid_for(rows * 32, 256, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                pred_levels[l], batch, hl, wl, num_classes, static_cast<float>(strides[l]), center, n_total, off, boxes, scores);
            DH_CUDA(cudaGetLastError());
            h->launches += 1;
        }
        off += static_cast<long long>(hl) * wl;
    }
    return DH_OK;
}

int dh_retina_decode(dh_handle_t h, const float* const* pred_levels, int batch, int pad_h, int pad_w, int n_levels,
                     const int32_t* strides, int n_anchors, const float* anchor_hw_dev, int num_classes, float* dets,
                     void* stream) {
    DH_CHECK_ARG(h && pred_levels && strides && anchor_hw_dev && dets, "dh_retina_decode: NULL argument");
    DH_CHECK_ARG(n_levels >= 1 && n_levels <= DH_MAX_LEVELS && n_anchors >= 1 && num_classes >= 1, "dh_retina_decode: bad configuration");
    DeviceGuard guard(h->device);
    long long n_total = 0;
    for (int l = 0; l < n_levels; ++l)
        n_total += static_cast<long long>(n_anchors) * static_cast<int>(static_cast<double>(pad_h) / strides[l]) *
                   static_cast<int>(static_cast<double>(pad_w) / strides[l]);
    long long off = 0;
    for (int l = 0; l < n_levels; ++l) {
        DH_CHECK_ARG(pred_levels[l] && strides[l] > 0, "dh_retina_decode: level %d", l);
        const int hl = static_cast<int>(static_cast<double>(pad_h) / strides[l]), wl = static_cast<int>(static_cast<double>(pad_w) / strides[l]);
        const long long rows = static_cast<long long>(batch) * n_anchors * hl * wl;
        if (rows > 0) {
            retina_decode_kernel<<<grid_for(rows * 32, 256, h->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                pred_levels[l], batch, n_anchors, hl, wl, num_classes, static_cast<float>(strides[l]),
                anchor_hw_dev + static_cast<long long>(l) * n_anchors * 2, n_total, off, dets);
            DH_CUDA(cudaGetLastError());
            h->launches += 1;
        }
        off += static_cast<long long>(n_anchors) * hl * wl;
    }
    return DH_OK;
}

int dh_select_topk(dh_handle_t h, const float* dets, int batch, long long n_total, int row_floats, int score_col,
                   const int32_t* seg_off_dev, int n_seg, int k, float min_score, int score_inclusive, float* out,
                   int32_t* out_src, void* stream) {
    DH_CHECK_ARG(h && dets && seg_off_dev && out, "dh_select_topk: NULL argument");
    DH_CHECK_ARG(batch >= 0 && n_total >= 0 && n_total < (1ll << 31) && row_floats >= 1 && score_col >= 0 && score_col < row_floats &&
                     n_seg >= 1 && k >= 1 && k <= 65535,
                 "dh_select_topk: bad sizes");
    if (batch == 0) return DH_OK;
    DeviceGuard guard(h->device);
    dim3 grid(n_seg, batch);
    select_topk_kernel<<<grid, kSelThreads, 0, static_cast<cudaStream_t>(stream)>>>(dets, n_total, row_floats, score_col, seg_off_dev, k, min_score,
                                                                                   score_inclusive, out, out_src, n_seg * k);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

}  // extern "C"
