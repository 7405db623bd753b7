// Pairwise IoU helpers and the CenterNet (soft-)NMS.
//
//   dh_compute_iou     RetinaNet/utils.py:42-83  -- float32 pairwise IoU of centre-size boxes
//   dh_bboxes_iou      CenterNet/tf_centernet_resnet_s8.py:22-42 -- float64 IoU of corner boxes, floored at
//                      float32 eps (one box against n, or n against n element-wise)
//   dh_centernet_nms   CenterNet/tf_centernet_resnet_s8.py:44-85 -- per-class greedy argmax-pop NMS in float64 on
//                      (xmin, ymin, w, h, score, class) rows, hard (`iou > thr` zeroes the score) or soft
//                      (`score *= exp(-iou^2 / sigma)`); rows whose score drops to <= 0 leave the pool.
//
// The NMS is inherently sequential in the number of kept boxes; one CTA walks the classes in ascending
// order and does the argmax (block reduction, first index on ties like np.argmax) and the re-weighting of
// the remaining boxes in parallel.
#include <cfloat>
#include <cstring>

#include "dh_common.cuh"
#include "dh_host.h"

namespace dh {

__global__ void compute_iou_kernel(const float* __restrict__ b1, int n, const float* __restrict__ b2, int m, float* __restrict__ out) {
    const long long total = static_cast<long long>(n) * m;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(e / m), j = static_cast<int>(e - static_cast<long long>(i) * m);
        const float* p = b1 + 4 * i;
        const float* q = b2 + 4 * j;
        const float ph0 = fdiv(p[2], 2.0f), ph1 = fdiv(p[3], 2.0f), qh0 = fdiv(q[2], 2.0f), qh1 = fdiv(q[3], 2.0f);
        const float lo0 = fmaxf(fsub(p[0], ph0), fsub(q[0], qh0)), lo1 = fmaxf(fsub(p[1], ph1), fsub(q[1], qh1));
        const float hi0 = fminf(fadd(p[0], ph0), fadd(q[0], qh0)), hi1 = fminf(fadd(p[1], ph1), fadd(q[1], qh1));
        const float inter = fmul(fmaxf(0.0f, fsub(hi0, lo0)), fmaxf(0.0f, fsub(hi1, lo1)));
        const float uni = fmaxf(fsub(fadd(fmul(p[2], p[3]), fmul(q[2], q[3])), inter), 1e-8f);
        out[e] = fminf(fmaxf(fdiv(inter, uni), 0.0f), 1.0f);
    }
}

__device__ __forceinline__ double iou64(const double* a, const double* b) {
    const double area_a = dmul(dsub(a[2], a[0]), dsub(a[3], a[1]));
    const double area_b = dmul(dsub(b[2], b[0]), dsub(b[3], b[1]));
    const double w = fmax(dsub(fmin(a[2], b[2]), fmax(a[0], b[0])), 0.0);
    const double h = fmax(dsub(fmin(a[3], b[3]), fmax(a[1], b[1])), 0.0);
    const double inter = dmul(w, h);
    const double v = ddiv(dmul(1.0, inter), dsub(dadd(area_a, area_b), inter));
    return fmax(v, static_cast<double>(FLT_EPSILON));  // (0/0 gives eps here, NaN in NumPy: zero-area pairs are out of contract)
}

__global__ void bboxes_iou_kernel(const double* __restrict__ b1, int n1, const double* __restrict__ b2, int n2, double* __restrict__ out) {
    const int n = max(n1, n2);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
        out[e] = iou64(b1 + 4 * (n1 == 1 ? 0 : e), b2 + 4 * (n2 == 1 ? 0 : e));
}

constexpr int kNms64Threads = 1024;

// rows: [n, 6] float64 (xmin, ymin, w, h, score, class).  classes: ascending list of the distinct class values.
__global__ void __launch_bounds__(kNms64Threads) centernet_nms_kernel(const double* __restrict__ rows, int n, const double* __restrict__ classes,
                                                                      int n_classes, double thr, double sigma, int soft,
                                                                      double* __restrict__ score /*[n] scratch*/, double* __restrict__ out,
                                                                      int* __restrict__ out_src, int* __restrict__ n_out) {
    __shared__ double s_val[kNms64Threads / 32];
    __shared__ int s_idx[kNms64Threads / 32];
    __shared__ int s_best;
    __shared__ double s_box[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int emitted = 0;
    for (int c = 0; c < n_classes; ++c) {
        const double cls = classes[c];
        // the pool: rows of this class, scores copied to scratch (negative infinity marks "not in the pool")
        for (int i = tid; i < n; i += kNms64Threads) score[i] = rows[6 * i + 5] == cls ? rows[6 * i + 4] : -INFINITY;
        __syncthreads();
        // note: -inf also stands for rows removed later; a genuine -inf score cannot be told apart (out of contract)
        while (true) {
            // argmax over the pool, first index on ties
            double bv = -INFINITY;
            int bi = 0x7fffffff;
            bool any = false;
            for (int i = tid; i < n; i += kNms64Threads) {
                const double v = score[i];
                if (v == -INFINITY) continue;
                if (!any || v > bv) bv = v, bi = i, any = true;
            }
            if (!any) bi = 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) bv = ov, bi = oi;
            }
            if (lane == 0) s_val[warp] = bv, s_idx[warp] = bi;
            __syncthreads();
            if (warp == 0) {
                bv = s_val[lane], bi = s_idx[lane];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) bv = ov, bi = oi;
                }
                if (lane == 0) {
                    s_best = bi;
                    if (bi != 0x7fffffff) {
                        const double* r = rows + 6 * bi;
                        s_box[0] = r[0], s_box[1] = r[1], s_box[2] = dadd(r[0], r[2]), s_box[3] = dadd(r[1], r[3]);  // :53-55
                        double* o = out + 6 * emitted;
                        o[0] = s_box[0], o[1] = s_box[1], o[2] = s_box[2], o[3] = s_box[3], o[4] = score[bi], o[5] = cls;
                        out_src[emitted] = bi;
                        score[bi] = -INFINITY;  // popped
                    }
                }
            }
            __syncthreads();
            const int best = s_best;
            if (best == 0x7fffffff) break;
            ++emitted;
            // re-weight the rest of the pool (:66-82)
            for (int i = tid; i < n; i += kNms64Threads) {
                const double v = score[i];
                if (v == -INFINITY) continue;
                const double* r = rows + 6 * i;
                const double bx[4] = {r[0], r[1], dadd(r[0], r[2]), dadd(r[1], r[3])};
                const double iou = iou64(s_box, bx);
                double w;
                if (soft)
                    w = exp(-ddiv(dmul(1.0, dmul(iou, iou)), sigma));
                else
                    w = iou > thr ? 0.0 : 1.0;
                const double nv = dmul(v, w);
                score[i] = nv > 0.0 ? nv : -INFINITY;  // `score > 0` keeps a row in the pool
            }
            __syncthreads();
        }
        __syncthreads();
    }
    if (tid == 0) *n_out = emitted;
}

}  // namespace dh

using namespace dh;

extern "C" {

int dh_compute_iou(dh_handle_t h, const float* boxes1, int n, const float* boxes2, int m, float* out, void* stream) {
    DH_CHECK_ARG(h && (n == 0 || boxes1) && (m == 0 || boxes2) && out, "dh_compute_iou: NULL argument");
    DH_CHECK_ARG(n >= 0 && m >= 0, "dh_compute_iou: bad sizes");
    const long long total = static_cast<long long>(n) * m;
    if (total == 0) return DH_OK;
    DeviceGuard guard(h);
    long long g = (total + 255) / 256;
    if (g > static_cast<long long>(h->sm_count) * 16) g = static_cast<long long>(h->sm_count) * 16;
    compute_iou_kernel<<<static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(boxes1, n, boxes2, m, out);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

int dh_bboxes_iou(dh_handle_t h, const double* boxes1, int n1, const double* boxes2, int n2, double* out, void* stream) {
    DH_CHECK_ARG(h && boxes1 && boxes2 && out, "dh_bboxes_iou: NULL argument");
    DH_CHECK_ARG(n1 >= 0 && n2 >= 0 && (n1 == n2 || n1 == 1 || n2 == 1), "dh_bboxes_iou: shapes %d and %d do not broadcast", n1, n2);
    const int n = (n1 == 0 || n2 == 0) ? 0 : (n1 > n2 ? n1 : n2);
    if (n == 0) return DH_OK;
    DeviceGuard guard(h);
    bboxes_iou_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(boxes1, n1, boxes2, n2, out);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

int dh_centernet_nms(dh_handle_t h, const double* rows, int n, const double* classes, int n_classes, double iou_threshold,
                     double sigma, int soft, double* out_rows, int32_t* out_src, int32_t* n_out, void* stream) {
    DH_CHECK_ARG(h && out_rows && out_src && n_out, "dh_centernet_nms: NULL argument");
    DH_CHECK_ARG(n >= 0 && n_classes >= 0 && (n == 0 || (rows && classes)), "dh_centernet_nms: bad arguments");
    DeviceGuard guard(h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0 || n_classes == 0) {
        DH_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), st));
        return DH_OK;
    }
    double* score = static_cast<double*>(scratch(h, static_cast<size_t>(n) * 8));
    if (!score) return DH_ERR_CUDA;
    centernet_nms_kernel<<<1, kNms64Threads, 0, st>>>(rows, n, classes, n_classes, iou_threshold, sigma, soft, score, out_rows, out_src, n_out);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

}  // extern "C"
