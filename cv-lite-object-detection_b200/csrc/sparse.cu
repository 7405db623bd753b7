// The offline COCO -> sparse FCOS target formatter of the reference, /format_COCO_annotations_fcos.py:66-183, for a
// batch of images in one call: per object the footprint rectangle on the img_dims canvas, per footprint cell seven
// COO entries ([y, x, scale, k] = b, t, l, r, centerness, 1 for k = 0..5 and the class entry [y, x, scale, label + 4]
// = 1), in the script's order (objects in input order, cells x-major, entries as listed).
//
// The output is one long write (20 bytes per entry, 140 per cell), so the shape is: (1) one thread per object derives its
// rectangle and cell count (float64 like the script: the divisions by the ratios and the truncations decide indices);
// (2) one CTA turns the counts into offsets (objects in order -> the order of the entries is fixed, no atomics);
// (3) a grid-wide pass where every warp takes 32 consecutive cells, finds their objects by binary search in the
// offsets (L2-resident), builds the 224 entries in shared memory and writes them out as consecutive 16-byte index
// vectors and consecutive floats -- fully coalesced stores, each byte written once.
#include "dh_common.cuh"
#include "dh_host.h"

namespace dh {

struct SparseObj {  // what the emit pass needs of one object
    long long x_low, x_upp, y_low, y_upp;  // as the script computes them (l, r, b, t are measured from these)
    int x0, y0, ny;                        // first cell and cells per column after NumPy's slice rules
    int scale, label;
};

// NumPy basic-slice normalisation of [lo:hi] against an axis of n: negative bounds count from the end, then both are
// clipped to [0, n]
__device__ __forceinline__ void slice_bounds(long long lo, long long hi, int n, int& first, int& count) {
    if (lo < 0) lo += n;
    if (hi < 0) hi += n;
    lo = lo < 0 ? 0 : (lo > n ? n : lo);
    hi = hi < 0 ? 0 : (hi > n ? n : hi);
    first = static_cast<int>(lo);
    count = hi > lo ? static_cast<int>(hi - lo) : 0;
}
// Python's int() of a float64 (truncation), kept inside +-2^40 (a box that far out has no cells on any canvas)
__device__ __forceinline__ long long trunc_ll(double v) {
    const double lim = 1099511627776.0;
    return static_cast<long long>(fmin(fmax(trunc(v), -lim), lim));
}

__global__ void sparse_count_kernel(const double* __restrict__ boxes, const int* __restrict__ nbox, const double* __restrict__ src_dims,
                                    int batch, int max_boxes, int img_w, int img_h, int num_scale, int min_side,
                                    SparseObj* __restrict__ objs, long long* __restrict__ cells) {
    const int total = batch * max_boxes;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gridDim.x * blockDim.x) {
        const int b = o / max_boxes, k = o - b * max_boxes;
        SparseObj rec;
        rec.x_low = rec.x_upp = rec.y_low = rec.y_upp = 0, rec.x0 = rec.y0 = rec.ny = 0, rec.scale = 0, rec.label = 0;
        long long n_cells = 0;
        const int n = nbox ? min(max(nbox[b], 0), max_boxes) : max_boxes;
        if (k < n) {
            const double* g = boxes + static_cast<long long>(o) * 5;
            const double w_ratio = ddiv(src_dims[2 * b], static_cast<double>(img_w));      // :86
            const double h_ratio = ddiv(src_dims[2 * b + 1], static_cast<double>(img_h));  // :87
            const double width = ddiv(g[2], w_ratio), height = ddiv(g[3], h_ratio);        // :98-99
            if (width >= 0.0 && height >= 0.0 && isfinite(width) && isfinite(height) && isfinite(g[0]) && isfinite(g[1])) {
                int sc = num_scale - 1;  // :103-123: the first slot both sides are strictly below; the last one unconditionally
                for (int s = 0; s < num_scale - 1; ++s) {
                    const double lim = static_cast<double>(static_cast<int>(static_cast<double>(min_side) / static_cast<double>(1 << (num_scale - 1 - s))));
                    if (width < lim && height < lim) {
                        sc = s;
                        break;
                    }
                }
                rec.x_low = trunc_ll(ddiv(g[0], w_ratio));                          // :126-127
                rec.x_upp = trunc_ll(dadd(static_cast<double>(rec.x_low), width));  // :128
                rec.y_low = trunc_ll(ddiv(g[1], h_ratio));
                rec.y_upp = trunc_ll(dadd(static_cast<double>(rec.y_low), height));
                int nx;
                slice_bounds(rec.x_low, rec.x_upp, img_w, rec.x0, nx);  // :147-149
                slice_bounds(rec.y_low, rec.y_upp, img_h, rec.y0, rec.ny);
                rec.scale = sc;
                rec.label = static_cast<int>(g[4]);
                n_cells = static_cast<long long>(nx) * rec.ny;
            }
        }
        objs[o] = rec;
        cells[o] = n_cells;
    }
}

// exclusive scan of cells[0..n) in place (one CTA); cells[n] = total; image offsets in entries
__global__ void __launch_bounds__(1024) sparse_scan_kernel(long long* __restrict__ cells, int n, int batch, int max_boxes,
                                                           long long* __restrict__ out_offsets) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const long long v = i < n ? cells[i] : 0;
        long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += u;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_tot[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long u = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += u;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const long long before = carry + (warp ? warp_tot[warp - 1] : 0) + inc - v;
        if (i < n) {
            cells[i] = before;
            if (i % max_boxes == 0) out_offsets[i / max_boxes] = before * 7;
        }
        __syncthreads();
        if (tid == 1023) carry = before + v;
        __syncthreads();
    }
    if (tid == 0) {
        cells[n] = carry;
        out_offsets[batch] = carry * 7;
    }
}

constexpr int kSparseWarps = 8;

__global__ void __launch_bounds__(kSparseWarps * 32) sparse_emit_kernel(const SparseObj* __restrict__ objs, const long long* __restrict__ cell_off,
                                                                        int n_obj, long long capacity, int4* __restrict__ out_idx,
                                                                        float* __restrict__ out_val) {
    __shared__ __align__(16) int4 s_idx[kSparseWarps][224];
    __shared__ float s_val[kSparseWarps][224];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long total = cell_off[n_obj];
    const long long n_groups = (total + 31) >> 5;
    for (long long grp = static_cast<long long>(blockIdx.x) * kSparseWarps + warp; grp < n_groups;
         grp += static_cast<long long>(gridDim.x) * kSparseWarps) {
        const long long cell = (grp << 5) + lane;
        if (cell < total) {
            // the object whose range [cell_off[o], cell_off[o+1]) holds this cell: the last o with cell_off[o] <= cell
            int lo = 0, hi = n_obj - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (__ldg(cell_off + mid) <= cell) lo = mid;
                else hi = mid - 1;
            }
            const SparseObj ob = objs[lo];
            const int local = static_cast<int>(cell - __ldg(cell_off + lo));
            const int cx = local / ob.ny;  // x-major: np.nonzero walks the [x, y, scale] array in row-major order (:151)
            const int x = ob.x0 + cx, y = ob.y0 + (local - cx * ob.ny);
            const long long l = x - ob.x_low, r = ob.x_upp - x, bb = y - ob.y_low, t = ob.y_upp - y;  // :158-161
            // compute_centerness (:8-11): int64 / int64 true division, sqrt and product in float64
            const double c = dmul(sqrt(ddiv(static_cast<double>(l < r ? l : r), static_cast<double>(l < r ? r : l))),
                                  sqrt(ddiv(static_cast<double>(bb < t ? bb : t), static_cast<double>(bb < t ? t : bb))));
            const float v[7] = {static_cast<float>(bb), static_cast<float>(t), static_cast<float>(l), static_cast<float>(r),
                                static_cast<float>(c), 1.0f, 1.0f};
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                s_idx[warp][lane * 7 + k] = make_int4(y, x, ob.scale, k < 6 ? k : ob.label + 4);  // :164-172
                s_val[warp][lane * 7 + k] = v[k];
            }
        }
        __syncwarp();
        const long long e0 = (grp << 5) * 7;
        const long long n_ent = min(static_cast<long long>(224), total * 7 - e0);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int e = k * 32 + lane;
            if (e < n_ent && e0 + e < capacity) {
                out_idx[e0 + e] = s_idx[warp][e];
                out_val[e0 + e] = s_val[warp][e];
            }
        }
        __syncwarp();
    }
}

}  // namespace dh

using namespace dh;

extern "C" {

int dh_fcos_sparse_encode(dh_handle_t h, const double* boxes, const int32_t* nbox, const double* src_dims, int batch, int max_boxes,
                          int img_width, int img_height, int num_scale, long long capacity, int32_t* out_indices, float* out_values,
                          long long* out_offsets, void* stream) {
    DH_CHECK_ARG(h && boxes && src_dims && out_offsets, "dh_fcos_sparse_encode: NULL argument");
    DH_CHECK_ARG(batch >= 0 && max_boxes >= 1 && img_width >= 1 && img_height >= 1 && num_scale >= 1 && num_scale <= 16,
                 "dh_fcos_sparse_encode: bad sizes");
    DH_CHECK_ARG(static_cast<long long>(batch) * max_boxes < (1ll << 30), "dh_fcos_sparse_encode: too many objects");
    DH_CHECK_ARG((out_indices == nullptr) == (out_values == nullptr) && capacity >= 0, "dh_fcos_sparse_encode: indices / values / capacity");
    DH_CHECK_ARG((reinterpret_cast<uintptr_t>(out_indices) & 15u) == 0, "dh_fcos_sparse_encode: out_indices must be 16-byte aligned");
    DeviceGuard guard(h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (batch == 0) {
        DH_CUDA(cudaMemsetAsync(out_offsets, 0, sizeof(long long), st));
        return DH_OK;
    }
    const int n_obj = batch * max_boxes;
    const size_t obj_bytes = (sizeof(SparseObj) * n_obj + 255) & ~size_t(255);
    char* sc = static_cast<char*>(scratch(h, obj_bytes + sizeof(long long) * (n_obj + 1)));
    if (!sc) return DH_ERR_CUDA;
    SparseObj* objs = reinterpret_cast<SparseObj*>(sc);
    long long* cells = reinterpret_cast<long long*>(sc + obj_bytes);
    const int min_side = img_width < img_height ? img_width : img_height;
    const int grid = (n_obj + 255) / 256 < h->sm_count * 4 ? (n_obj + 255) / 256 : h->sm_count * 4;
    sparse_count_kernel<<<grid, 256, 0, st>>>(boxes, nbox, src_dims, batch, max_boxes, img_width, img_height, num_scale, min_side, objs, cells);
    DH_CUDA(cudaGetLastError());
    sparse_scan_kernel<<<1, 1024, 0, st>>>(cells, n_obj, batch, max_boxes, out_offsets);
    DH_CUDA(cudaGetLastError());
    h->launches += 2;
    if (out_indices && capacity > 0) {
        // every entry is written once; the grid is sized for the capacity (the entry count itself stays on the device)
        const long long groups = (capacity / 7 + 31) / 32 + 1;
        const long long want = (groups + kSparseWarps - 1) / kSparseWarps;
        const int g = static_cast<int>(want < h->sm_count * 8ll ? (want < 1 ? 1 : want) : h->sm_count * 8ll);
        sparse_emit_kernel<<<g, kSparseWarps * 32, 0, st>>>(objs, cells, n_obj, capacity, reinterpret_cast<int4*>(out_indices), out_values);
        DH_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    return DH_OK;
}

}  // extern "C"
