// Inference-time decode: head outputs -> pixel boxes + scores, and the pre-NMS candidate selection.
//
//   dh_prediction_to_corners   the four prediction_to_corners variants of the reference
//                              (FCOS/fcos.py:112-134, fcos_center_v1.py:125-147,
//                               retinanet_module.py:428-451, tf_centernet_resnet_s8.py:210-241)
//   dh_fcos_decode             what FCOS/infer_fcos.py:35-57 hands to the NMS op: boxes [B,N,4], scores [B,N,C]
//   dh_retina_decode           RetinaNet.image_detections up to the threshold (retinanet_module.py:487-520):
//                              dets [B,N,6] = (y1, x1, y2, x2, max score, argmax label), level > anchor > row-major
//   dh_select_topk             per-segment (= per pyramid level) score threshold + exact top-k, stable in index
//                              order (linear 4096-bin histogram, boundary bin ranked exactly, ordered compaction)
//   launch_fcos_select         (dh_fcos_detect) the same selection fused with the FCOS decode, on raw logits: three
//                              exact, bit-identical variants -- estimate + one all-SM streaming pass + per-segment
//                              finish, a thread-block cluster per long segment, one CTA per segment
//
// These kernels are HBM-bound on one read of the head output; the decode kernels map one thread to an output
// row with the class loop vectorised where the row is 16-byte aligned.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>

#include "dh_common.cuh"
#include "dh_host.h"
#include "dh_infer.h"

namespace dh {

__device__ __forceinline__ float sigmoid_acc(float x) { return fdiv(1.0f, fadd(1.0f, expf(-x))); }
// Score of one (location, class) pair from the class and centerness channels of its head row.  `center` is the
// DH_FCOS_SCORE_* mode: 0 sigmoid(class), 1 sigmoid(centerness) * sigmoid(class) (FCOS/infer_fcos.py:44-51); 2 / 3 take
// the channels as probabilities -- a TARGET map fed back through the detector (show_heatmap,
// FCOS/train_fcos_center_voc.py:54-66): 2 the class value, 3 sqrt(class * centerness) in float64 like numpy.
__device__ __forceinline__ float fcos_pair_score(int center, float cls, float cen) {
    if (center == 2) return cls;
    if (center == 3) return static_cast<float>(sqrt(static_cast<double>(cls) * static_cast<double>(cen)));
    const float s = sigmoid_acc(cls);
    return center ? fmul(sigmoid_acc(cen), s) : s;
}

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
        const float cen = center ? q[4] : 0.0f;
        for (int c = lane; c < num_classes; c += 32) scores[orow * num_classes + c] = fcos_pair_score(center, q[5 + c], cen);
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

// ---- RetinaNet decode, streaming version ---------------------------------------------------------------------
// HBM-bound on one read of the head (25.8 MB per COCO image).  Each warp owns 32-row tiles: one elected lane
// pulls the tile (32 x ch floats, contiguous) into the warp's shared-memory stage with a single TMA bulk copy,
// two stages deep, and every lane then scans ITS row out of shared memory.  max_c sigmoid(x_c) is monotone in the
// logit, so the scan is a plain max over the logits; sigmoid is evaluated only for the winner and for the few
// logits close enough to it that their float32 sigmoid could collide with the winner's (the reference's
// np.argmax returns the FIRST index of the maximal score, so such a collision at a lower index must win).
constexpr int kDecWarps = 8;
struct RetinaDecodeArgs {
    const float* head[DH_MAX_LEVELS];   // [B, A, Hl, Wl, ch]
    long long rows[DH_MAX_LEVELS];      // B * A * Hl * Wl
    long long tile_begin[DH_MAX_LEVELS + 1];
    long long level_off[DH_MAX_LEVELS]; // first output row of the level inside an image
    int per_img[DH_MAX_LEVELS];         // A * Hl * Wl
    int cells[DH_MAX_LEVELS], wl[DH_MAX_LEVELS];
    FastDiv div_per_img[DH_MAX_LEVELS], div_cells[DH_MAX_LEVELS], div_wl[DH_MAX_LEVELS];
    float stride[DH_MAX_LEVELS];
    const float* anchor_hw;             // device [n_levels, A, 2]
    int n_levels, n_anchors, num_classes, ch, use_tma;
    long long n_total;
    float* scores;                      // optional [B, n_total]: the score column alone, for the selector that follows (dh_retina_detect)
};

// lower bound t such that x < t implies sigmoid_acc(x) < sigmoid_acc(m) strictly (conservative; see above)
__device__ __forceinline__ float sigmoid_collision_floor(float m) {
    const float e = expf(-m);
    if (!(e > 0.f)) return fminf(16.0f, m);  // exp underflowed: every logit >= ~16.7 gives exactly 1
    const float delta = 2.0f * log1pf(9.5367431640625e-07f * (1.0f + 1.0f / e)) + 1.0e-6f * fabsf(m);
    const float t = m - delta;
    return m > 16.0f ? fminf(16.0f, t) : t;
}

__global__ void __launch_bounds__(kDecWarps * 32) retina_decode_stream_kernel(const __grid_constant__ RetinaDecodeArgs a, float* __restrict__ dets) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = a.ch, C = a.num_classes;
    const int stage_floats = 32 * ch;
    float* stage0 = reinterpret_cast<float*>(smem) + static_cast<long long>(warp) * 2 * stage_floats;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(kDecWarps) * 2 * stage_floats * 4) + warp * 2;
    if (lane == 0) {
        mbar_init(bars, 1), mbar_init(bars + 1, 1);
        mbar_init_fence();
    }
    __syncwarp();
    const long long total_tiles = a.tile_begin[a.n_levels];
    const long long gw = static_cast<long long>(blockIdx.x) * kDecWarps + warp;
    const long long nw = static_cast<long long>(gridDim.x) * kDecWarps;
    auto locate = [&](long long tile, int& l, long long& r0, int& nrows) {
        l = 0;
        while (l + 1 < a.n_levels && tile >= a.tile_begin[l + 1]) ++l;
        r0 = (tile - a.tile_begin[l]) * 32;
        const long long left = a.rows[l] - r0;
        nrows = left < 32 ? static_cast<int>(left) : 32;
    };
    auto fetch = [&](long long tile, int st) {  // all lanes call
        int l, nrows;
        long long r0;
        locate(tile, l, r0, nrows);
        const float* src = a.head[l] + r0 * ch;
        float* dst = stage0 + st * stage_floats;
        if (a.use_tma) {
            if (lane == 0) {
                const uint32_t bytes = static_cast<uint32_t>(nrows) * ch * 4u;
                mbar_expect_tx(bars + st, bytes);
                bulk_g2s(dst, src, bytes, bars + st);
            }
        } else {
            for (int e = lane; e < nrows * ch; e += 32) dst[e] = __ldg(src + e);
        }
    };
    uint32_t parity[2] = {0u, 0u};
    long long tile = gw;
    if (tile < total_tiles) fetch(tile, 0);
    for (int it = 0; tile < total_tiles; ++it, tile += nw) {
        const int st = it & 1;
        if (tile + nw < total_tiles) fetch(tile + nw, st ^ 1);  // the other stage was drained before the __syncwarp below
        if (a.use_tma) {
            mbar_wait(bars + st, parity[st]);
            parity[st] ^= 1u;
        } else {
            __syncwarp();
        }
        int l, nrows;
        long long r0;
        locate(tile, l, r0, nrows);
        float* tile_smem = stage0 + st * stage_floats;
        float o6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        long long orow = 0;
        if (lane < nrows) {
            const float* q = tile_smem + lane * ch;
            // max logit, its first index, and the largest logit before that index
            float m = q[4], prev = -INFINITY;
            int cm = 0;
            if ((ch & 3) == 0) {
                const float4* q4 = reinterpret_cast<const float4*>(q);
                for (int v = 1; v < (ch >> 2); ++v) {
                    const float4 x = q4[v];
                    const int c0 = 4 * v - 4;
                    if (x.x > m) prev = m, m = x.x, cm = c0;
                    if (x.y > m) prev = m, m = x.y, cm = c0 + 1;
                    if (x.z > m) prev = m, m = x.z, cm = c0 + 2;
                    if (x.w > m) prev = m, m = x.w, cm = c0 + 3;
                }
            } else {
                for (int c = 1; c < C; ++c)
                    if (q[4 + c] > m) prev = m, m = q[4 + c], cm = c;
            }
            const float best = sigmoid_acc(m);
            int label = cm;
            if (best == 0.f) {
                label = 0;  // every score is 0: the first index wins
            } else if (cm > 0) {
                // below 8 the collision window is < 6e-3 wide; the general bound is only evaluated for large logits
                const float floor_x = m < 8.0f ? m - 0.01f : sigmoid_collision_floor(m);
                if (prev >= floor_x) {  // rare: an earlier logit is close enough to tie after the sigmoid
                    for (int c = 0; c < cm; ++c) {
                        const float x = q[4 + c];
                        if (x >= floor_x && sigmoid_acc(x) == best) {
                            label = c;
                            break;
                        }
                    }
                }
            }
            const uint32_t r = static_cast<uint32_t>(r0) + lane;
            const uint32_t b = fdiv_u32(r, a.div_per_img[l]);
            const uint32_t local = r - b * a.per_img[l];
            const uint32_t an = fdiv_u32(local, a.div_cells[l]);
            const uint32_t loc = local - an * a.cells[l];
            const uint32_t i = fdiv_u32(loc, a.div_wl[l]);
            CornerParams cp;
            cp.mode = 1, cp.stride = a.stride[l];
            cp.d0 = __ldg(a.anchor_hw + (l * a.n_anchors + an) * 2), cp.d1 = __ldg(a.anchor_hw + (l * a.n_anchors + an) * 2 + 1);
            const float4 o = corners_of(cp, static_cast<int>(i), static_cast<int>(loc - i * a.wl[l]), 0, q[0], q[1], q[2], q[3]);
            o6[0] = o.x, o6[1] = o.y, o6[2] = o.z, o6[3] = o.w, o6[4] = best, o6[5] = static_cast<float>(label);
            orow = static_cast<long long>(b) * a.n_total + a.level_off[l] + local;
            if (a.scores) a.scores[orow] = best;
        }
        // hand the 6-float rows to the stage buffer and write them out coalesced (the tile's output rows are
        // contiguous unless it straddles an image boundary)
        __syncwarp();
        const long long orow0 = __shfl_sync(0xffffffffu, orow, 0);
        const bool contiguous = __all_sync(0xffffffffu, lane >= nrows || orow == orow0 + lane);
        if (contiguous) {
            if (lane < nrows) {
#pragma unroll
                for (int c = 0; c < 6; ++c) tile_smem[lane * 6 + c] = o6[c];
            }
            __syncwarp();
            float* dst = dets + orow0 * 6;
            for (int e = lane; e < nrows * 6; e += 32) dst[e] = tile_smem[e];
        } else if (lane < nrows) {
            float* d = dets + orow * 6;
#pragma unroll
            for (int c = 0; c < 6; ++c) d[c] = o6[c];
        }
        __syncwarp();  // every lane is done with this stage before it is refilled
    }
}

// ---- exact per-segment top-k, stable in index order ------------------------------------------------------
// One CTA per (segment, image).  The k-th best passing score is located with three reads of the segment:
//   1. a 4096-bin histogram over a monotone *linear* map of the score (sigmoid scores spread over thousands of
//      bins; a radix histogram of the float bits would pile them onto ~40 exponent bins and serialise the atomics);
//   2. the entries of the boundary bin are collected into shared memory and ranked exactly by (score desc, index
//      asc); the entry of rank need-1 gives the cut (T_key, T_idx).  (A boundary bin too large for shared memory
//      is first narrowed by radix rounds on the float bits, and a run of identical scores by an ordered count.)
//   3. an ordered compaction: `take = key > T_key || (key == T_key && index <= T_idx)` is a pure per-element
//      predicate, so each thread handles 8 consecutive rows and a chunk of 8192 rows costs one block barrier.
__device__ __forceinline__ unsigned score_key(float s) {
    const unsigned u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone float -> uint
}
__device__ __forceinline__ int linear_bin(float s) {  // monotone non-decreasing in s
    return s <= 0.f ? 0 : (s >= 1.f ? 4095 : static_cast<int>(s * 4096.0f));
}

constexpr int kSelThreads = 1024;
constexpr int kSelItems = 8;     // consecutive rows per thread in the compaction pass
constexpr int kSelListCap = 2048;  // boundary-bin entries ranked in shared memory

struct SelShared {
    unsigned hist[4096];
    unsigned list_key[kSelListCap];
    int list_idx[kSelListCap];
    unsigned wtot[2][kSelThreads / 32];
    unsigned n_list, bin, need, above, t_key, n_pass;
    int t_idx;
};

// largest bin b with sum(hist[b..nbins)) >= need; returns the bin and the count strictly above it (block-wide)
template <int T = kSelThreads>
__device__ __forceinline__ void find_boundary(SelShared& sh, int nbins, unsigned need, int tid) {
    // thread t owns bins [t*per, t*per+per) counted from the top
    const int per = (nbins + T - 1) / T;
    unsigned mine = 0;
    for (int q = 0; q < per; ++q) {
        const int b = nbins - 1 - (tid * per + q);
        if (b >= 0) mine += sh.hist[b];
    }
    // inclusive scan over threads (descending bins)
    const int lane = tid & 31, warp = tid >> 5;
    unsigned incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) sh.wtot[0][warp] = incl;
    __syncthreads();
    unsigned before = 0;
    for (int w = 0; w < warp; ++w) before += sh.wtot[0][w];
    const unsigned excl = before + incl - mine;
    if (excl < need && excl + mine >= need) {  // the boundary lies in this thread's bins (exactly one thread)
        unsigned cum = excl;
        for (int q = 0; q < per; ++q) {
            const int b = nbins - 1 - (tid * per + q);
            if (b < 0) break;
            if (cum + sh.hist[b] >= need) {
                sh.bin = static_cast<unsigned>(b), sh.above = cum;
                break;
            }
            cum += sh.hist[b];
        }
    }
    __syncthreads();
}

// The selector proper, shared by the generic kernel (scores = one column of a row array) and the fused FCOS
// kernel (scores computed on the fly from the head logits).  `Src` provides
//   int n                               elements in the segment (indices 0 .. n-1)
//   float score(int i)                  the element's score (deterministic: it is evaluated up to three times)
//   void emit(int i, float s, int rank) write output slot `rank` from element i
//   void pad(int rank)                  mark output slot `rank` unused
// Selected elements are emitted in index order; exactly min(k, #passing) slots are filled.
template <class Src, int T = kSelThreads>
__device__ __forceinline__ void select_core(SelShared& sh, const Src& src, int k_slots, float min_score, int inclusive) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = src.n;
    const int k = min(k_slots, max(n, 0));  // a segment shorter than k: the slots beyond its length stay padded
    auto passes = [&](float s) { return inclusive ? (s >= min_score) : (s > min_score); };
    constexpr int U = 8;  // independent loads in flight per thread in the unordered passes

    // ---- 1. linear histogram -------------------------------------------------------------------------------
    for (int i = tid; i < 4096; i += T) sh.hist[i] = 0;
    if (tid == 0) sh.n_list = 0, sh.bin = 0xFFFFFFFFu, sh.above = 0, sh.n_pass = 0;
    __syncthreads();
    unsigned my_pass = 0;
    for (int base = 0; base < n; base += T * U) {
        float s[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * T + tid;
            s[u] = i < n ? src.score(i) : __int_as_float(0x7fc00000);  // NaN never passes
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (passes(s[u])) atomicAdd(&sh.hist[linear_bin(s[u])], 1u), ++my_pass;
    }
    my_pass = static_cast<unsigned>(warp_sum_i(static_cast<int>(my_pass)));
    if (lane == 0 && my_pass) atomicAdd(&sh.n_pass, my_pass);
    __syncthreads();
    unsigned T_key = 0u;  // default: fewer than k rows pass -> take everything that passes
    int T_idx = 0x7fffffff;
    if (k > 0) find_boundary<T>(sh, 4096, static_cast<unsigned>(k), tid);
    if (k > 0 && sh.bin != 0xFFFFFFFFu) {
        const int Bk = static_cast<int>(sh.bin);
        unsigned need = static_cast<unsigned>(k) - sh.above;  // still to take inside the boundary bin (>= 1)
        unsigned cnt = sh.hist[Bk];
        // ---- 2a. (rare) boundary bin larger than the shared-memory list: narrow it by radix rounds on the key bits
        unsigned prefix = 0u, decided = 0u;
        bool all_equal = false;
        for (int round = 0; cnt > kSelListCap && round < 3; ++round) {
            const int shift = round == 0 ? 20 : (round == 1 ? 8 : 0);
            const int bits = round == 2 ? 8 : 12;
            __syncthreads();
            for (int i = tid; i < (1 << bits); i += T) sh.hist[i] = 0;
            if (tid == 0) sh.bin = 0xFFFFFFFFu;
            __syncthreads();
            for (int i = tid; i < n; i += T) {
                const float s = src.score(i);
                if (!passes(s) || linear_bin(s) != Bk) continue;
                const unsigned key = score_key(s);
                if ((key & decided) == prefix) atomicAdd(&sh.hist[(key >> shift) & ((1u << bits) - 1u)], 1u);
            }
            __syncthreads();
            find_boundary<T>(sh, 1 << bits, need, tid);
            prefix |= sh.bin << shift;
            decided |= ((1u << bits) - 1u) << shift;
            need -= sh.above;
            cnt = sh.hist[sh.bin];
            if (round == 2) all_equal = true;  // every key bit decided: the remaining rows share one score
        }
        if (cnt > kSelListCap && all_equal) {
            // ---- 2b. (rarer) more identical scores than the list holds: the cut is the need-th of them in index order
            T_key = prefix;
            unsigned carry = 0;
            for (int base = 0; base < n; base += T) {
                const int i = base + tid;
                bool eq = false;
                if (i < n) {
                    const float s = src.score(i);
                    eq = passes(s) && score_key(s) == prefix;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, eq);
                if (lane == 0) sh.wtot[0][warp] = __popc(bal);
                __syncthreads();
                unsigned before = carry, tot = 0;
                for (int w = 0; w < T / 32; ++w) {
                    const unsigned v = sh.wtot[0][w];
                    if (w < warp) before += v;
                    tot += v;
                }
                const unsigned rank = before + __popc(bal & ((1u << lane) - 1u));
                if (eq && rank == need - 1) sh.t_idx = i;
                carry += tot;
                __syncthreads();
            }
            T_idx = sh.t_idx;
        } else {
            // ---- 2. collect the boundary entries, rank them exactly ------------------------------------------------
            __syncthreads();
            for (int base = 0; base < n; base += T * U) {
                float s[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + u * T + tid;
                    s[u] = i < n ? src.score(i) : __int_as_float(0x7fc00000);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (!passes(s[u]) || linear_bin(s[u]) != Bk) continue;
                    const unsigned key = score_key(s[u]);
                    if ((key & decided) != prefix) continue;
                    const unsigned slot = atomicAdd(&sh.n_list, 1u);
                    sh.list_key[slot] = key, sh.list_idx[slot] = base + u * T + tid;
                }
            }
            __syncthreads();
            const int n_list = static_cast<int>(sh.n_list);
            for (int e = tid; e < n_list; e += T) {
                const unsigned ke = sh.list_key[e];
                const int ie = sh.list_idx[e];
                unsigned rank = 0;
                for (int f = 0; f < n_list; ++f) {
                    const unsigned kf = sh.list_key[f];
                    rank += (kf > ke || (kf == ke && sh.list_idx[f] < ie)) ? 1u : 0u;
                }
                if (rank == need - 1) sh.t_key = ke, sh.t_idx = ie;
            }
            __syncthreads();
            T_key = sh.t_key, T_idx = sh.t_idx;
        }
    }
    __syncthreads();

    // ---- 3. ordered compaction ---------------------------------------------------------------------------------
    unsigned carry = 0;
    int it = 0;
    for (int base = 0; base < n; base += T * kSelItems, ++it) {
        const int i0 = base + tid * kSelItems;
        float s[kSelItems];
#pragma unroll
        for (int u = 0; u < kSelItems; ++u) s[u] = (i0 + u < n) ? src.score(i0 + u) : __int_as_float(0x7fc00000);
        unsigned flags = 0;
#pragma unroll
        for (int u = 0; u < kSelItems; ++u) {
            if (passes(s[u])) {
                const unsigned key = score_key(s[u]);
                if (key > T_key || (key == T_key && i0 + u <= T_idx)) flags |= 1u << u;
            }
        }
        const unsigned mine = __popc(flags);
        unsigned incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        unsigned* wt = sh.wtot[it & 1];  // parity double buffer: one barrier per chunk
        if (lane == 31) wt[warp] = incl;
        __syncthreads();
        unsigned before = carry, tot = 0;
        for (int w = 0; w < T / 32; ++w) {
            const unsigned v = wt[w];
            if (w < warp) before += v;
            tot += v;
        }
        unsigned rank = before + incl - mine;
#pragma unroll
        for (int u = 0; u < kSelItems; ++u) {
            if (!((flags >> u) & 1u)) continue;
            if (rank < static_cast<unsigned>(k)) src.emit(i0 + u, s[u], static_cast<int>(rank));
            ++rank;
        }
        carry += tot;
    }
    // pad the unused slots so that any threshold drops them
    const unsigned filled = min(static_cast<unsigned>(k), carry);
    for (int r = filled + tid; r < k_slots; r += T) src.pad(r);
}

// generic source: rows of `row_floats` floats, the score in column `score_col`; selected rows are copied whole
struct RowSource {
    const float* rows;  // first row of the segment
    const float* scores;  // or null: the score column on its own (the three passes then read 4 bytes per row, not a 32-byte sector)
    float* out;         // first output slot of the segment
    int* out_src;       // or null
    int row_floats, score_col, n, first_index;
    __device__ __forceinline__ float score(int i) const {
        return scores ? __ldg(scores + i) : __ldg(rows + static_cast<long long>(i) * row_floats + score_col);
    }
    __device__ __forceinline__ void emit(int i, float, int rank) const {
        const float* src = rows + static_cast<long long>(i) * row_floats;
        float* dst = out + static_cast<long long>(rank) * row_floats;
        for (int c = 0; c < row_floats; ++c) dst[c] = src[c];
        if (out_src) out_src[rank] = first_index + i;
    }
    __device__ __forceinline__ void pad(int rank) const {
        float* dst = out + static_cast<long long>(rank) * row_floats;
        for (int c = 0; c < row_floats; ++c) dst[c] = 0.f;
        dst[score_col] = -INFINITY;
        if (out_src) out_src[rank] = -1;
    }
};

struct SelSegs {
    int off[DH_MAX_LEVELS + 1];
};
__global__ void __launch_bounds__(kSelThreads) select_topk_segs_kernel(const float* __restrict__ dets, long long n_total, int row_floats,
                                                                       int score_col, SelSegs segs, int k_slots, float min_score, int inclusive,
                                                                       float* __restrict__ out, int out_rows, int* __restrict__ overflow,
                                                                       const float* __restrict__ scores) {
    __shared__ SelShared sh;
    const int b = blockIdx.x, seg = blockIdx.y;  // level-major dispatch: the long (fine-level) segments of every image start first
    const int lo = segs.off[seg], hi = segs.off[seg + 1];
    RowSource src;
    src.rows = dets + (static_cast<long long>(b) * n_total + lo) * row_floats;
    src.scores = scores ? scores + static_cast<long long>(b) * n_total + lo : nullptr;
    src.out = out + (static_cast<long long>(b) * out_rows + static_cast<long long>(seg) * k_slots) * row_floats;
    src.out_src = nullptr;
    src.row_floats = row_floats, src.score_col = score_col, src.n = hi - lo, src.first_index = lo;
    select_core(sh, src, k_slots, min_score, inclusive);
    if (overflow && threadIdx.x == 0 && sh.n_pass > static_cast<unsigned>(k_slots)) atomicOr(overflow + b, 1);
}

__global__ void __launch_bounds__(kSelThreads) select_topk_kernel(const float* __restrict__ dets, long long n_total, int row_floats, int score_col,
                                                                  const int* __restrict__ seg_off /*[n_seg+1] device*/, int k_slots,
                                                                  float min_score, int inclusive, float* __restrict__ out,
                                                                  int* __restrict__ out_src, int out_rows) {
    __shared__ SelShared sh;
    const int b = blockIdx.x, seg = blockIdx.y;
    const int lo = seg_off[seg], hi = seg_off[seg + 1];
    RowSource src;
    src.rows = dets + (static_cast<long long>(b) * n_total + lo) * row_floats;
    src.scores = nullptr;
    src.out = out + (static_cast<long long>(b) * out_rows + static_cast<long long>(seg) * k_slots) * row_floats;
    src.out_src = out_src ? out_src + static_cast<long long>(b) * out_rows + static_cast<long long>(seg) * k_slots : nullptr;
    src.row_floats = row_floats, src.score_col = score_col, src.n = hi - lo, src.first_index = lo;
    select_core(sh, src, k_slots, min_score, inclusive);
}

// Fused FCOS front end (FCOS/infer_fcos.py:35-57 + the pre-NMS selection): one CTA per (level, image) reads the
// head rows [Hl*Wl, C+5], scores every (location, class) pair on the fly -- sigmoid(class) [* sigmoid(centerness)],
// the same float32 expression as fcos_decode_kernel -- and emits the selected pairs as candidate rows
// (y1, x1, y2, x2, score, class) in (location, class) order.  Neither the [B, N, C] score tensor nor the decoded
// boxes of unselected locations are ever written.
struct FcosSource {
    const float* head;  // [rows, C+5] of this (image, level)
    float* out;         // [k, 6]
    int n, num_classes, ch, wl, center;
    float stride;
    FastDiv div_c;
    __device__ __forceinline__ float score(int i) const {
        const int loc = static_cast<int>(fdiv_u32(static_cast<uint32_t>(i), div_c));
        const int c = i - loc * num_classes;
        const float* q = head + static_cast<long long>(loc) * ch;
        return fcos_pair_score(center, __ldg(q + 5 + c), center ? __ldg(q + 4) : 0.0f);
    }
    __device__ __forceinline__ float logit(int i) const {
        const int loc = static_cast<int>(fdiv_u32(static_cast<uint32_t>(i), div_c));
        return __ldg(head + static_cast<long long>(loc) * ch + 5 + (i - loc * num_classes));
    }
    __device__ __forceinline__ void emit(int i, float s, int rank) const {
        const int loc = static_cast<int>(fdiv_u32(static_cast<uint32_t>(i), div_c));
        const int c = i - loc * num_classes;
        const float* q = head + static_cast<long long>(loc) * ch;
        CornerParams cp;
        cp.mode = 0, cp.stride = stride;
        const float4 o = corners_of(cp, loc / wl, loc % wl, 0, q[0], q[1], q[2], q[3]);
        float* dst = out + static_cast<long long>(rank) * 6;
        dst[0] = o.x, dst[1] = o.y, dst[2] = o.z, dst[3] = o.w, dst[4] = s, dst[5] = static_cast<float>(c);
    }
    __device__ __forceinline__ void pad(int rank) const {
        float* dst = out + static_cast<long long>(rank) * 6;
        dst[0] = dst[1] = dst[2] = dst[3] = 0.f, dst[4] = -INFINITY, dst[5] = 0.f;
    }
};

// Without centerness the score is a non-decreasing function of the class logit, so the three passes of the
// selector can run on raw logits (no transcendental per element).  Exactness needs care in two places where distinct
// logits collide after the float32 sigmoid: (1) the score threshold -- logits within 1e-3 of logit(thr) are decided
// with the real sigmoid; (2) ties at the k-th score must be cut by index -- the entries ranked exactly in shared
// memory are those of the boundary bin AND of enough neighbouring bins that no score tie can reach outside them
// (sigmoid_collision_floor bounds how far below a logit a colliding logit can lie).
// kVec: the head rows of one (image, level) are walked as float4 items in memory order (4 logits per load; the five
// regression / centerness channels of each row come along and are skipped); otherwise one (location, class) pair per
// load.  Memory order IS (location, class) order, so the ordered compaction emits the same sequence either way.
constexpr float kLogitBinScale = 4096.0f / 24.0f;  // bins of 0.0059 logit units from the threshold upwards

__device__ __forceinline__ bool logit_space_applies(const FcosSource& src, float min_score, int k) {
    return !src.center && min_score > 1.0e-6f && min_score < 0.999f && k > 0;
}

// The (image, level) segment seen as load items, with the logit-space threshold test and bin map.
__device__ __noinline__ bool logit_passes_exact(float x, float min_score, int inclusive) {
    const float sc = sigmoid_acc(x);
    return inclusive ? (sc >= min_score) : (sc > min_score);
}
template <bool kVec>
struct LogitItems {
    static constexpr int W = kVec ? 4 : 1;  // logits per load item
    struct Item {
        float x[W];  // the item's values as loaded (-inf for an item outside the range)
        int r, c5;   // row and (channel - 5) of slot 0 (vector items); unused for scalar items
    };
    const FcosSource* src;
    int n_items, inclusive;
    FastDiv div_ch;
    float min_score, thr_lo, thr_hi, thr_pass;  // x >= thr_pass: passes for sure; x < thr_lo: fails for sure

    __device__ __forceinline__ void init(const FcosSource& s, float min_score_, int inclusive_) {
        src = &s, min_score = min_score_, inclusive = inclusive_;
        const int rows = s.n / s.num_classes;
        n_items = kVec ? (rows * s.ch) >> 2 : s.n;
        div_ch = make_fastdiv(static_cast<uint32_t>(s.ch));
        const float x_thr = logf(min_score_ / (1.0f - min_score_));
        thr_lo = x_thr - 1.0e-3f, thr_hi = x_thr + 1.0e-3f;
        thr_pass = nextafterf(thr_hi, INFINITY);
    }
    // nothing but the load: the loads of a batch of items issue back to back
    __device__ __forceinline__ void fetch(int item, bool valid, Item& it) const {
        if (!kVec) {
            it.x[0] = valid ? src->logit(item) : -INFINITY;
            return;
        }
        float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (valid) v = __ldg(reinterpret_cast<const float4*>(src->head) + item);
        it.x[0] = v.x, it.x[kVec ? 1 : 0] = v.y, it.x[kVec ? 2 : 0] = v.z, it.x[kVec ? 3 : 0] = v.w;
    }
    __device__ __forceinline__ void locate(int item, Item& it) const {
        if (!kVec) return;
        const int p0 = item << 2;
        it.r = static_cast<int>(fdiv_u32(static_cast<uint32_t>(p0), div_ch));
        it.c5 = p0 - it.r * src->ch - 5;
    }
    // slot e of the item is a class logit: not one of the five regression / centerness channels of this row or, where
    // the item runs over the end of its row, of the next one
    __device__ __forceinline__ bool is_class(const Item& it, int e) const {
        return !kVec || static_cast<unsigned>(it.c5 + e) < static_cast<unsigned>(src->num_classes);
    }
    // the threshold test of a class logit (the float32 sigmoid decides within 1e-3 of the threshold logit)
    __device__ __forceinline__ bool passes(float x) const { return x >= thr_lo && (x >= thr_pass || logit_passes_exact(x, min_score, inclusive)); }
    // (location, class) pair index of a class-logit slot; only needed for the few entries near the cut
    __device__ __forceinline__ int pair_index(int item, const Item& it, int e) const {
        return kVec ? it.r * src->num_classes + it.c5 + e : item;
    }
    // the load item that holds (location, class) pair p
    __device__ __forceinline__ int item_of_pair(int p) const {
        if (!kVec) return p;
        const int r = static_cast<int>(fdiv_u32(static_cast<uint32_t>(p), src->div_c));
        return (r * src->ch + 5 + (p - r * src->num_classes)) >> 2;
    }
    __device__ __forceinline__ int bin_x(float x) const {  // monotone; x >= thr_lo for every entry that passes
        return min(max(__float2int_rz((x - thr_lo) * kLogitBinScale), 0), 4095);
    }
    // Logit bounds for the passes that follow the histogram: a class logit >= take_from passes the threshold and lies in
    // a bin above hi_bin; one < look_from fails the threshold or lies in a bin below lo_bin (0.01 bins of slack cover
    // the rounding of bin_x); the few in between are decided exactly.
    __device__ __forceinline__ float take_from(bool take_all, int hi_bin) const {
        return take_all ? thr_pass : fmaxf(thr_pass, thr_lo + (static_cast<float>(hi_bin) + 1.01f) / kLogitBinScale);
    }
    __device__ __forceinline__ float look_from(bool take_all, int lo_bin) const {
        return take_all ? thr_lo : fmaxf(thr_lo, thr_lo + (static_cast<float>(lo_bin) - 0.01f) / kLogitBinScale);
    }
};

// One batch of N load items: fetch them all, then every class-logit slot with x >= x_hot gets hot(u, e, x) -- the
// cheap, branch-free common case -- and the slots with x_lo <= x < x_hot are remembered in a bit mask and handed to
// slow(u, e, x) afterwards (rare: one branch per batch).  The passes are instruction-issue-bound before they are
// memory-bound, so the per-slot code is a handful of instructions.
template <bool kVec, int N, class Idx, class Hot, class Slow>
__device__ __forceinline__ void logit_scan(const LogitItems<kVec>& it, typename LogitItems<kVec>::Item (&q)[N], int i_hi, float x_hot,
                                           float x_lo, const Idx& idx, const Hot& hot, const Slow& slow) {
    constexpr int W = LogitItems<kVec>::W;
#pragma unroll
    for (int u = 0; u < N; ++u) it.fetch(idx(u), idx(u) < i_hi, q[u]);
    unsigned maybe = 0;
#pragma unroll
    for (int u = 0; u < N; ++u) {
        it.locate(idx(u), q[u]);
#pragma unroll
        for (int e = 0; e < W; ++e) {
            const float x = q[u].x[e];
            const bool cls = it.is_class(q[u], e);
            const bool h = cls && x >= x_hot;
            if (h) hot(u, e, x);
            if (cls && x >= x_lo && !h) maybe |= 1u << (u * W + e);
        }
    }
    if (maybe) {
#pragma unroll
        for (int u = 0; u < N; ++u)
#pragma unroll
            for (int e = 0; e < W; ++e)
                if ((maybe >> (u * W + e)) & 1u) slow(u, e, q[u].x[e]);
    }
}

// One thread, after find_boundary put the k-th best into bin Bk with `above` entries in the bins over it: the bins
// [lo, hi] whose entries can tie with the k-th score after the sigmoid, and how many of them are to be taken.
// False when the window cannot be bounded or holds more entries than the shared-memory list.
__device__ inline bool logit_tie_window(const unsigned* hist, int Bk, float thr_lo, unsigned k, unsigned above, int& lo, int& hi, unsigned& need) {
    const float lo_edge = thr_lo + static_cast<float>(Bk) / kLogitBinScale, hi_edge = thr_lo + static_cast<float>(Bk + 1) / kLogitBinScale;
    const int below = static_cast<int>(ceilf((lo_edge - sigmoid_collision_floor(lo_edge)) * kLogitBinScale)) + 1;
    int beyond = -1;
    for (int t = 1; t <= 8 && beyond < 0; ++t)
        if (sigmoid_collision_floor(hi_edge + static_cast<float>(t) / kLogitBinScale) > hi_edge) beyond = t + 1;
    lo = max(Bk - below, 0), hi = Bk + beyond;
    if (!(beyond > 0 && hi < 4095)) return false;
    unsigned inside = 0, between = 0;
    for (int q = lo; q <= hi; ++q) inside += hist[q];
    for (int q = Bk + 1; q <= hi; ++q) between += hist[q];
    need = k - (above - between);
    return inside <= kSelListCap;
}

// Rank the collected window entries exactly by (score desc, index asc); the entry of rank need-1 is the cut.
// `on_taken(index)` is called for every entry at or before the cut.
template <int T, class OnTaken>
__device__ __forceinline__ void rank_window(SelShared& sh, unsigned need, int tid, const OnTaken& on_taken) {
    const int n_list = static_cast<int>(sh.n_list);
    for (int e = tid; e < n_list; e += T) {
        const unsigned ke = sh.list_key[e];
        const int ie = sh.list_idx[e];
        unsigned rank = 0;
        for (int f = 0; f < n_list; ++f) {
            const unsigned kf = sh.list_key[f];
            rank += (kf > ke || (kf == ke && sh.list_idx[f] < ie)) ? 1u : 0u;
        }
        if (rank == need - 1) sh.t_key = ke, sh.t_idx = ie;
        if (rank < need) on_taken(ie);
    }
}

// ---- the three passes over the load items [i_lo, i_hi) of a segment ----------------------------------------------------
// 1. histogram of the passing logits; returns this thread's number of passing entries
template <bool kVec, int T>
__device__ __forceinline__ unsigned logit_histogram(unsigned* hist, const LogitItems<kVec>& it, int i_lo, int i_hi) {
    constexpr int U = kVec ? 4 : 8;  // independent loads in flight per thread in the unordered passes
    const int tid = threadIdx.x;
    unsigned n_pass = 0;
    for (int base = i_lo; base < i_hi; base += T * U) {
        typename LogitItems<kVec>::Item q[U];
        logit_scan<kVec, U>(
            it, q, i_hi, it.thr_pass, it.thr_lo, [&](int u) { return base + u * T + tid; },
            [&](int, int, float x) { atomicAdd(&hist[it.bin_x(x)], 1u), ++n_pass; },
            [&](int, int, float x) {
                if (it.passes(x)) atomicAdd(&hist[it.bin_x(x)], 1u), ++n_pass;
            });
    }
    return n_pass;
}
// 2. the entries of the tie window [lo_bin, hi_bin] are appended (exact score key, pair index) to a list, which may
//    live in another CTA of the cluster; returns this thread's number of entries above the window
template <bool kVec, int T>
__device__ __forceinline__ unsigned logit_collect(const LogitItems<kVec>& it, int i_lo, int i_hi, int lo_bin, int hi_bin, unsigned* n_list,
                                                  unsigned* list_key, int* list_idx) {
    constexpr int U = kVec ? 4 : 8;
    const int tid = threadIdx.x;
    const float x_take = it.take_from(false, hi_bin), x_look = it.look_from(false, lo_bin);
    unsigned n_above = 0;
    for (int base = i_lo; base < i_hi; base += T * U) {
        typename LogitItems<kVec>::Item q[U];
        logit_scan<kVec, U>(
            it, q, i_hi, x_take, x_look, [&](int u) { return base + u * T + tid; }, [&](int, int, float) { ++n_above; },
            [&](int u, int e, float x) {
                if (!it.passes(x)) return;
                const int bn = it.bin_x(x);
                if (bn > hi_bin) ++n_above;
                if (bn < lo_bin || bn > hi_bin) return;
                const unsigned slot = atomicAdd(n_list, 1u);
                list_key[slot] = score_key(sigmoid_acc(x));
                list_idx[slot] = it.pair_index(base + u * T + tid, q[u], e);
            });
    }
    return n_above;
}
// 3. ordered compaction: each thread owns kItems consecutive logits per chunk, the taken ones are emitted to slots
//    first_rank, first_rank + 1, ... in item order.  Returns the number taken.
template <bool kVec, int T, int kItems>
__device__ __forceinline__ unsigned logit_compact(SelShared& sh, const LogitItems<kVec>& it, int i_lo, int i_hi, bool take_all, int lo_bin,
                                                  int hi_bin, unsigned T_key, int T_idx, unsigned first_rank, int k) {
    constexpr int W = LogitItems<kVec>::W;
    constexpr int IPT = kItems / W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float x_take = it.take_from(take_all, hi_bin), x_look = it.look_from(take_all, lo_bin);
    unsigned carry = first_rank;
    int chunk = 0;
    for (int base = i_lo; base < i_hi; base += T * IPT, ++chunk) {
        typename LogitItems<kVec>::Item q[IPT];
        unsigned flags = 0;
        logit_scan<kVec, IPT>(
            it, q, i_hi, x_take, x_look, [&](int u) { return base + tid * IPT + u; }, [&](int u, int e, float) { flags |= 1u << (u * W + e); },
            [&](int u, int e, float x) {
                if (!it.passes(x)) return;
                bool take = take_all;
                if (!take) {
                    const int bn = it.bin_x(x);
                    if (bn > hi_bin) take = true;
                    else if (bn >= lo_bin) {
                        const unsigned key = score_key(sigmoid_acc(x));
                        take = key > T_key || (key == T_key && it.pair_index(base + tid * IPT + u, q[u], e) <= T_idx);
                    }
                }
                if (take) flags |= 1u << (u * W + e);
            });
        const unsigned mine = __popc(flags);
        unsigned incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        unsigned* wt = sh.wtot[chunk & 1];  // parity double buffer: one barrier per chunk
        if (lane == 31) wt[warp] = incl;
        __syncthreads();
        unsigned before = carry, tot = 0;
        for (int w = 0; w < T / 32; ++w) {
            const unsigned v = wt[w];
            if (w < warp) before += v;
            tot += v;
        }
        unsigned rank = before + incl - mine;
        if (flags) {  // emit the taken slots in order
#pragma unroll
            for (int u = 0; u < IPT; ++u)
#pragma unroll
                for (int e = 0; e < W; ++e) {
                    if (!((flags >> (u * W + e)) & 1u)) continue;
                    if (rank < static_cast<unsigned>(k))
                        it.src->emit(it.pair_index(base + tid * IPT + u, q[u], e), sigmoid_acc(q[u].x[e]), static_cast<int>(rank));
                    ++rank;
                }
        }
        carry += tot;
    }
    return carry - first_rank;
}

// One CTA does the whole segment.  Returns false (nothing written) when the shortcut does not apply; the caller then
// runs the generic exact selector.
template <bool kVec>
__device__ __forceinline__ bool fcos_select_logit_space(SelShared& sh, const FcosSource& src, int k_slots, float min_score, int inclusive) {
    const int tid = threadIdx.x;
    const int k = min(k_slots, max(src.n, 0));
    if (!logit_space_applies(src, min_score, k)) return false;
    LogitItems<kVec> it;
    it.init(src, min_score, inclusive);
    for (int i = tid; i < 4096; i += kSelThreads) sh.hist[i] = 0;
    if (tid == 0) sh.n_list = 0, sh.bin = 0xFFFFFFFFu, sh.above = 0, sh.n_pass = 0, sh.t_key = 0u, sh.t_idx = 0x7fffffff;
    __syncthreads();
    logit_histogram<kVec, kSelThreads>(sh.hist, it, 0, it.n_items);
    __syncthreads();
    find_boundary(sh, 4096, static_cast<unsigned>(k), tid);
    const bool take_all = sh.bin == 0xFFFFFFFFu;  // fewer than k pass
    int lo_bin = 0, hi_bin = 4095;
    if (!take_all) {
        if (tid == 0) {
            int lo, hi;
            unsigned need = 0;
            const bool ok = logit_tie_window(sh.hist, static_cast<int>(sh.bin), it.thr_lo, static_cast<unsigned>(k), sh.above, lo, hi, need);
            sh.wtot[1][0] = ok ? 1u : 0u, sh.wtot[1][1] = static_cast<unsigned>(lo), sh.wtot[1][2] = static_cast<unsigned>(hi);
            sh.need = need;
        }
        __syncthreads();
        if (!sh.wtot[1][0]) {
            __syncthreads();
            return false;
        }
        lo_bin = static_cast<int>(sh.wtot[1][1]), hi_bin = static_cast<int>(sh.wtot[1][2]);
        const unsigned need = sh.need;
        logit_collect<kVec, kSelThreads>(it, 0, it.n_items, lo_bin, hi_bin, &sh.n_list, sh.list_key, sh.list_idx);
        __syncthreads();
        rank_window<kSelThreads>(sh, need, tid, [](int) {});
        __syncthreads();
    }
    const unsigned taken = logit_compact<kVec, kSelThreads, kSelItems>(sh, it, 0, it.n_items, take_all, lo_bin, hi_bin, sh.t_key, sh.t_idx, 0u, k);
    const unsigned filled = min(static_cast<unsigned>(k), taken);
    for (int r = filled + tid; r < k_slots; r += kSelThreads) src.pad(r);
    return true;
}

struct FcosSelectArgs {
    const float* head[DH_MAX_LEVELS];
    int hl[DH_MAX_LEVELS], wl[DH_MAX_LEVELS];
    float stride[DH_MAX_LEVELS];
    int num_classes, center, k_slots, n_levels, inclusive, allow_logit_space;
    float min_score;
};
__device__ __forceinline__ FcosSource fcos_source(const FcosSelectArgs& a, int b, int l, float* cand) {
    FcosSource src;
    const int rows = a.hl[l] * a.wl[l];
    src.ch = a.num_classes + 5, src.num_classes = a.num_classes, src.wl = a.wl[l], src.center = a.center, src.stride = a.stride[l];
    src.head = a.head[l] + static_cast<long long>(b) * rows * src.ch;
    src.out = cand + (static_cast<long long>(b) * a.n_levels + l) * a.k_slots * 6;
    src.n = rows * a.num_classes;
    src.div_c = make_fastdiv(static_cast<uint32_t>(a.num_classes));
    return src;
}
__device__ __forceinline__ bool fcos_source_vec(const FcosSource& src) {  // uniform over everything that works on the segment
    return (reinterpret_cast<uintptr_t>(src.head) & 15u) == 0 && (((src.n / src.num_classes) * src.ch) & 3) == 0;
}
__global__ void __launch_bounds__(kSelThreads) fcos_select_kernel(FcosSelectArgs a, float* __restrict__ cand /*[B, L*k, 6]*/) {
    __shared__ SelShared sh;
    const int b = blockIdx.x, l = blockIdx.y;  // level-major dispatch: the 64 x 512 K-logit level-0 CTAs start first, the rest fill in
    const FcosSource src = fcos_source(a, b, l, cand);
    if (a.allow_logit_space) {
        if (fcos_source_vec(src) ? fcos_select_logit_space<true>(sh, src, a.k_slots, a.min_score, a.inclusive)
                                 : fcos_select_logit_space<false>(sh, src, a.k_slots, a.min_score, a.inclusive))
            return;
    }
    select_core(sh, src, a.k_slots, a.min_score, a.inclusive);
}

// ---- the same selection with a thread-block cluster per segment --------------------------------------------------------
// One CTA per (image, level) leaves most of the GPU idle: 64 images give 64 long level-0 segments (2.2 MB each, read three
// times) for 148 SMs.  Here a *group* of up to 8 CTAs of one cluster shares a segment: every CTA histograms, collects
// and compacts its contiguous slice of the load items, and the group meets in the shared memory of its leader CTA
// (distributed shared memory): the slice histograms are summed into the leader's, the leader finds the boundary bin and
// the tie window, every CTA appends its window entries to the leader's list, the leader ranks them and counts how many
// each slice keeps, and each CTA then knows the output slot its slice starts at.  A cluster holds one long segment or
// several short ones of the same image (host-side plan); the six cluster barriers are executed by every CTA.
constexpr int kSelCluster = 8;
#ifndef DH_CLUSTER_SEL_THREADS
#define DH_CLUSTER_SEL_THREADS 512
#endif
constexpr int kClThreads = DH_CLUSTER_SEL_THREADS;  // several CTAs per SM: one cluster's barriers and leader-only steps overlap another's streaming
constexpr int kClItems = 16;                         // consecutive logits per thread and chunk in the compaction pass (4 loads in flight)
struct FcosClusterPlan {
    int n_types;                                     // clusters per image
    signed char level[DH_MAX_LEVELS][kSelCluster];   // level worked on by rank r of cluster type t; -1: idle
    unsigned char lead[DH_MAX_LEVELS][kSelCluster];  // rank of the group's leader CTA
    unsigned char members[DH_MAX_LEVELS][kSelCluster];
};
struct ClusterCtl {
    unsigned mode, lo_bin, hi_bin, need, local_cnt;
    unsigned cnt_above[kSelCluster];   // per slice: entries above the tie window (or all passing entries when everything is taken)
    unsigned cnt_window[kSelCluster];  // per slice: window entries at or before the cut
};
enum : unsigned { kModeTakeAll = 1u, kModeRanked = 2u, kModeFallback = 3u };

template <bool kVec>
__device__ __forceinline__ void fcos_select_cluster_group(SelShared& sh, ClusterCtl& ctl, const FcosSource& src, int k_slots, float min_score,
                                                          int inclusive, bool active, int g, int G, int lead) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31;
    SelShared* lsh = cluster.map_shared_rank(&sh, lead);
    ClusterCtl* lctl = cluster.map_shared_rank(&ctl, lead);
    const bool leader = active && g == 0;
    const int k = active ? min(k_slots, max(src.n, 0)) : 0;
    LogitItems<kVec> it;
    int per = 1, i_lo = 0, i_hi = 0;
    if (active) {
        it.init(src, min_score, inclusive);
        per = max((it.n_items + G - 1) / G, 1);
        i_lo = min(g * per, it.n_items), i_hi = min(i_lo + per, it.n_items);
    }
    for (int i = tid; i < 4096; i += kClThreads) sh.hist[i] = 0;
    if (tid == 0) sh.n_list = 0, sh.bin = 0xFFFFFFFFu, sh.above = 0, sh.n_pass = 0, sh.t_key = 0u, sh.t_idx = 0x7fffffff;
    if (tid == 0) ctl.mode = 0u, ctl.local_cnt = 0u;
    if (tid < kSelCluster) ctl.cnt_above[tid] = 0u, ctl.cnt_window[tid] = 0u;
    __syncthreads();
    // ---- 1. slice histogram
    unsigned my_pass = logit_histogram<kVec, kClThreads>(sh.hist, it, i_lo, i_hi);
    my_pass = static_cast<unsigned>(warp_sum_i(static_cast<int>(my_pass)));
    if (lane == 0 && my_pass) atomicAdd(&sh.n_pass, my_pass);
    cluster.sync();  // A: every slice histogram is complete
    if (active && G > 1) {
        // CTA g sums its share of the bins over the group's histograms and leaves the totals in the leader's
        const int share = (4096 + G - 1) / G;
        const unsigned* peer[kSelCluster];
#pragma unroll
        for (int r = 0; r < kSelCluster; ++r) peer[r] = cluster.map_shared_rank(sh.hist, lead + min(r, G - 1));
        for (int i = g * share + tid; i < min((g + 1) * share, 4096); i += kClThreads) {
            unsigned sum = 0;
#pragma unroll
            for (int r = 0; r < kSelCluster; ++r)
                if (r < G) sum += peer[r][i];
            lsh->hist[i] = sum;
        }
    }
    cluster.sync();  // B: the leader holds the segment's histogram
    if (leader) {
        find_boundary<kClThreads>(sh, 4096, static_cast<unsigned>(k), tid);
        if (tid == 0) {
            if (k <= 0 || sh.bin == 0xFFFFFFFFu) {
                ctl.mode = kModeTakeAll;  // fewer than k pass
            } else {
                int lo, hi;
                unsigned need = 0;
                const bool ok = logit_tie_window(sh.hist, static_cast<int>(sh.bin), it.thr_lo, static_cast<unsigned>(k), sh.above, lo, hi, need);
                ctl.mode = ok ? kModeRanked : kModeFallback;
                ctl.lo_bin = static_cast<unsigned>(lo), ctl.hi_bin = static_cast<unsigned>(hi), ctl.need = need;
            }
        }
    }
    cluster.sync();  // C: mode and tie window are published
    const unsigned mode = active ? lctl->mode : 0u;
    const int lo_bin = mode == kModeRanked ? static_cast<int>(lctl->lo_bin) : 0;
    const int hi_bin = mode == kModeRanked ? static_cast<int>(lctl->hi_bin) : 4095;
    const unsigned need = mode == kModeRanked ? lctl->need : 0u;
    if (mode == kModeRanked) {
        // ---- 2. window entries go to the leader's list with their exact scores; count what lies above the window
        unsigned n_above = logit_collect<kVec, kClThreads>(it, i_lo, i_hi, lo_bin, hi_bin, &lsh->n_list, lsh->list_key, lsh->list_idx);
        n_above = static_cast<unsigned>(warp_sum_i(static_cast<int>(n_above)));
        if (lane == 0 && n_above) atomicAdd(&ctl.local_cnt, n_above);
        __syncthreads();
        if (tid == 0) lctl->cnt_above[g] = ctl.local_cnt;
    } else if (mode == kModeTakeAll) {
        if (tid == 0) lctl->cnt_above[g] = sh.n_pass;
    }
    cluster.sync();  // D: the leader's list and the per-slice counts are complete
    if (leader && mode == kModeRanked) {
        rank_window<kClThreads>(sh, need, tid, [&](int pair) { atomicAdd(&ctl.cnt_window[it.item_of_pair(pair) / per], 1u); });
    }
    cluster.sync();  // E: the cut and the per-slice window counts are published
    unsigned T_key = 0u, first_rank = 0u, total = 0u;
    int T_idx = 0x7fffffff;
    if (mode == kModeRanked || mode == kModeTakeAll) {
        T_key = lsh->t_key, T_idx = lsh->t_idx;
        for (int r = 0; r < G; ++r) {
            const unsigned c = lctl->cnt_above[r] + lctl->cnt_window[r];
            if (r < g) first_rank += c;
            total += c;
        }
    }
    cluster.sync();  // F: nobody reads another CTA's shared memory from here on
    if (mode == kModeRanked || mode == kModeTakeAll) {
        // ---- 3. ordered compaction of the slice into its slots
        logit_compact<kVec, kClThreads, kClItems>(sh, it, i_lo, i_hi, mode == kModeTakeAll, lo_bin, hi_bin, T_key, T_idx, first_rank, k);
        if (leader) {
            const unsigned filled = min(static_cast<unsigned>(k), total);
            for (int r = filled + tid; r < k_slots; r += kClThreads) src.pad(r);
        }
    } else if (leader) {
        // (rare) a tie window too large for the list: the leader redoes the whole segment with the generic exact selector
        __syncthreads();
        select_core<FcosSource, kClThreads>(sh, src, k_slots, min_score, inclusive);
    }
}

__global__ void __launch_bounds__(kClThreads, 1024 / kClThreads) fcos_select_cluster_kernel(const __grid_constant__ FcosSelectArgs a,
                                                                          const __grid_constant__ FcosClusterPlan plan, int batch,
                                                                          float* __restrict__ cand /*[B, L*k, 6]*/) {
    __shared__ SelShared sh;
    __shared__ ClusterCtl ctl;
    const int rank = static_cast<int>(cooperative_groups::this_cluster().block_rank());
    const int cid = blockIdx.x / kSelCluster;  // type-major: the clusters with the long segments start first
    const int type = cid / batch, b = cid - type * batch;
    const int l = plan.level[type][rank];
    const bool active = l >= 0;
    const int lead = plan.lead[type][rank], G = plan.members[type][rank];
    FcosSource src = fcos_source(a, b, active ? l : 0, cand);
    if (active && fcos_source_vec(src))
        fcos_select_cluster_group<true>(sh, ctl, src, a.k_slots, a.min_score, a.inclusive, active, rank - lead, G, lead);
    else
        fcos_select_cluster_group<false>(sh, ctl, src, a.k_slots, a.min_score, a.inclusive, active, rank - lead, G, lead);
}

static int grid_for(long long threads_needed, int block, int sm_count) {
    long long g = (threads_needed + block - 1) / block;
    const long long cap = static_cast<long long>(sm_count) * 16;
    if (g > cap) g = cap;
    return static_cast<int>(g < 1 ? 1 : g);
}

// ---- the same selection for many segments: estimate, one streaming pass, finish ----------------------------------------
// With dozens of long segments the three passes above are issue- and latency-bound.  Here the head is read ONCE by the
// whole GPU: (1) a sample of every 16th load item of a segment gives a logit bound x_est above which about 2.5 k logits
// are expected (the lower edge of a histogram bin; the threshold itself when the sample holds few passing entries or the
// segment is short) -- an estimate only, exactness never depends on it; (2) one streaming kernel over all (image,
// level, chunk) triples appends the class logits >= x_est with their pair index to the segment's candidate list (staged
// in shared memory, one global atomic per chunk); (3) one CTA per segment selects the exact top k among its candidates
// with the same machinery as above -- histogram, boundary bin, tie window ranked exactly -- marks the winners in a
// shared-memory bitmap over the pair indices and emits them in index order by a popcount scan.  A segment whose
// candidates do not provably contain its top k (list overflow, fewer than k candidates above an estimate that is not the
// threshold, a tie window reaching down to the estimate's bin or too large for the list) is redone on the spot by the
// same CTA with the one-CTA selector above.
#ifndef DH_PRE_CHUNK_ITEMS
#define DH_PRE_CHUNK_ITEMS 1024
#endif
#ifndef DH_PRE_THREADS
#define DH_PRE_THREADS 128
#endif
constexpr int kPreChunkItems = DH_PRE_CHUNK_ITEMS;  // load items per CTA of the streaming pass
constexpr int kPreThreads = DH_PRE_THREADS;
constexpr int kPreStage = 1024;         // candidates a chunk stages in shared memory (more go to the list one by one)
constexpr int kPreSampleStride = 16;
constexpr int kPreBitmapWords = 16384;  // 512 K (location, class) pairs per segment

struct PreselWork {
    float* x_est;    // [B*L] logit bound of the candidates
    int* est_bin;    // [B*L] its histogram bin; -1: the bound is the threshold itself (every passing entry is a candidate)
    unsigned* count; // [B*L] candidates appended
    float* cand_x;   // [B*L, cap]
    int* cand_pair;  // [B*L, cap]
    int cap, bitmap_words;  // list capacity per segment; words of the finish kernel's shared-memory bitmap
};

// (1) grid (B, L)
__global__ void __launch_bounds__(kSelThreads) fcos_presel_estimate_kernel(const __grid_constant__ FcosSelectArgs a, PreselWork w) {
    __shared__ SelShared sh;
    const int b = blockIdx.x, l = blockIdx.y, seg = b * a.n_levels + l, tid = threadIdx.x;
    const FcosSource src = fcos_source(a, b, l, nullptr);
    if (tid == 0) w.count[seg] = 0u;
    const int k = min(a.k_slots, max(src.n, 0));
    LogitItems<true> it;
    it.init(src, a.min_score, a.inclusive);
    if (src.n <= w.cap || !fcos_source_vec(src)) {  // short (or unaligned, hence short) segment: everything that passes
        if (tid == 0) w.est_bin[seg] = -1, w.x_est[seg] = it.thr_lo;
        return;
    }
    for (int i = tid; i < 4096; i += kSelThreads) sh.hist[i] = 0;
    if (tid == 0) sh.bin = 0xFFFFFFFFu, sh.above = 0;
    __syncthreads();
    const int n_sample = it.n_items / kPreSampleStride;
    for (int base = 0; base < n_sample; base += kSelThreads * 4) {
        LogitItems<true>::Item q[4];
        logit_scan<true, 4>(
            it, q, n_sample * kPreSampleStride, it.thr_pass, INFINITY, [&](int u) { return (base + u * kSelThreads + tid) * kPreSampleStride; },
            [&](int, int, float x) { atomicAdd(&sh.hist[it.bin_x(x)], 1u); }, [](int, int, float) {});
    }
    __syncthreads();
    // aim at 2.5 k candidates: the r-th largest of the sample with r = 2.5 k / stride
    const unsigned r = static_cast<unsigned>(max((5 * k) / (2 * kPreSampleStride), 8));
    find_boundary<kSelThreads>(sh, 4096, r, tid);
    if (tid == 0) {
        const bool all = sh.bin == 0xFFFFFFFFu || sh.bin == 0u;
        w.est_bin[seg] = all ? -1 : static_cast<int>(sh.bin);
        w.x_est[seg] = all ? it.thr_lo : it.thr_lo + static_cast<float>(sh.bin) / kLogitBinScale;
    }
}

// (2) one CTA per (level, image, chunk); the chunk table is level-major so that the long segments start first
struct PreselChunks {
    int first[DH_MAX_LEVELS + 1];  // first chunk id of each level (ids run level > image > chunk)
    int per_seg[DH_MAX_LEVELS];    // chunks per segment
};
template <bool kVec>
__device__ __forceinline__ void presel_collect_chunk(const FcosSource& src, const FcosSelectArgs& a, const PreselWork& w, int seg, int chunk,
                                                     float* st_x, int* st_pair, unsigned* st_n) {
    LogitItems<kVec> it;
    it.init(src, a.min_score, a.inclusive);
    const int tid = threadIdx.x;
    const int i_lo = chunk * kPreChunkItems, i_hi = min(i_lo + kPreChunkItems, it.n_items);
    const float x_est = w.x_est[seg];
    constexpr int U = kVec ? 4 : 8;
    for (int base = i_lo; base < i_hi; base += kPreThreads * U) {
        typename LogitItems<kVec>::Item q[U];
        logit_scan<kVec, U>(
            it, q, i_hi, x_est, INFINITY, [&](int u) { return base + u * kPreThreads + tid; },
            [&](int u, int e, float x) {
                const int pair = it.pair_index(base + u * kPreThreads + tid, q[u], e);
                const unsigned slot = atomicAdd(st_n, 1u);
                if (slot < kPreStage) {
                    st_x[slot] = x, st_pair[slot] = pair;
                } else {  // the stage is full (a dense chunk): append to the segment's list directly
                    const unsigned g = atomicAdd(&w.count[seg], 1u);
                    if (g < static_cast<unsigned>(w.cap)) {
                        w.cand_x[static_cast<long long>(seg) * w.cap + g] = x;
                        w.cand_pair[static_cast<long long>(seg) * w.cap + g] = pair;
                    }
                }
            },
            [](int, int, float) {});
    }
}
__global__ void __launch_bounds__(kPreThreads) fcos_presel_collect_kernel(const __grid_constant__ FcosSelectArgs a,
                                                                        const __grid_constant__ PreselChunks ct, int batch, PreselWork w) {
    __shared__ float st_x[kPreStage];
    __shared__ int st_pair[kPreStage];
    __shared__ unsigned st_n, st_base;
    int l = 0;
    while (l + 1 < a.n_levels && static_cast<int>(blockIdx.x) >= ct.first[l + 1]) ++l;
    const int local = blockIdx.x - ct.first[l];
    const int b = local / ct.per_seg[l], chunk = local - b * ct.per_seg[l];
    const int seg = b * a.n_levels + l, tid = threadIdx.x;
    if (tid == 0) st_n = 0u;
    __syncthreads();
    const FcosSource src = fcos_source(a, b, l, nullptr);
    if (fcos_source_vec(src)) presel_collect_chunk<true>(src, a, w, seg, chunk, st_x, st_pair, &st_n);
    else presel_collect_chunk<false>(src, a, w, seg, chunk, st_x, st_pair, &st_n);
    __syncthreads();
    const unsigned n = min(st_n, static_cast<unsigned>(kPreStage));
    if (n == 0u) return;
    if (tid == 0) st_base = atomicAdd(&w.count[seg], n);
    __syncthreads();
    const unsigned base = st_base;
    for (unsigned i = tid; i < n; i += kPreThreads)
        if (base + i < static_cast<unsigned>(w.cap)) {
            w.cand_x[static_cast<long long>(seg) * w.cap + base + i] = st_x[i];
            w.cand_pair[static_cast<long long>(seg) * w.cap + base + i] = st_pair[i];
        }
}

// (3) grid (B, L); dynamic shared memory: candidates (x, pair) [cap] + bitmap [kPreBitmapWords]
__global__ void __launch_bounds__(kSelThreads) fcos_presel_finish_kernel(const __grid_constant__ FcosSelectArgs a, PreselWork w,
                                                                       float* __restrict__ cand /*[B, L*k, 6]*/) {
    __shared__ SelShared sh;
    __shared__ unsigned s_pass;
    extern __shared__ __align__(16) unsigned char dyn[];
    float* cx = reinterpret_cast<float*>(dyn);
    int* cp = reinterpret_cast<int*>(dyn) + w.cap;
    unsigned* ckey = reinterpret_cast<unsigned*>(dyn) + 2 * w.cap;
    unsigned* bitmap = reinterpret_cast<unsigned*>(dyn) + 3 * w.cap;
    const int b = blockIdx.x, l = blockIdx.y, seg = b * a.n_levels + l, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const FcosSource src = fcos_source(a, b, l, cand);
    const int k = min(a.k_slots, max(src.n, 0));
    const unsigned count = w.count[seg];
    const int est_bin = w.est_bin[seg];
    const int words = (src.n + 31) >> 5;
    const bool done = [&]() -> bool {  // false: the candidates do not provably hold the segment's top k
    if (count > static_cast<unsigned>(w.cap) || words > w.bitmap_words) return false;
    LogitItems<false> it;  // only the threshold test and the bin map are used here
    it.init(src, a.min_score, a.inclusive);
    const int n = static_cast<int>(count);
    for (int i = tid; i < 4096; i += kSelThreads) sh.hist[i] = 0;
    for (int i = tid; i < words; i += kSelThreads) bitmap[i] = 0u;
    if (tid == 0) sh.n_list = 0, sh.bin = 0xFFFFFFFFu, sh.above = 0, sh.t_key = 0u, sh.t_idx = 0x7fffffff, s_pass = 0u;
    __syncthreads();
    // candidates -> shared memory with their exact score keys (0: fails the threshold), histogram of the passing ones
    unsigned my_pass = 0;
    for (int i = tid; i < n; i += kSelThreads) {
        const float x = w.cand_x[static_cast<long long>(seg) * w.cap + i];
        cx[i] = x, cp[i] = w.cand_pair[static_cast<long long>(seg) * w.cap + i];
        const bool ok = it.passes(x);
        ckey[i] = ok ? score_key(sigmoid_acc(x)) : 0u;
        if (ok) atomicAdd(&sh.hist[it.bin_x(x)], 1u), ++my_pass;
    }
    my_pass = static_cast<unsigned>(warp_sum_i(static_cast<int>(my_pass)));
    if (lane == 0 && my_pass) atomicAdd(&s_pass, my_pass);
    __syncthreads();
    const unsigned n_pass = s_pass;
    bool take_all = false;
    int lo_bin = 0, hi_bin = 4095;
    if (n_pass < static_cast<unsigned>(k)) {
        if (est_bin >= 0) return false;  // fewer than k candidates, and they are not everything that passes
        take_all = true;
    } else if (k > 0) {
        find_boundary<kSelThreads>(sh, 4096, static_cast<unsigned>(k), tid);
        if (tid == 0) {
            int lo, hi;
            unsigned need = 0;
            bool ok = logit_tie_window(sh.hist, static_cast<int>(sh.bin), it.thr_lo, static_cast<unsigned>(k), sh.above, lo, hi, need);
            ok = ok && lo > est_bin;  // every entry of the segment in the window's bins is a candidate
            sh.wtot[1][0] = ok ? 1u : 0u, sh.wtot[1][1] = static_cast<unsigned>(lo), sh.wtot[1][2] = static_cast<unsigned>(hi);
            sh.need = need;
        }
        __syncthreads();
        if (!sh.wtot[1][0]) return false;
        lo_bin = static_cast<int>(sh.wtot[1][1]), hi_bin = static_cast<int>(sh.wtot[1][2]);
        for (int i = tid; i < n; i += kSelThreads) {
            if (!ckey[i]) continue;
            const int bn = it.bin_x(cx[i]);
            if (bn < lo_bin || bn > hi_bin) continue;
            const unsigned slot = atomicAdd(&sh.n_list, 1u);
            sh.list_key[slot] = ckey[i], sh.list_idx[slot] = cp[i];
        }
        __syncthreads();
        rank_window<kSelThreads>(sh, sh.need, tid, [](int) {});
        __syncthreads();
    }
    const unsigned T_key = sh.t_key;
    const int T_idx = sh.t_idx;
    // winners -> bitmap over the pair indices
    for (int i = tid; i < n; i += kSelThreads) {
        const unsigned key = ckey[i];
        bool take = key != 0u;
        if (take && !take_all) {
            const int bn = it.bin_x(cx[i]);
            take = bn > hi_bin || (bn >= lo_bin && (key > T_key || (key == T_key && cp[i] <= T_idx)));
        }
        if (take) atomicOr(&bitmap[cp[i] >> 5], 1u << (cp[i] & 31));
        else ckey[i] = 0u;
    }
    __syncthreads();
    // ranks in index order: popcount scan over the bitmap (thread t owns words [t*wpt, (t+1)*wpt))
    const int wpt = (words + kSelThreads - 1) / kSelThreads;
    unsigned mine = 0;
    for (int q = 0; q < wpt; ++q) {
        const int wi = tid * wpt + q;
        if (wi < words) mine += __popc(bitmap[wi]);
    }
    unsigned incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) sh.wtot[0][warp] = incl;
    __syncthreads();
    unsigned before = 0, total = 0;
    for (int q = 0; q < kSelThreads / 32; ++q) {
        const unsigned v = sh.wtot[0][q];
        if (q < warp) before += v;
        total += v;
    }
    unsigned* prefix = sh.list_key;  // [kSelThreads] exclusive prefix per thread (the list is no longer needed)
    __syncthreads();
    prefix[tid] = before + incl - mine;
    __syncthreads();
    for (int i = tid; i < n; i += kSelThreads) {
        if (!ckey[i]) continue;
        const int pair = cp[i], wi = pair >> 5, owner = wi / wpt;
        unsigned rank = prefix[owner] + __popc(bitmap[wi] & ((1u << (pair & 31)) - 1u));
        for (int q = owner * wpt; q < wi; ++q) rank += __popc(bitmap[q]);
        if (rank < static_cast<unsigned>(k)) src.emit(pair, sigmoid_acc(cx[i]), static_cast<int>(rank));
    }
    const unsigned filled = min(static_cast<unsigned>(k), total);
    for (int r = filled + tid; r < a.k_slots; r += kSelThreads) src.pad(r);
    return true;
    }();
    if (done) return;
    // redo the segment with the one-CTA selector
    __syncthreads();
    if (fcos_source_vec(src) ? fcos_select_logit_space<true>(sh, src, a.k_slots, a.min_score, a.inclusive)
                             : fcos_select_logit_space<false>(sh, src, a.k_slots, a.min_score, a.inclusive))
        return;
    select_core(sh, src, a.k_slots, a.min_score, a.inclusive);
}

// Host-side plan: how many CTAs share each level's segment and which clusters they sit in.  A segment gets one CTA
// per 64 K head values (at most a whole cluster); levels are packed into clusters first-fit by decreasing size and the
// ranks left over go to the groups with the longest slices.
static FcosClusterPlan plan_fcos_clusters(const long long* values, int n_levels) {
    FcosClusterPlan plan;
    memset(&plan, 0, sizeof(plan));
    memset(plan.level, -1, sizeof(plan.level));
    int order[DH_MAX_LEVELS], want[DH_MAX_LEVELS];
    for (int l = 0; l < n_levels; ++l) {
        order[l] = l;
        const long long w = (values[l] + 65535) / 65536;
        want[l] = static_cast<int>(w < 1 ? 1 : (w > kSelCluster ? kSelCluster : w));
    }
    for (int i = 1; i < n_levels; ++i)  // insertion sort, longest first (stable)
        for (int j = i; j > 0 && values[order[j]] > values[order[j - 1]]; --j) {
            const int t = order[j];
            order[j] = order[j - 1], order[j - 1] = t;
        }
    int used[DH_MAX_LEVELS] = {0}, n_groups[DH_MAX_LEVELS] = {0};
    int group_level[DH_MAX_LEVELS][kSelCluster], group_members[DH_MAX_LEVELS][kSelCluster];
    for (int i = 0; i < n_levels; ++i) {
        const int l = order[i];
        int t = 0;
        while (t < plan.n_types && used[t] + want[l] > kSelCluster) ++t;
        if (t == plan.n_types) ++plan.n_types;
        group_level[t][n_groups[t]] = l, group_members[t][n_groups[t]] = want[l];
        ++n_groups[t], used[t] += want[l];
    }
    for (int t = 0; t < plan.n_types; ++t) {
        while (used[t] < kSelCluster) {
            int best = -1;
            long long best_slice = 16384;  // shorter slices are not worth another CTA
            for (int q = 0; q < n_groups[t]; ++q) {
                const long long slice = values[group_level[t][q]] / group_members[t][q];
                if (slice > best_slice) best = q, best_slice = slice;
            }
            if (best < 0) break;
            ++group_members[t][best], ++used[t];
        }
        int rank = 0;
        for (int q = 0; q < n_groups[t]; ++q) {
            const int lead = rank;
            for (int m = 0; m < group_members[t][q]; ++m, ++rank) {
                plan.level[t][rank] = static_cast<signed char>(group_level[t][q]);
                plan.lead[t][rank] = static_cast<unsigned char>(lead);
                plan.members[t][rank] = static_cast<unsigned char>(group_members[t][q]);
            }
        }
    }
    return plan;
}

// chunk table of the streaming pre-select: `items[l]` load items per segment of level l, one CTA per kPreChunkItems
static PreselChunks plan_presel_chunks(const long long* items, int n_levels, int batch) {
    PreselChunks ct;
    memset(&ct, 0, sizeof(ct));
    for (int l = 0; l < n_levels; ++l) {
        ct.per_seg[l] = static_cast<int>((items[l] + kPreChunkItems - 1) / kPreChunkItems);
        ct.first[l + 1] = ct.first[l] + ct.per_seg[l] * batch;
    }
    return ct;
}

int launch_fcos_select(dh_handle_s* h, const float* const* pred_levels, int batch, int pad_h, int pad_w, int n_levels,
                       const int32_t* strides, int num_classes, int center, int k, float min_score, int inclusive, float* cand,
                       cudaStream_t st) {
    FcosSelectArgs a;
    memset(&a, 0, sizeof(a));
    a.num_classes = num_classes, a.center = center, a.k_slots = k, a.n_levels = n_levels, a.inclusive = inclusive, a.min_score = min_score;
    a.allow_logit_space = h->fcos_select_mode == 1 ? 0 : 1;
    long long values[DH_MAX_LEVELS];
    for (int l = 0; l < n_levels; ++l) {
        a.head[l] = pred_levels[l];
        a.hl[l] = static_cast<int>(static_cast<double>(pad_h) / strides[l]);
        a.wl[l] = static_cast<int>(static_cast<double>(pad_w) / strides[l]);
        if (a.wl[l] < 1) a.wl[l] = 1, a.hl[l] = 0;
        a.stride[l] = static_cast<float>(strides[l]);
        values[l] = static_cast<long long>(a.hl[l]) * a.wl[l] * (num_classes + 5);
    }
    // same condition as logit_space_applies(): the cluster kernel has no scoring path of its own
    const bool logit_space = !center && min_score > 1.0e-6f && min_score < 0.999f && k > 0;
    // Which selector (all exact, all bit-identical):
    //   pre-select  reads the head once with every SM; for k <= 1024 it is the fastest at every batch size (B200, C4
    //               shape, whole dh_fcos_detect: 276 us for 64 images against 466 us with one CTA per segment);
    //   clusters    pay barriers and leader-only steps per segment: they beat one CTA per segment while the long
    //               segments alone cannot fill the GPU (2.1x faster at 1..8 images, break-even near 48);
    //   one CTA per (image, level) otherwise.
    const bool few_segments = static_cast<long long>(batch) * 3 <= h->sm_count;
    long long longest_pairs = 1;
    for (int l = 0; l < n_levels; ++l) longest_pairs = std::max(longest_pairs, static_cast<long long>(a.hl[l]) * a.wl[l] * num_classes);
    const bool presel_fits = k <= 1024 && (longest_pairs + 31) / 32 <= kPreBitmapWords;
    int launched = 1;
    const bool use_presel = logit_space && presel_fits && (h->fcos_select_mode == 4 || h->fcos_select_mode == 0);
    if (logit_space && !use_presel && (h->fcos_select_mode == 3 || (h->fcos_select_mode == 0 && few_segments))) {
        const FcosClusterPlan plan = plan_fcos_clusters(values, n_levels);
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(static_cast<unsigned>(kSelCluster) * plan.n_types * batch), cfg.blockDim = dim3(kClThreads), cfg.stream = st;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = kSelCluster, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
        cfg.attrs = &attr, cfg.numAttrs = 1;
        DH_CUDA(cudaLaunchKernelEx(&cfg, fcos_select_cluster_kernel, a, plan, batch, cand));
    } else if (use_presel) {
        const int segs = batch * n_levels;
        PreselWork w;
        w.cap = 4096;
        w.bitmap_words = static_cast<int>((longest_pairs + 31) / 32);
        const size_t head_bytes = (static_cast<size_t>(segs) * 16 + 255) & ~size_t(255);
        char* base = static_cast<char*>(scratch(h, head_bytes + static_cast<size_t>(segs) * w.cap * 8));
        if (!base) return DH_ERR_CUDA;
        w.x_est = reinterpret_cast<float*>(base), w.est_bin = reinterpret_cast<int*>(base) + segs;
        w.count = reinterpret_cast<unsigned*>(base) + 2 * segs;
        w.cand_x = reinterpret_cast<float*>(base + head_bytes);
        w.cand_pair = reinterpret_cast<int*>(base + head_bytes) + static_cast<size_t>(segs) * w.cap;
        long long items[DH_MAX_LEVELS];
        for (int l = 0; l < n_levels; ++l) {
            const bool vec = (reinterpret_cast<uintptr_t>(a.head[l]) & 15u) == 0 && (values[l] & 3) == 0;  // fcos_source_vec()
            items[l] = vec ? values[l] >> 2 : static_cast<long long>(a.hl[l]) * a.wl[l] * num_classes;
        }
        const PreselChunks ct = plan_presel_chunks(items, n_levels, batch);
        dim3 grid(batch, n_levels);
        fcos_presel_estimate_kernel<<<grid, kSelThreads, 0, st>>>(a, w);
        if (ct.first[n_levels] > 0) fcos_presel_collect_kernel<<<ct.first[n_levels], kPreThreads, 0, st>>>(a, ct, batch, w);
        const size_t dyn = (static_cast<size_t>(3) * w.cap + w.bitmap_words) * 4;
        DH_ONCE_PER_DEVICE(h) {  // the largest it can be: cap candidates x 3 words + the whole bitmap
            DH_CUDA(cudaFuncSetAttribute(fcos_presel_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (3 * 4096 + kPreBitmapWords) * 4));
        }
        fcos_presel_finish_kernel<<<grid, kSelThreads, dyn, st>>>(a, w, cand);
        DH_CUDA(cudaGetLastError());
        launched = 3;
    } else {
        dim3 grid(batch, n_levels);
        fcos_select_kernel<<<grid, kSelThreads, 0, st>>>(a, cand);
        DH_CUDA(cudaGetLastError());
    }
    h->launches += launched;
    return DH_OK;
}

int launch_select_segs(dh_handle_s* h, const float* dets, int batch, long long n_total, int row_floats, int score_col, const int* seg_off_host,
                       int n_seg, int k, float min_score, int inclusive, float* out, int* overflow, cudaStream_t st, const float* scores) {
    SelSegs segs;
    memset(&segs, 0, sizeof(segs));
    for (int s = 0; s <= n_seg; ++s) segs.off[s] = seg_off_host[s];
    dim3 grid(batch, n_seg);
    select_topk_segs_kernel<<<grid, kSelThreads, 0, st>>>(dets, n_total, row_floats, score_col, segs, k, min_score, inclusive, out, n_seg * k,
                                                          overflow, scores);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

}  // namespace dh

using namespace dh;

extern "C" {

int dh_plan_fcos_select(const long long* values, int n_levels, int batch, signed char* level, unsigned char* lead,
                        unsigned char* members, int* chunk_first) {
    DH_CHECK_ARG(values && n_levels >= 1 && n_levels <= DH_MAX_LEVELS && batch >= 0, "dh_plan_fcos_select: bad arguments");
    for (int l = 0; l < n_levels; ++l) DH_CHECK_ARG(values[l] >= 0, "dh_plan_fcos_select: values[%d] is negative", l);
    const FcosClusterPlan plan = plan_fcos_clusters(values, n_levels);
    for (int t = 0; t < n_levels; ++t)
        for (int r = 0; r < kSelCluster; ++r) {
            if (level) level[t * kSelCluster + r] = t < plan.n_types ? plan.level[t][r] : static_cast<signed char>(-1);
            if (lead) lead[t * kSelCluster + r] = t < plan.n_types ? plan.lead[t][r] : static_cast<unsigned char>(0);
            if (members) members[t * kSelCluster + r] = t < plan.n_types ? plan.members[t][r] : static_cast<unsigned char>(0);
        }
    if (chunk_first) {
        long long items[DH_MAX_LEVELS];
        for (int l = 0; l < n_levels; ++l) items[l] = values[l] >> 2;
        const PreselChunks ct = plan_presel_chunks(items, n_levels, batch);
        for (int l = 0; l <= n_levels; ++l) chunk_first[l] = ct.first[l];
    }
    return plan.n_types;
}

int dh_prediction_to_corners(dh_handle_t h, const float* pred, int batch, int height, int width, int sub, int ch_in, int mode,
                             float stride, float d0, float d1, const float* scales, float* out, void* stream) {
    DH_CHECK_ARG(h && pred && out, "dh_prediction_to_corners: NULL argument");
    DH_CHECK_ARG(mode >= 0 && mode <= 3, "dh_prediction_to_corners: mode %d", mode);
    DH_CHECK_ARG(batch >= 0 && height >= 0 && width >= 0 && sub >= 1 && sub <= 8 && ch_in >= 4, "dh_prediction_to_corners: bad sizes");
    DH_CHECK_ARG(mode != 3 || scales, "dh_prediction_to_corners: mode 3 needs scales");
    DH_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "dh_prediction_to_corners: out must be 16-byte aligned");
    const long long rows = static_cast<long long>(batch) * height * width * sub;
    if (rows == 0) return DH_OK;
    DeviceGuard guard(h);
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
    DeviceGuard guard(h);
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
    return dh::retina_decode_scores(h, pred_levels, batch, pad_h, pad_w, n_levels, strides, n_anchors, anchor_hw_dev, num_classes, dets, nullptr,
                                    nullptr, stream);
}

}  // extern "C"

// dh_retina_decode; `scores` (optional, [B, N]) also receives the score column on its own when the streaming kernel runs
// (*wrote_scores says whether it did).
int dh::retina_decode_scores(dh_handle_s* h, const float* const* pred_levels, int batch, int pad_h, int pad_w, int n_levels,
                             const int32_t* strides, int n_anchors, const float* anchor_hw_dev, int num_classes, float* dets, float* scores,
                             int* wrote_scores, void* stream) {
    if (wrote_scores) *wrote_scores = 0;
    DH_CHECK_ARG(h && pred_levels && strides && anchor_hw_dev && dets, "dh_retina_decode: NULL argument");
    DH_CHECK_ARG(n_levels >= 1 && n_levels <= DH_MAX_LEVELS && n_anchors >= 1 && num_classes >= 1, "dh_retina_decode: bad configuration");
    DeviceGuard guard(h);
    long long n_total = 0;
    for (int l = 0; l < n_levels; ++l)
        n_total += static_cast<long long>(n_anchors) * static_cast<int>(static_cast<double>(pad_h) / strides[l]) *
                   static_cast<int>(static_cast<double>(pad_w) / strides[l]);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ch = num_classes + 4;
    const size_t stream_smem = static_cast<size_t>(kDecWarps) * 2 * 32 * ch * 4 + kDecWarps * 2 * 8;
    if (stream_smem <= 200 * 1024 && batch > 0) {  // streaming kernel: TMA-staged 32-row tiles, one lane per row
        RetinaDecodeArgs a;
        memset(&a, 0, sizeof(a));
        a.n_levels = n_levels, a.n_anchors = n_anchors, a.num_classes = num_classes, a.ch = ch, a.n_total = n_total;
        a.anchor_hw = anchor_hw_dev;
        a.scores = scores;
        a.use_tma = (ch * 4) % 16 == 0 ? 1 : 0;
        long long off = 0, tiles = 0;
        for (int l = 0; l < n_levels; ++l) {
            DH_CHECK_ARG(pred_levels[l] && strides[l] > 0, "dh_retina_decode: level %d", l);
            const int hl = static_cast<int>(static_cast<double>(pad_h) / strides[l]), wl = static_cast<int>(static_cast<double>(pad_w) / strides[l]);
            a.head[l] = pred_levels[l];
            a.cells[l] = hl * wl > 0 ? hl * wl : 1, a.wl[l] = wl > 0 ? wl : 1, a.stride[l] = static_cast<float>(strides[l]);
            a.per_img[l] = n_anchors * hl * wl > 0 ? n_anchors * hl * wl : 1;
            a.div_per_img[l] = make_fastdiv(static_cast<uint32_t>(a.per_img[l]));
            a.div_cells[l] = make_fastdiv(static_cast<uint32_t>(a.cells[l]));
            a.div_wl[l] = make_fastdiv(static_cast<uint32_t>(a.wl[l]));
            DH_CHECK_ARG(static_cast<long long>(batch) * n_anchors * hl * wl < (1ll << 31), "dh_retina_decode: level %d has too many rows", l);
            a.rows[l] = static_cast<long long>(batch) * n_anchors * hl * wl;
            a.level_off[l] = off;
            a.tile_begin[l] = tiles;
            tiles += (a.rows[l] + 31) / 32;
            off += static_cast<long long>(n_anchors) * hl * wl;
            if (reinterpret_cast<uintptr_t>(pred_levels[l]) & 15u) a.use_tma = 0;
        }
        a.tile_begin[n_levels] = tiles;
        if (tiles == 0) return DH_OK;
        DH_ONCE_PER_DEVICE(h) {
            DH_CUDA(cudaFuncSetAttribute(retina_decode_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        }
        int per_sm = 1;
        DH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, retina_decode_stream_kernel, kDecWarps * 32, stream_smem));
        if (per_sm < 1) per_sm = 1;
        long long grid = static_cast<long long>(h->sm_count) * per_sm;
        const long long need = (tiles + kDecWarps - 1) / kDecWarps;
        if (grid > need) grid = need;
        retina_decode_stream_kernel<<<static_cast<unsigned>(grid), kDecWarps * 32, stream_smem, st>>>(a, dets);
        DH_CUDA(cudaGetLastError());
        h->launches += 1;
        if (wrote_scores) *wrote_scores = scores != nullptr;
        return DH_OK;
    }
    long long off = 0;
    for (int l = 0; l < n_levels; ++l) {
        DH_CHECK_ARG(pred_levels[l] && strides[l] > 0, "dh_retina_decode: level %d", l);
        const int hl = static_cast<int>(static_cast<double>(pad_h) / strides[l]), wl = static_cast<int>(static_cast<double>(pad_w) / strides[l]);
        const long long rows = static_cast<long long>(batch) * n_anchors * hl * wl;
        if (rows > 0) {
            retina_decode_kernel<<<grid_for(rows * 32, 256, h->sm_count), 256, 0, st>>>(
                pred_levels[l], batch, n_anchors, hl, wl, num_classes, static_cast<float>(strides[l]),
                anchor_hw_dev + static_cast<long long>(l) * n_anchors * 2, n_total, off, dets);
            DH_CUDA(cudaGetLastError());
            h->launches += 1;
        }
        off += static_cast<long long>(n_anchors) * hl * wl;
    }
    return DH_OK;
}

extern "C" {

int dh_select_topk(dh_handle_t h, const float* dets, int batch, long long n_total, int row_floats, int score_col,
                   const int32_t* seg_off_dev, int n_seg, int k, float min_score, int score_inclusive, float* out,
                   int32_t* out_src, void* stream) {
    DH_CHECK_ARG(h && dets && seg_off_dev && out, "dh_select_topk: NULL argument");
    DH_CHECK_ARG(batch >= 0 && n_total >= 0 && n_total < (1ll << 31) && row_floats >= 1 && score_col >= 0 && score_col < row_floats &&
                     n_seg >= 1 && k >= 1 && k <= 65535,
                 "dh_select_topk: bad sizes");
    if (batch == 0) return DH_OK;
    DeviceGuard guard(h);
    dim3 grid(batch, n_seg);
    select_topk_kernel<<<grid, kSelThreads, 0, static_cast<cudaStream_t>(stream)>>>(dets, n_total, row_floats, score_col, seg_off_dev, k, min_score,
                                                                                   score_inclusive, out, out_src, n_seg * k);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

}  // extern "C"
