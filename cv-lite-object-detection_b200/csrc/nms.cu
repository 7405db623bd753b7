// Greedy NMS on the device: sort by score, 64x64-blocked IoU bit matrix, single-warp sweep.
//
// Replaces RetinaNet.cpu_nms (RetinaNet/retinanet_module.py:453-481, class-agnostic, keeps `ovr <= thr`)
// and the per-class / capped NMS that FCOS inference delegates to TensorFlow's
// combined_non_max_suppression (FCOS/infer_fcos.py:58-61).  Both are one sweep in global score order:
// a per-class NMS is the same sweep with "same class" folded into the suppression predicate, and the
// per-class / total caps are counters inside the sweep.
//
//   kernel 1  nms_sort_kernel   one CTA per image: (score, index) keys -> bitonic sort in shared memory
//                               (descending score, ascending index on ties = a stable sort), threshold,
//                               gather boxes into score order.
//   kernel 2  nms_mask_kernel   grid (col blocks, row blocks, images): bit j of mask[i][w] says box
//                               64*w+j (later in order) is suppressed by box i; warp-ballot builds words.
//   kernel 3  nms_sweep_kernel  one CTA per image walks the order 64 boxes at a time: one thread resolves the
//                               block's greedy chain from its diagonal mask words, all threads OR the kept rows
//                               into the shared-memory `removed` bit vector.
//
// Float arithmetic follows the reference operation order with contraction disabled, so kept indices
// are bit-identical to the NumPy loop.
#include <cstring>

#include "dh_common.cuh"
#include "dh_host.h"

namespace dh {

constexpr int kNmsMaxN = 16384;  // candidates per image (shared-memory sort)
constexpr int kSortThreads = 1024;

struct NmsParams {
    int n_max;        // row stride of dets in boxes: dets is [B, n_max, row_floats]
    int row_floats;   // >= 5 (mode 0) or >= 6 (per-class)
    int per_class;    // 0: class-agnostic cpu_nms rule; 1: same-class IoU > thr rule
    int inclusive;    // 1: keep score >= min_score; 0: keep score > min_score
    float iou_thr, min_score;
    int max_per_class, max_total, num_classes;
    int max_out;  // capacity of keep[b]
};

// ---- kernel 1 -----------------------------------------------------------------------------------------
// Keys are distinct (the candidate index sits in the low word), so the sorted order is unique and any correct sort
// gives the same result.  The fast path is a bucket sort in shared memory with six block barriers: 4096 buckets over
// the span of the score keys (descending), a prefix over the bucket counts, a scatter into the bucket segments, and
// each key's rank inside its bucket by counting the larger keys of the segment.  Scores piled onto few buckets (a
// segment longer than kSortMaxBucket) fall back to the bitonic network, which does 91 barrier-separated stages for
// 8192 keys.
constexpr int kSortBuckets = 4096;
constexpr int kSortMaxBucket = 384;

__device__ __forceinline__ unsigned long long nms_sort_key(const float* d, int row_floats, int i, int inclusive, float min_score) {
    const float s = d[static_cast<long long>(i) * row_floats + 4];
    const bool ok = inclusive ? (s >= min_score) : (s > min_score);
    if (!ok) return 0ull;  // sorts to the end
    // monotone map float -> uint (handles negatives too), then descending sort of the 64-bit key
    unsigned u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    const unsigned long long k = (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned>(i));
    return k == 0ull ? 1ull : k;
}

// ---- slab masks: a conservative "can this pair be suppressed at all" test in three instructions -------------------
// Each axis of an image's candidates is cut into 32 slabs over the candidates' own coordinate range; a box carries, per
// axis, the bit mask of the slabs its CORE touches -- the box shrunk about its centre by the factor (1 - thr).  Why
// cores: for two proper boxes, inter / union > thr needs inter > thr * max(area_a, area_c); inter = ix * iy with
// iy <= min(h_a, h_c), so the overlap on the x axis must exceed thr * max(w_a, w_c) (same on y), and ix is at most
// (w_a + w_c) / 2 - |centre distance|: the centres are closer than (1 - thr) (w_a + w_c) / 2, i.e. the cores intersect.
// The slab index is a monotone function of the coordinate, so two intervals that intersect share a slab:
// (mask_a & mask_c) == 0 on either axis proves that neither predicate suppresses the pair (both compute
// inter / (union [+ 1e-8]); thr = 0 gives the plain "do they overlap" test).  With thr = 0.5 a core has a quarter of the
// box's area: where boxes are large against the image (coarse pyramid levels) nearly every pair overlaps, few cores do.
// Boxes that are not proper (inverted corners, zero extent, NaN) get all-ones masks: they always take the exact predicate.
__device__ __forceinline__ unsigned ordered_bits(float v) {  // monotone float -> uint
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_float(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
__device__ __forceinline__ unsigned slab_range_mask(float lo_v, float hi_v, float lo, float scale) {
    const int s0 = min(max(__float2int_rd((lo_v - lo) * scale), 0), 31), s1 = min(max(__float2int_rd((hi_v - lo) * scale), 0), 31);
    return (s1 >= 31 ? 0xffffffffu : ((1u << (s1 + 1)) - 1u)) & ~((1u << s0) - 1u);
}

__global__ void __launch_bounds__(kSortThreads) nms_sort_kernel(const float* __restrict__ dets, const int* __restrict__ n_valid,
                                                                NmsParams p, int n_pow2, int force_bitonic, float4* __restrict__ sorted_boxes,
                                                                int* __restrict__ sorted_cls, int* __restrict__ order,
                                                                int* __restrict__ n_cand, uint2* __restrict__ sorted_slabs,
                                                                int* __restrict__ filter_ok) {
    extern __shared__ unsigned long long keys[];  // [n_pow2]
    __shared__ unsigned bucket_start[kSortBuckets], bucket_fill[kSortBuckets];
    __shared__ unsigned warp_tot[kSortThreads / 32];
    __shared__ unsigned key_min, key_max, biggest;
    __shared__ unsigned rng[4];  // ordered bits of {min, max} of axis 0 and of axis 1 over the valid candidates
    __shared__ unsigned n_improper;  // valid candidates that are not proper boxes (they defeat the slab filter)
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = n_valid ? min(n_valid[b], p.n_max) : p.n_max;
    const float* d = dets + static_cast<long long>(b) * p.n_max * p.row_floats;
    auto emit = [&](int pos, unsigned long long key) {
        const int src = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(key & 0xFFFFFFFFull));
        const float* r = d + static_cast<long long>(src) * p.row_floats;
        sorted_boxes[static_cast<long long>(b) * p.n_max + pos] = make_float4(r[0], r[1], r[2], r[3]);
        sorted_cls[static_cast<long long>(b) * p.n_max + pos] = p.per_class ? __float2int_rz(r[5]) : 0;
        order[static_cast<long long>(b) * p.n_max + pos] = src;
        if (sorted_slabs) {
            // only a proper box (c2 > c0, c3 > c1, finite) gets tight masks: for two such boxes with an empty intersection both
            // suppression predicates answer "no" (positive areas, positive denominator); anything else -- inverted corners,
            // zero extent, NaN -- keeps all-ones masks and always takes the exact predicate
            uint2 m = make_uint2(0xffffffffu, 0xffffffffu);
            if (r[2] > r[0] && r[3] > r[1] && fabsf(r[0]) < 3.0e38f && fabsf(r[1]) < 3.0e38f && fabsf(r[2]) < 3.0e38f && fabsf(r[3]) < 3.0e38f) {
                const float lo0 = ordered_float(rng[0]), hi0 = ordered_float(rng[1]), lo1 = ordered_float(rng[2]), hi1 = ordered_float(rng[3]);
                const float sc0 = hi0 > lo0 ? 32.0f / (hi0 - lo0) : 0.f, sc1 = hi1 > lo1 ? 32.0f / (hi1 - lo1) : 0.f;
                if (sc0 < 3.0e38f && sc1 < 3.0e38f) {
                    // the box shrunk about its centre by (1 - thr): see "cores" above.  thr is taken 1e-4 low (the rounding
                    // of the exact predicate) and the core a little wide (the rounding of these few operations)
                    const float keep = 1.0f - fminf(fmaxf(p.iou_thr * (1.0f - 1.0e-4f), 0.0f), 1.0f);
                    const float h0 = 0.5f * keep * (r[2] - r[0]) * (1.0f + 1.0e-5f) + 1.0e-5f * (hi0 - lo0);
                    const float h1 = 0.5f * keep * (r[3] - r[1]) * (1.0f + 1.0e-5f) + 1.0e-5f * (hi1 - lo1);
                    const float m0 = 0.5f * (r[0] + r[2]), m1 = 0.5f * (r[1] + r[3]);
                    m = make_uint2(slab_range_mask(m0 - h0, m0 + h0, lo0, sc0), slab_range_mask(m1 - h1, m1 + h1, lo1, sc1));
                }
            }
            sorted_slabs[static_cast<long long>(b) * p.n_max + pos] = m;
        }
    };
    // span of the score keys
    if (tid == 0) key_min = 0xFFFFFFFFu, key_max = 0u, biggest = 0u, rng[0] = rng[2] = 0xFFFFFFFFu, rng[1] = rng[3] = 0u, n_improper = 0u;
    for (int i = tid; i < kSortBuckets; i += kSortThreads) bucket_fill[i] = 0u;
    __syncthreads();
    unsigned lo = 0xFFFFFFFFu, hi = 0u;
    unsigned r0 = 0xFFFFFFFFu, r1 = 0u, r2 = 0xFFFFFFFFu, r3 = 0u, bad = 0u;
    for (int i = tid; i < n; i += kSortThreads) {
        const unsigned long long k = nms_sort_key(d, p.row_floats, i, p.inclusive, p.min_score);
        if (k) {
            lo = min(lo, static_cast<unsigned>(k >> 32)), hi = max(hi, static_cast<unsigned>(k >> 32));
            if (sorted_slabs) {
                const float* r = d + static_cast<long long>(i) * p.row_floats;
                const float c0 = r[0], c1 = r[1], c2 = r[2], c3 = r[3];
                if (c0 == c0 && c1 == c1 && c2 == c2 && c3 == c3 && fabsf(c0) < 3.0e38f && fabsf(c1) < 3.0e38f && fabsf(c2) < 3.0e38f && fabsf(c3) < 3.0e38f) {
                    r0 = min(r0, ordered_bits(fminf(c0, c2))), r1 = max(r1, ordered_bits(fmaxf(c0, c2)));
                    r2 = min(r2, ordered_bits(fminf(c1, c3))), r3 = max(r3, ordered_bits(fmaxf(c1, c3)));
                }
                bad += !(c2 > c0 && c3 > c1 && fabsf(c0) < 3.0e38f && fabsf(c1) < 3.0e38f && fabsf(c2) < 3.0e38f && fabsf(c3) < 3.0e38f);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)), hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    if (lane == 0 && lo <= hi) atomicMin(&key_min, lo), atomicMax(&key_max, hi);
    if (sorted_slabs) {
        r0 = __reduce_min_sync(0xffffffffu, r0), r1 = __reduce_max_sync(0xffffffffu, r1);
        r2 = __reduce_min_sync(0xffffffffu, r2), r3 = __reduce_max_sync(0xffffffffu, r3);
        bad = __reduce_add_sync(0xffffffffu, bad);
        if (lane == 0) atomicMin(&rng[0], r0), atomicMax(&rng[1], r1), atomicMin(&rng[2], r2), atomicMax(&rng[3], r3), atomicAdd(&n_improper, bad);
    }
    __syncthreads();
    const unsigned kmin = key_min, span = key_max >= kmin ? key_max - kmin : 0u;
    const int shift = max(0, (32 - __clz(span)) - 12);  // (u - kmin) >> shift < 4096
    auto bucket_of = [&](unsigned long long k) { return kSortBuckets - 1 - static_cast<int>((static_cast<unsigned>(k >> 32) - kmin) >> shift); };
    for (int i = tid; i < n; i += kSortThreads) {
        const unsigned long long k = nms_sort_key(d, p.row_floats, i, p.inclusive, p.min_score);
        if (k) atomicAdd(&bucket_fill[bucket_of(k)], 1u);
    }
    __syncthreads();
    // exclusive prefix over the buckets (thread t owns 4 consecutive buckets)
    unsigned c[kSortBuckets / kSortThreads], mine = 0, most = 0;
#pragma unroll
    for (int q = 0; q < kSortBuckets / kSortThreads; ++q) c[q] = bucket_fill[tid * (kSortBuckets / kSortThreads) + q], mine += c[q], most = max(most, c[q]);
    unsigned incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) most = max(most, __shfl_xor_sync(0xffffffffu, most, o));
    if (lane == 31) warp_tot[warp] = incl;
    if (lane == 0) atomicMax(&biggest, most);
    __syncthreads();
    unsigned before = 0, m_total = 0;
    for (int w = 0; w < kSortThreads / 32; ++w) {
        const unsigned v = warp_tot[w];
        if (w < warp) before += v;
        m_total += v;
    }
    const int m = static_cast<int>(m_total);
    if (biggest <= kSortMaxBucket && !force_bitonic) {
        unsigned run = before + incl - mine;
#pragma unroll
        for (int q = 0; q < kSortBuckets / kSortThreads; ++q) {
            bucket_start[tid * (kSortBuckets / kSortThreads) + q] = run, bucket_fill[tid * (kSortBuckets / kSortThreads) + q] = run;
            run += c[q];
        }
        __syncthreads();
        for (int i = tid; i < n; i += kSortThreads) {
            const unsigned long long k = nms_sort_key(d, p.row_floats, i, p.inclusive, p.min_score);
            if (k) keys[atomicAdd(&bucket_fill[bucket_of(k)], 1u)] = k;
        }
        __syncthreads();
        if (tid == 0) n_cand[b] = m;
        if (tid == 0 && filter_ok) filter_ok[b] = n_improper * 8u <= static_cast<unsigned>(m);  // improper boxes take every pair: beyond 1 in 8 the filter costs more than it saves
        for (int q = tid; q < m; q += kSortThreads) {
            const unsigned long long k = keys[q];
            const int bk = bucket_of(k);
            const unsigned s0 = bucket_start[bk], s1 = bucket_fill[bk];  // after the scatter the fill cursor is the segment's end
            unsigned r = 0;
            for (unsigned j = s0; j < s1; ++j) r += keys[j] > k ? 1u : 0u;
            emit(static_cast<int>(s0 + r), k);
        }
        return;
    }
    // bitonic network over n_pow2 keys
    __syncthreads();
    for (int i = tid; i < n_pow2; i += kSortThreads) keys[i] = i < n ? nms_sort_key(d, p.row_floats, i, p.inclusive, p.min_score) : 0ull;
    __syncthreads();
    for (int size = 2; size <= n_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (n_pow2 >> 1); t += kSortThreads) {
                const int lo_i = ((t / stride) * (stride << 1)) + (t % stride);
                const int hi_i = lo_i + stride;
                const bool desc = ((lo_i & size) == 0);
                const unsigned long long a = keys[lo_i], cc = keys[hi_i];
                if (desc ? (a < cc) : (a > cc)) keys[lo_i] = cc, keys[hi_i] = a;
            }
            __syncthreads();
        }
    }
    // valid keys are a prefix (m of them); gather boxes in score order
    if (tid == 0) n_cand[b] = m;
    if (tid == 0 && filter_ok) filter_ok[b] = n_improper * 8u <= static_cast<unsigned>(m);
    for (int i = tid; i < m; i += kSortThreads) emit(i, keys[i]);
}

// ---- suppression predicates ---------------------------------------------------------------------------
// RetinaNet/retinanet_module.py:461-479: areas from raw corners, ovr = inter / (a_i + a_j - inter + 1e-8),
// the later box survives only if ovr <= thr (a NaN does not survive).
__device__ __forceinline__ float box_area(const float4& b) { return fmul(fsub(b.z, b.x), fsub(b.w, b.y)); }
__device__ __forceinline__ bool suppress_agnostic(const float4& a, float area_a, const float4& c, float area_c, float thr) {
    const float w_raw = fsub(fminf(a.z, c.z), fmaxf(a.x, c.x)), h_raw = fsub(fminf(a.w, c.w), fmaxf(a.y, c.y));
    if (!(w_raw > 0.f && h_raw > 0.f)) {
        // no overlap: inter = 0, so ovr = 0 / (area_a + area_c + 1e-8), which is <= thr unless thr < 0 or the
        // denominator is 0 / NaN (degenerate boxes) -- the same answer as the full expression, without the division
        const float den = fadd(fadd(area_a, area_c), 1e-8f);
        return thr < 0.f || !(den < 0.f || den > 0.f);
    }
    const float inter = fmul(w_raw, h_raw);
    const float den = fadd(fsub(fadd(area_a, area_c), inter), 1e-8f);
    // ovr = inter / den decides; away from the threshold the division-free comparison gives the same answer
    // (|inter - thr*den| far above the rounding error of either side), so the IEEE division runs only near it
    const float rhs = fmul(thr, den);
    const float gap = fabsf(fsub(inter, rhs)), tol = fmul(1.0e-6f, fadd(fabsf(inter), fabsf(rhs)));
    if (den > 0.f && gap > tol) return inter > rhs;
    const float ovr = fdiv(inter, den);
    return !(ovr <= thr);
}
__device__ __forceinline__ bool suppress_agnostic(const float4& a, const float4& c, float thr) {
    return suppress_agnostic(a, box_area(a), c, box_area(c), thr);
}
// combined-NMS rule (TensorFlow's op, restated; parity unpinned): corner order normalised, degenerate
// boxes never suppress, IoU > thr suppresses.
__device__ __forceinline__ bool suppress_iou(const float4& a0, const float4& c0, float thr) {
    const float ay1 = fminf(a0.x, a0.z), ax1 = fminf(a0.y, a0.w), ay2 = fmaxf(a0.x, a0.z), ax2 = fmaxf(a0.y, a0.w);
    const float cy1 = fminf(c0.x, c0.z), cx1 = fminf(c0.y, c0.w), cy2 = fmaxf(c0.x, c0.z), cx2 = fmaxf(c0.y, c0.w);
    const float area_a = fmul(fsub(ay2, ay1), fsub(ax2, ax1)), area_c = fmul(fsub(cy2, cy1), fsub(cx2, cx1));
    if (!(area_a > 0.f) || !(area_c > 0.f)) return false;
    const float ih = fmaxf(0.f, fsub(fminf(ay2, cy2), fmaxf(ay1, cy1)));
    const float iw = fmaxf(0.f, fsub(fminf(ax2, cx2), fmaxf(ax1, cx1)));
    const float inter = fmul(ih, iw);
    const float uni = fsub(fadd(area_a, area_c), inter);
    return uni > 0.f && fdiv(inter, uni) > thr;
}

// ---- kernel 2 -----------------------------------------------------------------------------------------
// One CTA takes the 64 boxes of column block cb against four row blocks (256 threads, one row each).  Nearly all of the
// 400 M pairs of a 64 x 5 000-candidate batch cannot reach the threshold, so the exact predicate must not run for them:
// the column block's slab masks are turned into "which of my 64 boxes touch slab s" words (64 ballots per axis), a row
// ORs the words of the slabs it touches on each axis and ANDs the two -- the 64-bit set of columns whose cores meet its
// own, for ~30 instructions -- and only those take the exact predicate (identical arithmetic, identical bits).  A pair
// that is filtered out has two boxes of positive finite area whose IoU is provably below the threshold, for which both
// predicates answer "not suppressed" (a negative threshold switches the filter off).
constexpr int kMaskRowBlocks = 4;
__global__ void __launch_bounds__(64 * kMaskRowBlocks) nms_mask_kernel(const float4* __restrict__ sorted_boxes, const int* __restrict__ sorted_cls,
                                                                       const uint2* __restrict__ sorted_slabs, const int* __restrict__ filter_ok,
                                                                       const int* __restrict__ n_cand, NmsParams p, int words, int dense_at,
                                                                       unsigned long long* __restrict__ mask) {
    const int b = blockIdx.z, cb = blockIdx.x, rb = blockIdx.y * kMaskRowBlocks + (threadIdx.x >> 6);
    const int m = n_cand[b];
    if (cb * 64 >= m || blockIdx.y * kMaskRowBlocks > cb) return;  // only later boxes (j > i) can be suppressed
    __shared__ float4 cbox[64];
    __shared__ float carea[64];
    __shared__ int ccls[64];
    __shared__ unsigned slab_cols[2][32][2];  // [axis][slab][column half]: columns of the block that touch the slab
    const int tid = threadIdx.x, t = tid & 63, lane = tid & 31;
    const long long base = static_cast<long long>(b) * p.n_max;
    const bool filter = sorted_slabs != nullptr && !(p.iou_thr < 0.f) && filter_ok[b] != 0;
    __shared__ uint2 cslab[64];
    // this thread's row: asked for ahead of the barriers below
    const int i = rb * 64 + t;
    const bool row_ok = rb <= cb && i < m;
    const float4 a = row_ok ? sorted_boxes[base + i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int ac = row_ok ? sorted_cls[base + i] : 0;
    const uint2 row_sm = (row_ok && filter) ? sorted_slabs[base + i] : make_uint2(0u, 0u);
    if (tid < 64) {
        const int j = cb * 64 + t;
        uint2 sm = make_uint2(0u, 0u);
        if (j < m) {
            const float4 c = sorted_boxes[base + j];
            cbox[t] = c, carea[t] = box_area(c), ccls[t] = sorted_cls[base + j];
            if (filter) sm = sorted_slabs[base + j];
        }
        cslab[t] = sm;
    }
    __syncthreads();
    if (filter) {  // 128 ballots, dealt to the warps: warp w takes slabs 4w .. 4w + 3 of both axes for both column halves
        const int w = tid >> 5;
        for (int s = w; s < 32; s += (64 * kMaskRowBlocks) >> 5) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const uint2 sm = cslab[32 * hf + lane];
                const unsigned b0 = __ballot_sync(0xffffffffu, (sm.x >> s) & 1u), b1 = __ballot_sync(0xffffffffu, (sm.y >> s) & 1u);
                if (lane == 0) slab_cols[0][s][hf] = b0, slab_cols[1][s][hf] = b1;
            }
        }
        __syncthreads();
    }
    if (!row_ok) return;
    const float area_a = box_area(a);
    const int jn = min(64, m - cb * 64);
    unsigned long long todo = jn >= 64 ? ~0ull : ((1ull << jn) - 1ull);
    if (rb == cb) todo &= t >= 63 ? 0ull : (~0ull << (t + 1));
    if (filter) {
        const uint2 sm = row_sm;
        unsigned x0 = 0u, x1 = 0u, y0 = 0u, y1 = 0u;
        for (unsigned r = sm.x; r; r &= r - 1) {
            const int s = __ffs(r) - 1;
            x0 |= slab_cols[0][s][0], x1 |= slab_cols[0][s][1];
        }
        for (unsigned r = sm.y; r; r &= r - 1) {
            const int s = __ffs(r) - 1;
            y0 |= slab_cols[1][s][0], y1 |= slab_cols[1][s][1];
        }
        todo &= (static_cast<unsigned long long>(x1 & y1) << 32) | (x0 & y0);
    }
    // the columns left take the exact predicate, after one more cheap and safe reject for two proper boxes: IoU cannot
    // exceed min(area) / max(area), so a pair whose areas differ by more than the threshold allows is never suppressed
    // (0.999 covers the rounding of either side).  Two 32-bit halves: no 64-bit shifts in the loops.
    const bool a_proper = a.z > a.x && a.w > a.y;
    const float k_area = (filter && p.iou_thr > 0.f) ? 0.999f * p.iou_thr : -1.0f;  // < 0: reject off
    unsigned half[2] = {static_cast<unsigned>(todo), static_cast<unsigned>(todo >> 32)}, out[2] = {0u, 0u};
    // A warp whose rows keep most of their columns (no filter, or boxes that are not proper) walks the columns in step --
    // every lane reads the same column box, a broadcast -- and each lane skips what it does not need; otherwise every
    // lane walks its own few columns.
    const int most = __reduce_max_sync(__activemask(), __popc(half[0]) + __popc(half[1]));
    if (!filter || most > dense_at) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int t_end = min(32, jn - 32 * hh);
            for (int tt = 0; tt < t_end; ++tt) {
                if (!((half[hh] >> tt) & 1u)) continue;
                const int c = 32 * hh + tt;
                if (a_proper && k_area > 0.f && cbox[c].z > cbox[c].x && cbox[c].w > cbox[c].y && fminf(area_a, carea[c]) < k_area * fmaxf(area_a, carea[c])) continue;
                const bool sup = p.per_class ? (ccls[c] == ac && suppress_iou(a, cbox[c], p.iou_thr)) : suppress_agnostic(a, area_a, cbox[c], carea[c], p.iou_thr);
                if (sup) out[hh] |= 1u << tt;
            }
        }
    } else {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            for (unsigned r = half[hh]; r; r &= r - 1) {
                const int c = 32 * hh + __ffs(r) - 1;
                const float4 cbx = cbox[c];
                const float ca = carea[c];
                if (a_proper && k_area > 0.f && cbx.z > cbx.x && cbx.w > cbx.y && fminf(area_a, ca) < k_area * fmaxf(area_a, ca)) continue;
                const bool sup = p.per_class ? (ccls[c] == ac && suppress_iou(a, cbx, p.iou_thr)) : suppress_agnostic(a, area_a, cbx, ca, p.iou_thr);
                if (sup) out[hh] |= 1u << (c & 31);
            }
        }
    }
    const unsigned long long bits = (static_cast<unsigned long long>(out[1]) << 32) | out[0];
    // block-major layout [image][row block][column block][row in block]: a row block's words right of the diagonal are
    // one contiguous range (a single bulk copy for the sweep) and this store is coalesced
    mask[(((static_cast<long long>(b) * words + rb) * words + cb) << 6) + t] = bits;
}

// ---- kernel 3 -----------------------------------------------------------------------------------------
// One CTA per image walks the score order 64 boxes at a time.  Inside a 64-block the greedy dependency needs only
// the block's diagonal mask words, so one thread resolves it from shared memory (64 short steps); the rows of the
// boxes it kept are then OR-ed into the `removed` bit vector by all threads in parallel (coalesced along the row).
// Only mask words at or right of the diagonal are ever read, so nms_mask_kernel never has to clear the rest.
constexpr int kSweepThreads = 256;
constexpr int kSweepStageWords = 96;  // a row block's mask words are staged in shared memory when words <= this (n <= 6144)
constexpr int kSweepDepth = 3;        // stage buffers: blocks w + 1 and w + 2 are in flight while block w is swept
__global__ void __launch_bounds__(kSweepThreads) nms_sweep_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ sorted_cls,
                                                                  const int* __restrict__ order, const int* __restrict__ n_cand, NmsParams p,
                                                                  int words, int staged, int serial_chain, int* __restrict__ keep,
                                                                  int* __restrict__ n_keep) {
    // dynamic shared memory: [kSweepDepth][words][64] staged mask words (when `staged`), then class_count (caps)
    extern __shared__ __align__(128) unsigned char sweep_smem[];
    __shared__ unsigned long long removed[kNmsMaxN / 64];
    __shared__ int bcls[2][64];
    __shared__ unsigned long long s_km;
    __shared__ uint64_t bars[kSweepDepth];
    __shared__ int s_kept, s_stop;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int m = n_cand[b];
    const int nblk = (m + 63) >> 6;
    const long long base = static_cast<long long>(b) * p.n_max;
    const bool caps = p.per_class && p.max_per_class > 0;
    const int max_total = p.max_total > 0 ? min(p.max_total, p.max_out) : p.max_out;
    const size_t buf_words = staged ? static_cast<size_t>(words) * 64 : 0;
    unsigned long long* stage = reinterpret_cast<unsigned long long*>(sweep_smem);
    int* class_count = reinterpret_cast<int*>(sweep_smem + kSweepDepth * buf_words * 8);
    // mask words of row block rb from its diagonal column on: [column block - rb][row in block]
    auto block_words = [&](int rb) { return mask + (((static_cast<long long>(b) * words + rb) * words + rb) << 6); };
    if (caps)
        for (int c = tid; c < p.num_classes; c += kSweepThreads) class_count[c] = 0;
    for (int w = tid; w < nblk; w += kSweepThreads) removed[w] = 0ull;
    if (tid == 0) {
        s_kept = 0, s_stop = (max_total <= 0);
        for (int q = 0; q < kSweepDepth; ++q) mbar_init(&bars[q], 1);
        mbar_init_fence();
    }
    if (caps && tid < min(64, m)) bcls[0][tid] = sorted_cls[base + tid];
    __syncthreads();
    auto prefetch_block = [&](int blk) {  // thread 0: one TMA bulk copy per block into stage[blk % depth]
        if (!staged || blk >= nblk) return;
        const uint32_t bytes = static_cast<uint32_t>(nblk - blk) * 512u;
        const int buf = blk % kSweepDepth;
        mbar_expect_tx(&bars[buf], bytes);
        bulk_g2s(stage + buf * buf_words, block_words(blk), bytes, &bars[buf]);
    };
    if (tid == 0) prefetch_block(0), prefetch_block(1);
    int w = 0;
    for (; w < nblk && !s_stop; ++w) {
        const int first = w << 6, nb = min(64, m - first), cur = w & 1;
        const int kept_before = s_kept;  // written before the barriers that end the previous block
        const int buf = w % kSweepDepth;
        const int my_order = tid < nb ? __ldg(order + base + first + tid) : 0;  // issued early: its latency hides behind the chain walk
        if (staged) mbar_wait(&bars[buf], static_cast<uint32_t>((w / kSweepDepth) & 1));  // requested two blocks ago
        const unsigned long long* src = staged ? stage + buf * buf_words : block_words(w);
        if (tid < 32 && !caps && !serial_chain) {
            // The greedy chain of the block, resolved in parallel rounds by one warp (lane l holds diagonal rows l and l + 32).
            // K = kept for sure, U = undecided.  A box of U no box of K | U suppresses is kept; one a box of K suppresses is
            // gone.  The lowest box of U has only decided boxes before it, so every round decides at least that one; with
            // the few suppressions a 64-block holds, two or three rounds settle it (the serial walk takes 64 dependent steps).
            const unsigned long long d_lo = src[lane], d_hi = src[lane + 32];
            const unsigned long long valid = nb >= 64 ? ~0ull : ((1ull << nb) - 1ull);
            unsigned long long K = 0ull, U = valid & ~removed[w];
            while (U) {
                const unsigned long long maybe = K | U;
                const unsigned long long vm = (((maybe >> lane) & 1ull) ? d_lo : 0ull) | (((maybe >> (lane + 32)) & 1ull) ? d_hi : 0ull);
                const unsigned long long vk = (((K >> lane) & 1ull) ? d_lo : 0ull) | (((K >> (lane + 32)) & 1ull) ? d_hi : 0ull);
                const unsigned long long s_maybe = (static_cast<unsigned long long>(__reduce_or_sync(0xffffffffu, static_cast<unsigned>(vm >> 32))) << 32) |
                                                   __reduce_or_sync(0xffffffffu, static_cast<unsigned>(vm));
                const unsigned long long s_kept = (static_cast<unsigned long long>(__reduce_or_sync(0xffffffffu, static_cast<unsigned>(vk >> 32))) << 32) |
                                                  __reduce_or_sync(0xffffffffu, static_cast<unsigned>(vk));
                const unsigned long long new_k = U & ~s_maybe, new_r = U & s_kept;
                K |= new_k;
                U &= ~(new_k | new_r);
            }
            int kept = kept_before + __popcll(K);
            bool live = true;
            if (kept >= max_total) {  // the output is full inside this block: keep its first boxes only
                int room = max_total - kept_before;
                unsigned long long kk = 0ull;
                for (unsigned long long r = K; r && room > 0; r &= r - 1, --room) kk |= r & (~r + 1ull);
                K = kk, kept = max_total, live = false;
            }
            if (tid == 0) {
                s_km = K, s_kept = kept;
                if (!live) s_stop = 1;
            }
        } else if (tid == 0 && (caps || serial_chain)) {
            // one thread walks the greedy chain with the 64 diagonal words in registers: each step is a bit test and a
            // predicated OR, no memory access on the dependent path
            unsigned long long d[64];
#pragma unroll
            for (int r = 0; r < 64; ++r) d[r] = src[r];
            unsigned long long rem = removed[w], km = 0ull;
            int kept = kept_before;
            bool live = true;
#pragma unroll
            for (int r = 0; r < 64; ++r) {
                bool take = live && r < nb && !((rem >> r) & 1ull);
                if (caps && take) {
                    const int cc = bcls[cur][r];
                    if (cc >= 0 && cc < p.num_classes) {
                        if (class_count[cc] >= p.max_per_class) take = false;  // over the per-class cap: never selected, suppresses nobody
                        else class_count[cc] += 1;
                    }
                }
                if (take) {
                    km |= 1ull << r;
                    rem |= d[r];
                    if (++kept >= max_total) live = false;
                }
            }
            s_km = km, s_kept = kept;
            if (!live) s_stop = 1;
        } else if (caps && tid >= 64 && tid < 128 && first + tid < m) {
            bcls[cur ^ 1][tid - 64] = sorted_cls[base + first + tid];  // the next block's classes
        }
        __syncthreads();
        const unsigned long long km = s_km;
        if (tid < nb && ((km >> tid) & 1ull))
            keep[static_cast<long long>(b) * p.max_out + kept_before + __popcll(km & ((1ull << tid) - 1ull))] = my_order;
        if (!s_stop && km) {
            // OR the kept rows into `removed`: one warp per mask column, lane l takes rows l and l + 32 (coalesced /
            // conflict-free), two redux instructions combine the warp and lane 0 updates the word (no atomics needed)
            const bool k0 = (km >> lane) & 1ull, k1 = (km >> (lane + 32)) & 1ull;
            for (int col = 1 + (tid >> 5); col < nblk - w; col += kSweepThreads / 32) {
                const unsigned long long* cw = src + (static_cast<size_t>(col) << 6);
                const unsigned long long v = (k0 ? cw[lane] : 0ull) | (k1 ? cw[lane + 32] : 0ull);
                const unsigned lo = __reduce_or_sync(0xffffffffu, static_cast<unsigned>(v));
                const unsigned hi = __reduce_or_sync(0xffffffffu, static_cast<unsigned>(v >> 32));
                if (lane == 0 && (lo | hi)) removed[w + col] |= (static_cast<unsigned long long>(hi) << 32) | lo;
            }
        }
        __syncthreads();
        if (tid == 0) prefetch_block(w + 2);  // buffer (w + 2) % depth was drained one block ago
    }
    if (tid == 0) {
        n_keep[b] = s_kept;
        // blocks < w were waited for; an early stop leaves the prefetches of blocks w and w + 1 in flight: let them land
        // before the CTA (and its shared memory) goes away
        if (staged)
            for (int blk = w; blk <= w + 1; ++blk)
                if (blk < nblk) mbar_wait(&bars[blk % kSweepDepth], static_cast<uint32_t>((blk / kSweepDepth) & 1));
    }
}

// ---- lazy variant for capped outputs: kernels 2 + 3 fused, one CTA per image --------------------------------
// With an output cap (FCOS: 100 per class / 100 total) the sweep ends after a few 64-blocks, so suppression is
// evaluated only where the sweep actually goes: a block's candidates are first tested against the boxes kept so
// far (held in shared memory, at most kLazyMaxKept), then the block's own 64x64 tile is evaluated and one thread
// resolves its greedy chain.  The 12.5 M-pair mask matrix of a 5 000-candidate image is never formed.
constexpr int kLazyThreads = 256;
constexpr int kLazyMaxKept = 1024;
__device__ __forceinline__ bool suppresses(const NmsParams& p, const float4& a, int ac, const float4& c, int cc) {
    if (p.per_class) return cc == ac && suppress_iou(a, c, p.iou_thr);
    return suppress_agnostic(a, c, p.iou_thr);
}
__global__ void __launch_bounds__(kLazyThreads) nms_lazy_kernel(const float4* __restrict__ sorted_boxes, const int* __restrict__ sorted_cls,
                                                                const int* __restrict__ order, const int* __restrict__ n_cand, NmsParams p,
                                                                int* __restrict__ keep, int* __restrict__ n_keep) {
    extern __shared__ int class_count[];  // [num_classes] when per-class caps are on
    __shared__ float4 kbox[kLazyMaxKept];
    __shared__ int kcls[kLazyMaxKept];
    __shared__ unsigned long long diag[64];
    __shared__ float4 bbox[64];
    __shared__ int bcls[64];
    __shared__ unsigned long long s_dead, s_km;
    __shared__ int s_kept, s_stop;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int m = n_cand[b];
    const int nblk = (m + 63) >> 6;
    const long long base = static_cast<long long>(b) * p.n_max;
    const bool caps = p.per_class && p.max_per_class > 0;
    const int max_total = min(p.max_total > 0 ? min(p.max_total, p.max_out) : p.max_out, kLazyMaxKept);
    if (caps)
        for (int c = tid; c < p.num_classes; c += kLazyThreads) class_count[c] = 0;
    if (tid == 0) s_kept = 0, s_stop = (max_total <= 0);
    __syncthreads();
    for (int w = 0; w < nblk && !s_stop; ++w) {
        const int first = w << 6, nb = min(64, m - first);
        const int kept_before = s_kept;
        if (tid < 64) {
            diag[tid] = 0ull;
            if (tid < nb) bbox[tid] = sorted_boxes[base + first + tid], bcls[tid] = sorted_cls[base + first + tid];
        }
        if (tid == 0) s_dead = 0ull;
        __syncthreads();
        {  // candidates of this block against everything kept so far: thread (c, part) walks a quarter of the kept list
            const int c = tid >> 2, part = tid & 3;
            if (c < nb) {
                const float4 cb = bbox[c];
                const int cc = bcls[c];
                bool dead = false;
                for (int q = part; q < kept_before && !dead; q += 4) dead = suppresses(p, kbox[q], kcls[q], cb, cc);
                if (dead) atomicOr(&s_dead, 1ull << c);
            }
        }
        __syncthreads();
        const unsigned long long alive = ~s_dead;
        {  // the block's own tile: thread (r, q) tests row r against columns 16 q .. 16 q + 15 that come after it
            const int r = tid >> 2, q = tid & 3;
            if (r < nb && ((alive >> r) & 1ull)) {
                const float4 a = bbox[r];
                const int ac = bcls[r];
                unsigned long long bits = 0ull;
                for (int c = max(16 * q, r + 1); c < min(16 * q + 16, nb); ++c)
                    if (((alive >> c) & 1ull) && suppresses(p, a, ac, bbox[c], bcls[c])) bits |= 1ull << c;
                if (bits) atomicOr(&diag[r], bits);
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long rem = ~alive, km = 0ull;
            int kept = kept_before;
            for (int r = 0; r < nb; ++r) {
                if ((rem >> r) & 1ull) continue;
                if (caps) {
                    const int c = bcls[r];
                    if (c >= 0 && c < p.num_classes) {
                        if (class_count[c] >= p.max_per_class) continue;  // never selected, suppresses nobody
                        class_count[c] += 1;
                    }
                }
                kbox[kept] = bbox[r], kcls[kept] = bcls[r];
                km |= 1ull << r;
                rem |= diag[r];
                if (++kept >= max_total) {
                    s_stop = 1;
                    break;
                }
            }
            s_km = km, s_kept = kept;
        }
        __syncthreads();
        const unsigned long long km = s_km;
        if (tid < nb && ((km >> tid) & 1ull))
            keep[static_cast<long long>(b) * p.max_out + kept_before + __popcll(km & ((1ull << tid) - 1ull))] = order[base + first + tid];
        __syncthreads();
    }
    if (tid == 0) n_keep[b] = s_kept;
}

}  // namespace dh

using namespace dh;

extern "C" int dh_nms(dh_handle_t h, const float* dets, const int32_t* n_valid, int batch, int n_max, int row_floats,
                      int mode, float iou_thr, float min_score, int score_inclusive, int num_classes, int max_per_class,
                      int max_total, int32_t* keep, int max_out, int32_t* n_keep, void* stream) {
    DH_CHECK_ARG(h && dets && keep && n_keep, "dh_nms: NULL argument");
    DH_CHECK_ARG(batch >= 0 && n_max >= 0 && max_out >= 1, "dh_nms: bad sizes");
    DH_CHECK_ARG(mode == DH_NMS_AGNOSTIC || mode == DH_NMS_PER_CLASS, "dh_nms: mode %d", mode);
    DH_CHECK_ARG(row_floats >= (mode == DH_NMS_PER_CLASS ? 6 : 5), "dh_nms: row_floats %d too small", row_floats);
    if (n_max > kNmsMaxN)
        return set_error(DH_ERR_CAPACITY, "dh_nms: %d candidates per image > %d; apply a pre-NMS top-k first", n_max, kNmsMaxN);
    DH_CHECK_ARG(mode != DH_NMS_PER_CLASS || max_per_class <= 0 || (num_classes >= 1 && num_classes <= 8192),
                 "dh_nms: per-class caps need 1..8192 classes");
    if (batch == 0) return DH_OK;
    DeviceGuard guard(h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_max == 0) {
        DH_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int32_t) * batch, st));
        return DH_OK;
    }
    NmsParams p;
    p.n_max = n_max, p.row_floats = row_floats, p.per_class = (mode == DH_NMS_PER_CLASS), p.inclusive = score_inclusive;
    p.iou_thr = iou_thr, p.min_score = min_score, p.max_per_class = max_per_class, p.max_total = max_total;
    p.num_classes = num_classes, p.max_out = max_out;
    int n_pow2 = 64;
    while (n_pow2 < n_max) n_pow2 <<= 1;
    const int words = (((n_max + 63) / 64) + 1) & ~1;  // even: every mask row is 16-byte aligned for the TMA bulk copies
    // scratch: sorted boxes (16 B), classes, order, counts, mask
    const size_t per_img = static_cast<size_t>(n_max);
    size_t off_cls = static_cast<size_t>(batch) * per_img * 16;
    size_t off_ord = off_cls + static_cast<size_t>(batch) * per_img * 4;
    size_t off_cnt = off_ord + static_cast<size_t>(batch) * per_img * 4;
    size_t off_flag = off_cnt + static_cast<size_t>(batch) * 4;
    size_t off_slab = (off_flag + static_cast<size_t>(batch) * 4 + 255) & ~size_t(255);
    size_t off_mask = (off_slab + static_cast<size_t>(batch) * per_img * 8 + 255) & ~size_t(255);
    // a small output cap ends the sweep after a few blocks: suppression is evaluated lazily, one CTA per image, and the
    // words^2 mask matrix (3 MB per image at 5 000 candidates) is never formed -- nor allocated
    const int eff_total = max_total > 0 ? (max_total < max_out ? max_total : max_out) : max_out;
    const bool lazy = eff_total <= kLazyMaxKept && (h->nms_kernel == 1 || (h->nms_kernel == 0 && eff_total <= 512));
    size_t total = lazy ? off_mask : off_mask + static_cast<size_t>(batch) * words * words * 64 * 8;
    char* sc = static_cast<char*>(scratch(h, total));
    if (!sc) return DH_ERR_CUDA;
    float4* sboxes = reinterpret_cast<float4*>(sc);
    int* scls = reinterpret_cast<int*>(sc + off_cls);
    int* order = reinterpret_cast<int*>(sc + off_ord);
    int* ncand = reinterpret_cast<int*>(sc + off_cnt);
    uint2* slabs = (lazy || !h->nms_filter) ? nullptr : reinterpret_cast<uint2*>(sc + off_slab);
    int* filter_ok = reinterpret_cast<int*>(sc + off_flag);
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(sc + off_mask);
    DH_ONCE_PER_DEVICE(h) {
        DH_CUDA(cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNmsMaxN * 8));
    }
    nms_sort_kernel<<<batch, kSortThreads, static_cast<size_t>(n_pow2) * 8, st>>>(dets, n_valid, p, n_pow2, h->nms_sort == 1 ? 1 : 0, sboxes, scls, order,
                                                                                    ncand, slabs, slabs ? filter_ok : nullptr);
    DH_CUDA(cudaGetLastError());
    const size_t cap_smem = (p.per_class && max_per_class > 0) ? static_cast<size_t>(num_classes) * 4 : 0;
    if (lazy) {
        nms_lazy_kernel<<<batch, kLazyThreads, cap_smem, st>>>(sboxes, scls, order, ncand, p, keep, n_keep);
        DH_CUDA(cudaGetLastError());
        h->launches += 2;
        return DH_OK;
    }
    dim3 grid(words, (words + kMaskRowBlocks - 1) / kMaskRowBlocks, batch);
    nms_mask_kernel<<<grid, 64 * kMaskRowBlocks, 0, st>>>(sboxes, scls, slabs, filter_ok, ncand, p, words, h->nms_filter > 1 ? h->nms_filter : 64,
                                                          mask);
    DH_CUDA(cudaGetLastError());
    const int staged = words <= kSweepStageWords ? 1 : 0;
    const size_t stage_bytes = staged ? static_cast<size_t>(kSweepDepth) * 64 * words * 8 : 0;  // [depth][words][64] words
    const size_t sweep_smem = stage_bytes + ((p.per_class && max_per_class > 0) ? static_cast<size_t>(num_classes) * 4 : 0);
    DH_ONCE_PER_DEVICE(h) {
        DH_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    nms_sweep_kernel<<<batch, kSweepThreads, sweep_smem, st>>>(mask, scls, order, ncand, p, words, staged, h->nms_chain, keep, n_keep);
    DH_CUDA(cudaGetLastError());
    h->launches += 3;
    return DH_OK;
}
