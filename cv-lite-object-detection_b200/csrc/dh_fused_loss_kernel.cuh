// Fused encode + loss, "stream + correct" formulation (the C5 step: format_data -> model_loss of one
// training step, FCOS/train_fcos.py:142-154, RetinaNet/train_retinanet_coco.py:198-208, with the targets
// never written anywhere).
//
// More than 99 % of the rows of a dense-head target map are all-zero, and for an all-zero row the loss needs
// no label at all: sum_c (1-a) * sigmoid(x_c)^g * softplus(x_c) over the class logits.  So the kernel splits
// the work of a 32-row warp tile in two:
//   stream   every element of the tile is read once from HBM (128-bit loads, U independent loads in flight
//            per lane) and accumulated as if its label were zero -- a tight loop with no shared-memory
//            traffic, no row bookkeeping and no block barrier; two of every four reciprocals run as a
//            quadratic seed + 2 Newton steps on the FMA pipe so that the MUFU pipe (ex2, lg2, rcp) is not
//            the limiter (tools/focal_stream_probe.cu: 5.65 TB/s against 5.15 TB/s with 3 MUFU per element);
//   correct  the encoder policy matches the tile's rows against the GT boxes that can touch it (per-warp
//            candidate list, ballot-compacted, no block barrier).  A matched row is kept in registers as
//            <= 5 regression values + a class bitmask (CompactSink); its owner lane re-reads that one row
//            (L2-resident) and adds  term(y, x) - term(0, x)  for the channels whose label is not zero, plus
//            the box-regression loss.  Rows are matched by the same policy code as the encoders, so the
//            target semantics cannot drift from dh_*_encode.
// Warps of a CTA never synchronise inside a chunk; partial sums stay in registers for the whole chunk and the
// per-chunk partials are reduced in a fixed order by loss_finalize_images (deterministic results).
#pragma once
#include "dh_loss_kernel.cuh"

namespace dh {

__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_fast(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 1/w for w in [1, 2] on the FMA pipe: minimax quadratic seed (rel. error 1.0e-2) + two Newton steps (-> 1e-8)
__device__ __forceinline__ float rcp_unit_newton(float w) {
    float r = fmaf(fmaf(0.32323232f, w, -1.45454545f), w, 2.12121212f);
    float t = fmaf(-w, r, 1.0f);
    r = fmaf(r, t, r);
    t = fmaf(-w, r, 1.0f);
    return fmaf(r, t, r);
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// label == 0 element of the classification loss, accumulated in log2 units.  kCls selects the loss:
//   0 focal, any gamma   acc += sigmoid(x)^gamma * softplus(x) / ln 2      (the caller scales by (1 - alpha) * ln 2)
//   1 focal, gamma == 2  the same with the power as one multiply
//   2 sigmoid BCE        acc += softplus(x) / ln 2                          (the caller scales by ln 2)
// With kGrad the derivative is returned too, in natural units before the caller's scale ((1 - alpha) * w_cls for
// focal: d/dx = s^g * (g * (1 - s) * softplus(x) + s);  w_cls for BCE: d/dx = s).
template <int kCls, bool kNewton, bool kGrad>
__device__ __forceinline__ float stream_term(float x, float gamma, float& acc) {
    const float u = x * kLog2e;
    const float e = ex2_fast(-fabsf(u));  // exp(-|x|)
    const float w = 1.0f + e;
    const float lg = lg2_fast(w);                    // log2(1 + exp(-|x|))
    const float sp2 = lg + fmaxf(u, 0.f);            // softplus(x) / ln 2
    if (kCls == 2 && !kGrad) {
        acc += sp2;
        return 0.f;
    }
    const float inv = kNewton ? rcp_unit_newton(w) : rcp_fast(w);
    const float s = (x < 0.f ? e : 1.0f) * inv;      // sigmoid(x)
    if (kCls == 2) {
        acc += sp2;
        return s;
    }
    const float pw = kCls == 1 ? s * s : ex2_fast(gamma * lg2_fast(s));
    acc = fmaf(pw, sp2, acc);
    if (!kGrad) return 0.f;
    const float om = (x < 0.f ? 1.0f : e) * inv;     // 1 - sigmoid(x)
    return pw * fmaf((kCls == 1 ? 2.0f : gamma) * kLn2 * om, sp2, s);
}
// ---- packed fp32 (sm_100a FFMA2 / FMUL2 / FADD2): two elements per FMA-pipe instruction --------------------------
// The streaming pass is bound by instruction issue, not by HBM: with the reciprocal on the FMA pipe (2 MUFU per
// element) and the FMA-pipe work packed two-wide, the loop of tools/focal_stream_probe.cu moves from 5.6 to 6.3 TB/s.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// two label == 0 focal elements with gamma == 2: acc += sigmoid(x)^2 * softplus(x) / ln 2 (both lanes of `acc`); with
// kGrad the derivatives s^2 * (2 (1 - s) softplus(x) + s) come back in (d0, d1), before the caller's scale
template <bool kGrad>
__device__ __forceinline__ void stream_pair_g2(float x0, float x1, f32x2& acc, float& d0, float& d1) {
    const f32x2 one = pack2(1.0f, 1.0f);
    float u0, u1;
    unpack2(mul2(pack2(x0, x1), pack2(kLog2e, kLog2e)), u0, u1);
    const float e0 = ex2_fast(-fabsf(u0)), e1 = ex2_fast(-fabsf(u1));  // exp(-|x|)
    const f32x2 w = add2(pack2(e0, e1), one);
    float w0, w1;
    unpack2(w, w0, w1);
    // 1 / w for w in [1, 2]: quadratic seed + two Newton steps, all on the FMA pipe
    const f32x2 nw = mul2(w, pack2(-1.0f, -1.0f));
    f32x2 r = fma2(fma2(pack2(0.32323232f, 0.32323232f), w, pack2(-1.45454545f, -1.45454545f)), w, pack2(2.12121212f, 2.12121212f));
    f32x2 t = fma2(nw, r, one);
    r = fma2(r, t, r);
    t = fma2(nw, r, one);
    const f32x2 inv = fma2(r, t, r);
    const f32x2 sp2 = add2(pack2(lg2_fast(w0), lg2_fast(w1)), pack2(fmaxf(u0, 0.f), fmaxf(u1, 0.f)));  // softplus(x) / ln 2
    const f32x2 s = mul2(pack2(x0 < 0.f ? e0 : 1.0f, x1 < 0.f ? e1 : 1.0f), inv);                      // sigmoid(x)
    const f32x2 pw = mul2(s, s);
    acc = fma2(pw, sp2, acc);
    if (kGrad) {
        const f32x2 om = mul2(pack2(x0 < 0.f ? 1.0f : e0, x1 < 0.f ? 1.0f : e1), inv);  // 1 - sigmoid(x)
        unpack2(mul2(pw, fma2(mul2(om, pack2(2.0f * kLn2, 2.0f * kLn2)), sp2, s)), d0, d1);
    }
}

// The same two elements when both logits are <= -0.7, i.e. e = exp(x) <= 0.4966 -- all but a few of the label == 0
// logits of a detector (prior -4.6): sigmoid(x)^2 softplus(x) / ln 2 = e^3 ln(1 + e) / (e (1 + e)^2 ln 2) = e^3 P(e) with a
// degree-7 polynomial (least-squares fit of the relative error on Chebyshev nodes of [0, 0.5]; max relative error
// 6.4e-7 in float32 Horner form).  1 MUFU + 5.5 packed FMA-pipe instructions per element instead of 2 MUFU + ~12.
constexpr float kSmallLogit = -0.7f;
__device__ __forceinline__ void stream_pair_g2_small(float x0, float x1, f32x2& acc) {
    float u0, u1;
    unpack2(mul2(pack2(x0, x1), pack2(kLog2e, kLog2e)), u0, u1);
    const f32x2 e = pack2(ex2_fast(u0), ex2_fast(u1));  // exp(x)
    f32x2 p = fma2(pack2(-3.1173593997955322f, -3.1173593997955322f), e, pack2(8.865461349487305f, 8.865461349487305f));
    p = fma2(p, e, pack2(-12.291608810424805f, -12.291608810424805f));
    p = fma2(p, e, pack2(11.74155330657959f, 11.74155330657959f));
    p = fma2(p, e, pack2(-9.157456398010254f, -9.157456398010254f));
    p = fma2(p, e, pack2(6.245364189147949f, 6.245364189147949f));
    p = fma2(p, e, pack2(-3.606579303741455f, -3.606579303741455f));
    p = fma2(p, e, pack2(1.4426943063735962f, 1.4426943063735962f));
    acc = fma2(mul2(mul2(e, e), e), p, acc);
}
// ... and its derivative, in natural units before the caller's scale: d/dx [sigmoid^2 softplus] = e^3 Q(e), a second
// degree-7 fit (max relative error 2.0e-6 in float32 Horner form over [0, 0.5])
__device__ __forceinline__ void stream_pair_g2_small_grad(float x0, float x1, f32x2& acc, float& d0, float& d1) {
    float u0, u1;
    unpack2(mul2(pack2(x0, x1), pack2(kLog2e, kLog2e)), u0, u1);
    const f32x2 e = pack2(ex2_fast(u0), ex2_fast(u1));  // exp(x)
    f32x2 p = fma2(pack2(-3.1173593997955322f, -3.1173593997955322f), e, pack2(8.865461349487305f, 8.865461349487305f));
    f32x2 q = fma2(pack2(-17.49591827392578f, -17.49591827392578f), e, pack2(48.32270050048828f, 48.32270050048828f));
    p = fma2(p, e, pack2(-12.291608810424805f, -12.291608810424805f));
    q = fma2(q, e, pack2(-63.491676330566406f, -63.491676330566406f));
    p = fma2(p, e, pack2(11.74155330657959f, 11.74155330657959f));
    q = fma2(q, e, pack2(55.418338775634766f, 55.418338775634766f));
    p = fma2(p, e, pack2(-9.157456398010254f, -9.157456398010254f));
    q = fma2(q, e, pack2(-37.818424224853516f, -37.818424224853516f));
    p = fma2(p, e, pack2(6.245364189147949f, 6.245364189147949f));
    q = fma2(q, e, pack2(21.622814178466797f, 21.622814178466797f));
    p = fma2(p, e, pack2(-3.606579303741455f, -3.606579303741455f));
    q = fma2(q, e, pack2(-9.998870849609375f, -9.998870849609375f));
    p = fma2(p, e, pack2(1.4426943063735962f, 1.4426943063735962f));
    q = fma2(q, e, pack2(2.999994993209839f, 2.999994993209839f));
    const f32x2 e3 = mul2(mul2(e, e), e);
    acc = fma2(e3, p, acc);
    unpack2(mul2(e3, q), d0, d1);
}
// smallest of the four values' bit patterns as unsigned integers: >= bits(kSmallLogit) exactly when every value is
// <= kSmallLogit (negative floats order by magnitude as unsigned integers; anything with the sign bit clear is smaller)
__device__ __forceinline__ unsigned min_bits4(const float4& x) {
    return min(min(__float_as_uint(x.x), __float_as_uint(x.y)), min(__float_as_uint(x.z), __float_as_uint(x.w)));
}

// smooth-L1(0, sigmoid(x)): the centerness term of an all-zero row (FCOS/fcos.py:483-486)
__device__ __forceinline__ float cen_l1_zero(float x, float delta) {
    const float s = sigmoid_f(x);
    return s < delta ? 0.5f * s * s : s;
}

constexpr int kTraceSlots = 12;  // dh_set_trace: longs per CTA
constexpr int kMaxChunkTiles = 18;  // tiles (of 256 rows) per scheduler chunk at most (launch_fused_g: 16 in the tiered plans, up to 18
                                    // when equal chunks of that size give every CTA exactly one -- 32 COCO images: 576 chunks of 18 tiles)
constexpr int kMaxSegments = kMaxChunkTiles;  // ... so a chunk meets at most as many maps
// A tile (256 rows) is streamed as 2 spans; the chunk's last tiles as 4, 4, 8 and up to 32 (see span_plan)
constexpr int kMaxSpans = (kMaxChunkTiles - 4) * 2 + 4 + 4 + 8 + 32;
struct FusedSmemLayout {
    int rec_off, raw_off, cand_off, seg_off, run_off, misc_off, args_off, total;
};
// Candidate list of one segment (the run of a chunk's tiles inside one map), built once per chunk by one warp and read by
// all eight: the boxes that can match the map at all, in ascending GT order, with the cell rectangle each can touch.
constexpr int kRunCap = 8;      // boxes of one row a lane keeps privately (more: the segment's whole list is matched)
constexpr int kMaxPairs = 1024;  // (row, box) candidates one chunk can hold before it falls back to visiting every row
struct SegLists {
    int* nmap;                // [kMaxSegments] boxes that can match the segment's map (zero between chunks)
    int* npairs;              // [1] pairs appended (zero between chunks); > kMaxPairs: overflow
    unsigned short* box;      // [kMaxSegments][box_cap] those boxes, in no particular order
    uint32_t* pairs;          // [kMaxPairs] chunk row << 16 | box, in no particular order
    unsigned short* next;     // [kMaxPairs] the pair appended before this one for the same row (kNoPair: none)
    uint32_t* head;           // [kMaxChunkTiles * 256 / 2] per chunk row, 16 bits each: the last pair appended for it (kNoPair between
                              // chunks); two rows share a word (head_exchange), which keeps the kernel under the 48 KB carve-out step
    uint32_t* rowbits;        // [kMaxChunkTiles * 8] one bit per chunk row: has a pair (zero between chunks)
    unsigned short* rowlist;  // [kMaxPairs] the chunk rows that have pairs, ascending (scan_rows)
    int* tile_tab;            // [kMaxChunkTiles][4] per tile of the chunk: map, first row, segment, rows
    float* span_sum;          // [kMaxSpans][2] the label-free sums {class, centerness} of each span of the chunk
    int2* span_tab;           // [kMaxSpans] the chunk's spans: {tile of the chunk, first unit | units << 16} (build_tile_tab)
    int* span_ctr;            // [4] [0] next span to hand out (zero between chunks), [1] spans of the chunk
};
constexpr uint32_t kNoPair = 0xffffu;
// the row's 16-bit slot of the head table: swap in `val`, return what was there (shared-memory atomics are 32 bits wide:
// a compare-and-swap on the word the row shares with its neighbour; rows of a chunk rarely collide)
__device__ __forceinline__ uint32_t head_exchange(uint32_t* head, int row, uint32_t val) {
    uint32_t* w = head + (row >> 1);
    const int sh = (row & 1) * 16;
    uint32_t old = *w, assumed;
    do {
        assumed = old;
        old = atomicCAS(w, assumed, (assumed & ~(0xffffu << sh)) | (val << sh));
    } while (old != assumed);
    return (old >> sh) & 0xffffu;
}
__device__ __forceinline__ uint32_t head_take(uint32_t* head, int row) { return head_exchange(head, row, kNoPair); }
template <class P>
__host__ __device__ inline FusedSmemLayout fused_smem_layout(int box_cap) {
    FusedSmemLayout l;
    l.rec_off = 0;
    l.raw_off = (static_cast<int>(sizeof(typename P::Rec)) * box_cap + 127) & ~127;
    l.cand_off = l.raw_off + ((box_cap * 20 + 127) & ~127);
    l.seg_off = l.cand_off;
    l.run_off = l.seg_off + 128 + kMaxSegments * box_cap * 2 + kMaxPairs * 8 + kMaxChunkTiles * (DH_THREADS * 2 + 32 + 16) + kMaxSpans * 16 + 16;  // SegLists
    l.misc_off = l.run_off + (DH_THREADS / 2) * kRunCap * 2;  // resolve_pass (run by the four helper warps): the boxes of the row a lane resolves
    l.args_off = l.misc_off + 512;
    l.total = l.args_off + ((static_cast<int>(sizeof(LossArgs<P>)) + 127) & ~127);
    // Mind the total: it decides the shared-memory carve-out of the SM, and what the carve-out leaves is the L1 that the
    // loads in flight land in.  Four CTAs of <= 40 KB (+ 1 KB each that the system keeps) fit the 164 KB carve-out (92 KB of
    // L1); up to 48 KB the 196 KB one (60 KB of L1: 2-5 % slower from 48 COCO images up -- 565 us against 537 at 128, 1 038
    // against 1 020 at 256); at 53 KB per CTA the SM takes the 228 KB carve-out, 28 KB of L1 remain, and the streaming pass
    // loses a fifth of its rate (1 218 us at 256 images).  RetinaNet-COCO (box_cap 128) comes to 39 984 bytes: the 16-bit
    // head table, the run lists sized for the four helper warps and kMaxChunkTiles = 18 are chosen against that step.
    return l;
}

struct StreamAcc {
    float c0, c1, c2, c3;  // class focal, log2 units (4 chains for ILP)
    float cen;             // centerness of all-zero rows, natural units
    f32x2 p0, p1;          // packed-math chains of the gamma == 2 vector path (two lanes each), log2 units
};

// ---- the label-free pass over one warp tile: rows [0, nrows) x ch floats starting at `p` --------------------
// Lane l takes items l, l + 32, l + 64, ... (an item = one 128-bit load in the vector pass, one float in the
// scalar pass); `c` tracks the item's position inside its row incrementally (no division in the loop).
template <int kCls, bool kGrad>
__device__ __forceinline__ void stream_vec_item(const float4& x, bool is_reg, float gamma, float gscale, StreamAcc& a, float4* g) {
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!is_reg) {  // float4 0 of a row = the 4 regression channels
        if (kCls == 1) {
            stream_pair_g2<kGrad>(x.x, x.y, a.p0, d.x, d.y);
            stream_pair_g2<kGrad>(x.z, x.w, a.p1, d.z, d.w);
        } else {
            d.x = stream_term<kCls, true, kGrad>(x.x, gamma, a.c0);
            d.y = stream_term<kCls, true, kGrad>(x.y, gamma, a.c1);
            d.z = stream_term<kCls, true, kGrad>(x.z, gamma, a.c2);
            d.w = stream_term<kCls, false, kGrad>(x.w, gamma, a.c3);
        }
    }
    if (kGrad) __stcs(g, make_float4(d.x * gscale, d.y * gscale, d.z * gscale, d.w * gscale));
}
// `c4` = position of this lane's next item inside its row (in float4s), advanced by 32 mod vpr per item: no division
// in the loop, and correct for every row length (a lane may meet several regression float4s per tile, or none)
// the elements of a float4 whose channel is below n_skip (box regression, centerness) become -inf: they add exactly 0 to
// the label-free class sum and get a zero gradient.  c0 = channel of the float4's first element.
__device__ __forceinline__ void mask_skip(float4& x, int c0, int ch, int n_skip) {
    int c = c0;
    if (c < n_skip) x.x = -INFINITY;
    c = c + 1 == ch ? 0 : c + 1;
    if (c < n_skip) x.y = -INFINITY;
    c = c + 1 == ch ? 0 : c + 1;
    if (c < n_skip) x.z = -INFINITY;
    c = c + 1 == ch ? 0 : c + 1;
    if (c < n_skip) x.w = -INFINITY;
}
// kAny = false: rows are whole float4s (ch % 4 == 0) and start with exactly the 4 regression channels -- `vpr` float4s per
// row, c4 = the item's position in its row.  kAny = true: any row length and channel layout over a 16-byte-aligned flat
// range -- `vpr` is then the row length in floats, c4 the channel of the item's first element, `n_skip` the channels
// that are not class logits (the centerness channel is added by a separate pass over the rows, stream_cen).
template <int kCls, bool kGrad, int U, bool kAny = false>
__device__ __forceinline__ void stream_vec(const float* __restrict__ p, float* __restrict__ gout, int n_items, int vpr, int c4, int step,
                                           int lane, float gamma, float gscale, StreamAcc& a, int n_skip = 4) {
    const float4* __restrict__ ptr = reinterpret_cast<const float4*>(p) + lane;
    float4* __restrict__ gptr = reinterpret_cast<float4*>(gout) + lane;
    const int n_mine = (n_items - lane + 31) >> 5;  // items of this lane (may be <= 0 in a ragged tile)
    int k = 0;
    if constexpr (kCls == 1) {
        // gamma == 2: when every class logit of the warp's batch is <= kSmallLogit (one vote per 128 U elements; 96 % of the
        // batches at logits ~ N(-4.6, 1)) the batch takes the polynomial form of the term (and of its derivative)
        const unsigned small_bits = __float_as_uint(kSmallLogit);
        const int n_all = n_items >> 5;  // items EVERY lane has: the vote below needs a warp-uniform trip count
#pragma unroll 1
        for (; k + U <= n_all; k += U, ptr += 32 * U, gptr += 32 * U) {
            float4 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = __ldcs(ptr + 32 * u);
            unsigned lowest = 0xFFFFFFFFu;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                // float4 0 of a row = the 4 regression channels: as -inf they add exactly 0 in either form (e = 0), which
                // is cheaper than branching around them
                if constexpr (kAny) mask_skip(x[u], c4, vpr, n_skip);
                else if (c4 == 0) x[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                lowest = min(lowest, min_bits4(x[u]));
                c4 += step;
                if (c4 >= vpr) c4 -= vpr;
            }
            if (__all_sync(0xffffffffu, lowest >= small_bits)) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if constexpr (kGrad) {
                        float4 d;
                        stream_pair_g2_small_grad(x[u].x, x[u].y, a.p0, d.x, d.y);
                        stream_pair_g2_small_grad(x[u].z, x[u].w, a.p1, d.z, d.w);
                        __stcs(gptr + 32 * u, make_float4(d.x * gscale, d.y * gscale, d.z * gscale, d.w * gscale));
                    } else {
                        stream_pair_g2_small(x[u].x, x[u].y, a.p0);
                        stream_pair_g2_small(x[u].z, x[u].w, a.p1);
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) stream_vec_item<kCls, kGrad>(x[u], false, gamma, gscale, a, gptr + 32 * u);
            }
        }
    } else {
#pragma unroll 1
        for (; k + U <= n_mine; k += U, ptr += 32 * U, gptr += 32 * U) {  // U independent loads in flight, no predicates
            float4 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = __ldcs(ptr + 32 * u);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if constexpr (kAny) {
                    mask_skip(x[u], c4, vpr, n_skip);
                    stream_vec_item<kCls, kGrad>(x[u], false, gamma, gscale, a, gptr + 32 * u);
                } else {
                    stream_vec_item<kCls, kGrad>(x[u], c4 == 0, gamma, gscale, a, gptr + 32 * u);
                }
                c4 += step;
                if (c4 >= vpr) c4 -= vpr;
            }
        }
    }
#pragma unroll 1
    for (; k < n_mine; ++k, ptr += 32, gptr += 32) {
        float4 x = __ldcs(ptr);
        if constexpr (kAny) {
            mask_skip(x, c4, vpr, n_skip);
            stream_vec_item<kCls, kGrad>(x, false, gamma, gscale, a, gptr);
        } else {
            stream_vec_item<kCls, kGrad>(x, c4 == 0, gamma, gscale, a, gptr);
        }
        c4 += step;
        if (c4 >= vpr) c4 -= vpr;
    }
}
// kAny companion: the centerness channel of rows [0, nrows) as if its label were zero (one element per row, lane <-> row;
// the sectors were just streamed by this warp), and the up to three floats a flat range leaves behind its last float4.
template <int kCls, bool kGrad>
__device__ __forceinline__ void stream_cen_and_tail(const float* __restrict__ p, float* __restrict__ gout, int nrows, int ch, int lane,
                                                    const LossSpec& sp, StreamAcc& a) {
    if (kGrad) __syncwarp();  // the float4 stores above wrote a zero where the centerness gradient goes now
    if (sp.cen_mode == 1 || sp.cen_mode == 2) {
#pragma unroll 1
        for (int r = lane; r < nrows; r += 32) {
            const float x = p[static_cast<long long>(r) * ch + sp.reg_ch];
            if (sp.cen_mode == 1) {
                a.cen += cen_l1_zero(x, sp.delta);
                if (kGrad) gout[static_cast<long long>(r) * ch + sp.reg_ch] = sp.w_cen * cen_l1_grad(0.f, x, sp.delta);
            } else {
                a.cen += focal_term(0.f, x, sp.alpha, sp.gamma);
                if (kGrad) gout[static_cast<long long>(r) * ch + sp.reg_ch] = sp.w_cen * focal_grad(0.f, x, sp.alpha, sp.gamma);
            }
        }
    }
    const int n = nrows * ch, done = n & ~3;
    if (lane < n - done) {
        const int e = done + lane, c = e % ch;
        const int n_skip = sp.reg_ch + (sp.cen_mode != 0 ? 1 : 0);
        float d = 0.f;
        if (c >= n_skip) d = stream_term<kCls, false, kGrad>(p[e], sp.gamma, a.c3) * ((kCls == 2 ? 1.0f : 1.0f - sp.alpha) * sp.w_cls);
        if (kGrad && !(c == sp.reg_ch && (sp.cen_mode == 1 || sp.cen_mode == 2))) gout[e] = d;  // (the loop above wrote a centerness element)
    }
}

template <int kCls, bool kNewton, bool kGrad>
__device__ __forceinline__ void stream_scalar_item(float x, int c, const LossSpec& sp, float& cls_acc, StreamAcc& a, float* g) {
    float d = 0.f;
    if (c >= sp.reg_ch) {
        if (c == sp.reg_ch && sp.cen_mode != 0) {
            if (sp.cen_mode == 1) {
                a.cen += cen_l1_zero(x, sp.delta);
                if (kGrad) d = sp.w_cen * cen_l1_grad(0.f, x, sp.delta);
            } else if (sp.cen_mode == 2) {
                a.cen += focal_term(0.f, x, sp.alpha, sp.gamma);
                if (kGrad) d = sp.w_cen * focal_grad(0.f, x, sp.alpha, sp.gamma);
            }
        } else {
            d = stream_term<kCls, kNewton, kGrad>(x, sp.gamma, cls_acc) * ((kCls == 2 ? 1.0f : 1.0f - sp.alpha) * sp.w_cls);
        }
    }
    if (kGrad) __stcs(g, d);
}
template <int kCls, bool kGrad, int U>
__device__ __forceinline__ void stream_scalar(const float* __restrict__ p, float* __restrict__ gout, int nrows, int ch, int c, int step,
                                              int lane, const LossSpec& sp, StreamAcc& a) {
    const float* __restrict__ base = p + lane;
    float* __restrict__ gbase = gout + lane;
    const int n_mine = (nrows * ch - lane + 31) >> 5;
    int k = 0;
#pragma unroll 1
    for (; k + U <= n_mine; k += U) {
        float x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = __ldcs(base + 32 * (k + u));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u & 1) stream_scalar_item<kCls, true, kGrad>(x[u], c, sp, a.c1, a, gbase + 32 * (k + u));
            else stream_scalar_item<kCls, false, kGrad>(x[u], c, sp, a.c0, a, gbase + 32 * (k + u));
            c += step;
            if (c >= ch) c -= ch;
        }
    }
#pragma unroll 1
    for (; k < n_mine; ++k) {
        stream_scalar_item<kCls, false, kGrad>(__ldcs(base + 32 * k), c, sp, a.c2, a, gbase + 32 * k);
        c += step;
        if (c >= ch) c -= ch;
    }
}

// The owner lane of a matched row: box-regression loss + (term(y, x) - term(0, x)) for the labelled channels.
// Rare (well under 1 % of the rows), so it is kept out of line: its registers do not weigh on the streaming loop.
__device__ __noinline__ LossAcc correct_row(const LossSpec& sp, const float* __restrict__ prow, float* __restrict__ grow, CompactSink sink,
                                            float gy, float gx) {
    const int cls0 = sp.reg_ch + (sp.cen_mode != 0 ? 1 : 0);
    LossAcc acc = {0.f, 0.f, 0.f, 1};
    float x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = prow[k];
    if (sp.reg_mode == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc.reg += smooth_l1_term(sink.r[k], x[k], sp.delta);
        if (grow) {
#pragma unroll
            for (int k = 0; k < 4; ++k) grow[k] = sp.w_reg * smooth_l1_grad(sink.r[k], x[k], sp.delta);
        }
    } else {
        acc.reg += iou_loss_term(sink.r, x, gy, gx, sp.reg_mode);
        if (grow) {
            float g[4];
            iou_loss_grad(sink.r, x, gy, gx, g, sp.reg_mode);
#pragma unroll
            for (int k = 0; k < 4; ++k) grow[k] = sp.w_reg * g[k];
        }
    }
    if (sp.cen_mode == 1) {
        const float s = sigmoid_f(prow[4]);
        acc.cen += smooth_l1_term(sink.r[4], s, sp.delta) - smooth_l1_term(0.f, s, sp.delta);
        if (grow) grow[4] = sp.w_cen * cen_l1_grad(sink.r[4], prow[4], sp.delta);
    } else if (sp.cen_mode == 2) {
        const float xc = prow[4];
        acc.cen += focal_term(sink.r[4], xc, sp.alpha, sp.gamma) - focal_term(0.f, xc, sp.alpha, sp.gamma);
        if (grow) grow[4] = sp.w_cen * focal_grad(sink.r[4], xc, sp.alpha, sp.gamma);
    }
#pragma unroll 1
    for (int wd = 0; wd < kCompactClassWords; ++wd) {
        uint32_t m = sink.m[wd];
#pragma unroll 1
        while (m) {
            const int c = __ffs(m) - 1 + 32 * wd;
            m &= m - 1;
            const float xc = prow[cls0 + c];
            acc.cls += cls_term(sp, 1.0f, xc) - cls_term(sp, 0.f, xc);
            if (grow) grow[cls0 + c] = sp.w_cls * cls_grad(sp, 1.0f, xc);
        }
    }
    return acc;
}

// This warp's slice of tile `cur`: rows [r0, r0 + nrows) of map ti.m, or nrows <= 0 when the tile has fewer rows.
template <class P>
__device__ __forceinline__ const float* warp_tile(const LossArgs<P>& a, const TileCursor& cur, int img, int warp, TileInfo& ti,
                                                  float** grad = nullptr) {
    cursor_info(a.tt, cur, ti);
    ti.nrows = min(32, ti.nrows - 32 * warp);
    ti.r0 += 32 * warp;
    const MapDesc& md = a.tt.maps[ti.m];
    const long long off = static_cast<long long>(img) * md.image_stride + static_cast<long long>(ti.r0) * a.tt.ch;
    if (grad) *grad = a.grad_maps[ti.m] ? a.grad_maps[ti.m] + off : nullptr;
    return md.pred + off;
}

// ---- correct: the rows of the chunk that receive targets --------------------------------------------------------
// Nearly every row has no target at all and the SMs are busy issuing the streaming pass, so this part must cost
// instructions in proportion to the rows that DO get one -- and keep its lanes full.  Per chunk (at most 8 tiles of 256
// rows, walked as "segments" = runs of tiles inside one map):
//   pair   (before the streaming pass) the (segment, 32-box group) pairs are dealt to the warps; a warp finds the boxes
//          of its group that can match the segment's map at all (P::map_hit) and the cell rectangle each can touch
//          (P::cell_bounds), then walks every such rectangle with its 32 lanes, one cell per lane, and appends
//          (row, box) to the chunk's pair list where the box may paint the cell (P::pair_hit: a cheap superset test);
//          and chains the pair to the row's earlier ones (an atomicExch on the row's slot of a shared-memory table);
//          and sets the row's bit in the chunk's row map;
//   resolve  (after the streaming pass) one warp lists the marked rows in ascending order; then 32 rows per warp pass: a
//          lane walks its row's chain, runs the exact row matcher (P::match_row) over those boxes and adds the correction of the row
//          (correct_row); the matcher does not depend on the order of the boxes, so neither does the result.
// A chunk with more than kMaxPairs candidates (an IoU threshold near zero, say) visits every row instead (visit_dense).
__device__ __forceinline__ SegLists seg_lists(unsigned char* base, int box_cap) {
    SegLists L;
    L.nmap = reinterpret_cast<int*>(base);
    L.npairs = L.nmap + kMaxSegments;
    L.head = reinterpret_cast<uint32_t*>(base + 128);
    L.pairs = L.head + kMaxChunkTiles * DH_THREADS / 2;
    L.rowbits = L.pairs + kMaxPairs;
    L.tile_tab = reinterpret_cast<int*>(L.rowbits + kMaxChunkTiles * 8);
    L.span_sum = reinterpret_cast<float*>(L.tile_tab + kMaxChunkTiles * 4);
    L.span_tab = reinterpret_cast<int2*>(L.span_sum + kMaxSpans * 2);
    L.span_ctr = reinterpret_cast<int*>(L.span_tab + kMaxSpans);
    L.next = reinterpret_cast<unsigned short*>(L.span_ctr + 4);
    L.rowlist = L.next + kMaxPairs;
    L.box = L.rowlist + kMaxPairs;
    return L;
}

// Tile `tile` of image `img` and the segment of the chunk it belongs to.
template <class P>
__device__ __forceinline__ int locate_tile(const LossArgs<P>& a, int img, int t_begin, int tile, TileCursor& cur) {
    cursor_init(a.tt, static_cast<long long>(img) * a.tt.tiles_per_image + t_begin, cur);
    int seg = 0;
    for (int t = t_begin; t < tile; ++t) {
        const int m = cur.m;
        cursor_next(a.tt, cur);
        seg += cur.m != m;
    }
    return seg;
}

// tile t of the chunk: where it lies (every chunk; read by the streaming pass and by resolve_pass)
template <class P>
__device__ __forceinline__ void build_tile_tab(const LossArgs<P>& a, const SegLists& L, int img, int t_begin, int t_end) {
    if (static_cast<int>(threadIdx.x) < t_end - t_begin) {
        TileCursor cur;
        const int seg_t = locate_tile<P>(a, img, t_begin, t_begin + threadIdx.x, cur);
        const int r0 = cur.t * a.tt.rows_per_tile;
        L.tile_tab[threadIdx.x * 4 + 0] = cur.m, L.tile_tab[threadIdx.x * 4 + 1] = r0, L.tile_tab[threadIdx.x * 4 + 2] = seg_t;
        L.tile_tab[threadIdx.x * 4 + 3] = min(a.tt.rows_per_tile, a.tt.maps[cur.m].rows - r0);
    }
}

// The spans the warps take from the chunk's counter, coarse first and ever finer towards the chunk's end: a warp streams
// at an eighth of the CTA's share of the HBM rate, so when the counter runs dry the warps wait for the one that took the
// last span for about half a span's time -- 15 us per chunk with 128-row spans, which the few chunks per CTA of a small
// batch cannot hide.  Tile e (counted from the chunk's END) is cut into 2 spans for e >= 4 and into 2^min(fine, ...) spans
// nearer the end: 4 for e = 2, 3, 8 for e = 1, 32 for the last tile.  Measured (B200, 32 / 64 COCO images): fine = 2 -- the
// last four tiles in 64-row spans -- 178 / 302 us; every tile in 2 spans 185 / 308; down to single load batches (fine = 5)
// 185 / 309: a span's fixed cost (the counter, two warp reductions, an empty load pipeline) eats what the shorter wait
// gives.  A chunk of one to four tiles (the fine tiers that end a launch) is cut into at least 24 spans whatever `fine`
// says: with four 64-row spans to a tile half of the warps of a one-tile chunk had nothing to take.
// A span is measured in `unit`s: 128-bit items of the tile when
// the rows are whole float4s (then every span but a tile's last is a whole number of load batches of `batch` items: no
// ragged round trip), else rows (multiples of 32).  Thread t writes the spans of tile t; returns nothing -- span_ctr[1].
__device__ __forceinline__ int spans_of_tile(int e, int n_tiles, int units, int batch, int fine, int& per_span) {
    int shift = min(fine, e >= 4 ? 1 : (e >= 2 ? 2 : (e == 1 ? 3 : 5)));  // 2, 4, 8, 32 spans wanted
    while (shift < 5 && (n_tiles << shift) < 24) ++shift;                // a short chunk: three spans per warp all the same
    const unsigned share = (static_cast<unsigned>(units) + (1u << shift) - 1u) >> shift;
    per_span = static_cast<int>((share + batch - 1u) / static_cast<unsigned>(batch)) * batch;
    return per_span > 0 ? static_cast<int>((static_cast<unsigned>(units) + per_span - 1u) / static_cast<unsigned>(per_span)) : 0;
}
// (called by every thread of warp 0, between the chunk-end barriers: it is on every warp's critical path -- one thread per
// tile, two divisions each, a shuffle scan for the first span of each tile)
template <class P>
__device__ __forceinline__ void build_span_tab(const LossArgs<P>& a, const SegLists& L, int n_tiles, bool items, int batch) {
    const int t = threadIdx.x;
    if (t >= 32) return;
    int per = 0, mine = 0, units = 0;
    if (t < n_tiles) {
        const int rows = L.tile_tab[t * 4 + 3];
        units = items ? rows * (a.tt.ch >> 2) : rows;
        mine = spans_of_tile(n_tiles - 1 - t, n_tiles, units, items ? batch : 32, a.span_fine, per);
    }
    int first = mine;
#pragma unroll
    for (int d = 1; d < kMaxChunkTiles; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, first, d);
        if (t >= d) first += u;
    }
    first -= mine;
    for (int q = 0; q < mine; ++q) L.span_tab[first + q] = make_int2(t, (q * per) | (min(per, units - q * per) << 16));
    if (t == n_tiles - 1) L.span_ctr[1] = first + mine;
}

template <class P>
__device__ __forceinline__ void pair_pass(const LossArgs<P>& a, const typename P::Rec* recs, int n_boxes, const SegLists& L,
                                          int img, int t_begin, int t_end, int n_workers) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rpt = a.tt.rows_per_tile;
    const int n_groups = (n_boxes + 31) >> 5;
    TileCursor cur;
    cursor_init(a.tt, static_cast<long long>(img) * a.tt.tiles_per_image + t_begin, cur);
    int seg = 0, item = 0;
#pragma unroll 1
    for (int tile = t_begin; tile < t_end; ++seg) {
        const MapDesc& md = a.tt.maps[cur.m];
        const int nseg = min(md.n_tiles - cur.t, t_end - tile);
        const int row0 = cur.t * rpt;                                  // the segment's rows in the map: [row0, row1]
        const int row1 = min(row0 + nseg * rpt, md.rows) - 1;
        const int local0 = (tile - t_begin) * rpt - row0;              // map row -> row of the chunk
        const int cell0 = static_cast<int>(fdiv_u32(static_cast<uint32_t>(row0), md.div_sub));
        const int cell1 = static_cast<int>(fdiv_u32(static_cast<uint32_t>(row1), md.div_sub));
        const int seg_i0 = static_cast<int>(fdiv_u32(static_cast<uint32_t>(cell0), md.div_width));
        const int seg_i1 = static_cast<int>(fdiv_u32(static_cast<uint32_t>(cell1), md.div_width));
#pragma unroll 1
        for (int g = 0; g < n_groups; ++g, ++item) {
            if (item % n_workers != warp) continue;                    // (warp-uniform) this warp's items
            const int k = g * 32 + lane;
            int ilo = 0, ihi = -1, jlo = 0, jhi = -1;
            bool hit = k < n_boxes && P::map_hit(a.pp, recs[k], md.level, md.anchor) &&
                       P::cell_bounds(a.pp, recs[k], md, md.level, md.anchor, ilo, ihi, jlo, jhi);
            ilo = max(ilo, seg_i0), ihi = min(ihi, seg_i1);           // the map rows the segment holds
            hit = hit && ilo <= ihi;
            unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (bal) {
                int base = 0;
                if (lane == 0) base = atomicAdd(L.nmap + seg, __popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (hit) L.box[seg * a.box_cap + base + __popc(bal & ((1u << lane) - 1u))] = static_cast<unsigned short>(k);
            }
            // every rectangle in turn, one cell (its `sub` rows) per lane
            while (bal) {
                const int src = __ffs(bal) - 1;
                bal &= bal - 1;
                const int bk = g * 32 + src;
                const int bi0 = __shfl_sync(0xffffffffu, ilo, src), bi1 = __shfl_sync(0xffffffffu, ihi, src);
                const int bj0 = __shfl_sync(0xffffffffu, jlo, src), bj1 = __shfl_sync(0xffffffffu, jhi, src);
                const int nw = bj1 - bj0 + 1, n = (bi1 - bi0 + 1) * nw * md.sub;
                const float inv = 1.0f / static_cast<float>(nw * md.sub);
                const typename P::Rec& r = recs[bk];
#pragma unroll 1
                for (int c0 = 0; c0 < n; c0 += 32) {
                    const int c = c0 + lane;
                    bool ok = false;
                    int row = 0;
                    if (c < n) {
                        const int di = __float2int_rz((static_cast<float>(c) + 0.5f) * inv);  // c / (nw * sub), exact for c < 2^20
                        const int rem = c - di * nw * md.sub;
                        const int i = bi0 + di, j = bj0 + static_cast<int>(fdiv_u32(static_cast<uint32_t>(rem), md.div_sub));
                        row = (i * md.width + bj0) * md.sub + rem;
                        ok = row >= row0 && row <= row1 && P::pair_hit(a.pp, r, md, md.level, md.anchor, row, i, j);
                    }
                    const unsigned okb = __ballot_sync(0xffffffffu, ok);
                    if (okb) {
                        int at = 0;
                        if (lane == 0) at = atomicAdd(L.npairs, __popc(okb));
                        at = __shfl_sync(0xffffffffu, at, 0) + __popc(okb & ((1u << lane) - 1u));
                        if (ok && at < kMaxPairs) {  // append, and chain to the row's earlier pairs
                            L.pairs[at] = (static_cast<uint32_t>(local0 + row) << 16) | static_cast<uint32_t>(bk);
                            L.next[at] = static_cast<unsigned short>(head_exchange(L.head, local0 + row, static_cast<uint32_t>(at)));
                            atomicOr(L.rowbits + ((local0 + row) >> 5), 1u << ((local0 + row) & 31));
                        }
                    }
                }
            }
        }
        tile += nseg;
        cur.t += nseg;  // (cursor_next, nseg times)
        if (cur.t == md.n_tiles) {
            cur.t = 0;
            if (++cur.m == a.tt.n_maps) cur.m = 0, ++cur.b;
        }
    }
}

// The chunk rows that have pairs, in ascending order (one warp; the row map goes back to zero).  Returns their number.
__device__ __forceinline__ int scan_rows(const SegLists& L) {
    const int lane = threadIdx.x & 31;
    constexpr int kWords = kMaxChunkTiles * 8, kPer = (kWords + 31) / 32;  // words per lane (rounded UP: 144 words are 4.5 per lane)
    uint32_t w[kPer];
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
        const int i = lane * kPer + u;
        w[u] = i < kWords ? L.rowbits[i] : 0u;
        cnt += __popc(w[u]);
        if (i < kWords) L.rowbits[i] = 0u;
    }
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    int at = incl - cnt;
#pragma unroll
    for (int u = 0; u < kPer; ++u)
        for (uint32_t m = w[u]; m; m &= m - 1) L.rowlist[at++] = static_cast<unsigned short>((lane * kPer + u) * 32 + __ffs(m) - 1);
    return __shfl_sync(0xffffffffu, incl, 31);
}

// Rows are resolved in ascending order, 32 per warp pass (lane <-> row), so the order of the float sums does not depend on
// the order in which the pairs were appended: results are run-to-run deterministic.
template <class P>
__device__ __noinline__ LossAcc resolve_pass(const LossArgs<P>& a, const typename P::Rec* recs, const SegLists& L, int n_rows, int img,
                                             unsigned short* scratch, int n_workers) {
    LossAcc acc = {0.f, 0.f, 0.f, 0};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rpt = a.tt.rows_per_tile;
    unsigned short* mine = scratch + threadIdx.x * kRunCap;  // this lane's boxes (a row seldom has more than two or three)
#pragma unroll 1
    for (int p0 = warp * 32; p0 < n_rows; p0 += 32 * n_workers) {  // warp-uniform trip count
        const int p = p0 + lane;
        int pairs = 0;
        if (p < n_rows) {
            const int crow = L.rowlist[p];
            int len = 0;
            for (uint32_t q = head_take(L.head, crow); q != kNoPair; q = L.next[q], ++len)
                if (len < kRunCap) mine[len] = static_cast<unsigned short>(L.pairs[q] & 0xffffu);
            const int tix = crow / rpt;
            const int m = L.tile_tab[tix * 4 + 0], r0 = L.tile_tab[tix * 4 + 1], seg = L.tile_tab[tix * 4 + 2];
            const MapDesc& md = a.tt.maps[m];
            TileInfo ti;
            ti.b = img, ti.m = m, ti.r0 = r0, ti.nrows = L.tile_tab[tix * 4 + 3];
            ti.level = md.level, ti.anchor = md.anchor, ti.height = md.height, ti.width = md.width, ti.sub = md.sub;
            const unsigned short* boxes = mine;
            if (len > kRunCap) boxes = L.box + seg * a.box_cap, len = L.nmap[seg];  // (rare) every box that can match the map
            const int row = r0 + (crow - tix * rpt);
            const long long off = static_cast<long long>(img) * md.image_stride + static_cast<long long>(row) * a.tt.ch;
            CompactSink sink;
            sink.clear();
            pairs = P::match_row(a.pp, ti, md, row, sink, recs, boxes, len);
            if (pairs > 0) {
                const int cell = static_cast<int>(fdiv_u32(row, md.div_sub));
                const int i = static_cast<int>(fdiv_u32(cell, md.div_width));
                const LossAcc d = correct_row(a.spec, md.pred + off, a.grad_maps[m] ? a.grad_maps[m] + off : nullptr, sink,
                                              static_cast<float>(i), static_cast<float>(cell - i * md.width));
                acc.cls += d.cls, acc.reg += d.reg, acc.cen += d.cen, acc.npos += d.npos;
            }
        }
        TileInfo tb;
        tb.b = img;
        P::tile_epilogue(a.pp, tb, pairs);
    }
    return acc;
}

// Fallback for a chunk whose pair list overflowed: warp w visits every row of tile w with the segment's whole box list.
template <class P>
__device__ __noinline__ LossAcc visit_dense(const LossArgs<P>& a, const typename P::Rec* recs, const SegLists& L, int img, int t_begin,
                                            int t_end, int n_workers) {
    LossAcc acc = {0.f, 0.f, 0.f, 0};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int tile = t_begin + warp; tile < t_end; tile += n_workers) {
    TileCursor cur;
    const int seg = locate_tile<P>(a, img, t_begin, tile, cur);
    const MapDesc& md = a.tt.maps[cur.m];
    TileInfo ti;
    cursor_info(a.tt, cur, ti);
    const long long off = static_cast<long long>(img) * md.image_stride + static_cast<long long>(ti.r0) * a.tt.ch;
    const float* __restrict__ gp = md.pred + off;
    float* gg = a.grad_maps[ti.m] ? a.grad_maps[ti.m] + off : nullptr;
    const unsigned short* cand = L.box + seg * a.box_cap;
    const int ncand = L.nmap[seg];
    int pairs = 0;
#pragma unroll 1
    for (int r = lane; r < ti.nrows; r += 32) {
        CompactSink sink;
        sink.clear();
        const int row = ti.r0 + r;
        const int n = P::match_row(a.pp, ti, md, row, sink, recs, cand, ncand);
        if (n > 0) {
            pairs += n;
            const int cell = static_cast<int>(fdiv_u32(row, md.div_sub));
            const int i = static_cast<int>(fdiv_u32(cell, md.div_width));
            const LossAcc d = correct_row(a.spec, gp + static_cast<long long>(r) * a.tt.ch, gg ? gg + static_cast<long long>(r) * a.tt.ch : nullptr,
                                          sink, static_cast<float>(i), static_cast<float>(cell - i * md.width));
            acc.cls += d.cls, acc.reg += d.reg, acc.cen += d.cen, acc.npos += d.npos;
        }
    }
    P::tile_epilogue(a.pp, ti, pairs);
    }
    return acc;
}

// ---- stream: every element of the chunk as if its label were zero --------------------------------------------------
// The chunk's tiles are cut into spans of 64 or 128 rows which the warps TAKE from a shared-memory counter: a warp that
// was busy with something else (the helper warps pair boxes with rows and resolve the target rows first) simply takes
// fewer, and the warps reach the chunk-end barrier together.  A span's sums are reduced over the warp and stored per
// span; the chunk's sum is formed from them in span order afterwards, so it does not depend on which warp took which
// span: results stay run-to-run deterministic.
template <class P, int kCls, bool kGrad, int kMode /*0 scalar, 1 whole-float4 rows, 2 any layout over an aligned flat range*/>
__device__ __noinline__ void stream_spans(const LossArgs<P>& a, const SegLists& L, int img) {
    const int lane = threadIdx.x & 31;
    const int ch = a.tt.ch;
    const int n_spans = L.span_ctr[1];
    const int vpr = ch >> 2;
    const int step = kMode == 1 ? 32 % vpr : (kMode == 2 ? 128 % ch : 32 % ch);
    const int n_skip = a.spec.reg_ch + (a.spec.cen_mode != 0 ? 1 : 0);
    const float gamma = a.spec.gamma, gscale = (kCls == 2 ? 1.0f : 1.0f - a.spec.alpha) * a.spec.w_cls;
    const LossSpec sp = a.spec;
#pragma unroll 1
    for (;;) {
        int s = 0;
        if (lane == 0) s = atomicAdd(L.span_ctr, 1);
        s = __shfl_sync(0xffffffffu, s, 0);
        if (s >= n_spans) break;
        const int2 span = L.span_tab[s];
        const int tix = span.x, first = span.y & 0xffff, count = span.y >> 16;  // units: items (kMode 1) or rows
        const int m = L.tile_tab[tix * 4 + 0], r0 = L.tile_tab[tix * 4 + 1];
        StreamAcc sa = {0.f, 0.f, 0.f, 0.f, 0.f, 0ull, 0ull};
        if (count > 0) {
            const MapDesc& md = a.tt.maps[m];
            const long long e = static_cast<long long>(img) * md.image_stride + static_cast<long long>(r0) * ch +
                                static_cast<long long>(first) * (kMode == 1 ? 4 : ch);
            float* gg = (kGrad && a.grad_maps[m]) ? a.grad_maps[m] + e : nullptr;
            if constexpr (kMode == 1) {
                stream_vec<kCls, kGrad, kGrad ? 5 : 7>(md.pred + e, gg, count, vpr, (first + lane) % vpr, step, lane, gamma, gscale, sa);
            } else if constexpr (kMode == 2) {
                stream_vec<kCls, kGrad, kGrad ? 5 : 7, true>(md.pred + e, gg, (count * ch) >> 2, ch, (4 * lane) % ch, step, lane, gamma, gscale, sa, n_skip);
                stream_cen_and_tail<kCls, kGrad>(md.pred + e, gg, count, ch, lane, sp, sa);
            } else {
                stream_scalar<kCls, kGrad, 4>(md.pred + e, gg, count, ch, lane % ch, step, lane, sp, sa);
            }
        }
        float q0, q1, q2, q3;
        unpack2(sa.p0, q0, q1);
        unpack2(sa.p1, q2, q3);
        const float cls = warp_sum(((sa.c0 + sa.c1) + (sa.c2 + sa.c3)) + ((q0 + q1) + (q2 + q3)));
        const float cen = warp_sum(sa.cen);
        if (lane == 0) L.span_sum[2 * s] = cls, L.span_sum[2 * s + 1] = cen;
    }
}

__device__ __forceinline__ void helper_barrier(int n_threads) { asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory"); }

// The chunk -> (tier, image, first tile) map of the tiered scheduler (LossArgs::tiers).
template <class P>
__device__ __forceinline__ void chunk_span(const LossArgs<P>& a, long long chunk, int& img, int& sub, int& t_begin, int& t_end) {
    int tk = 0;
#pragma unroll
    for (int q = 1; q < kMaxChunkTiers; ++q)
        if (q < a.n_tiers && chunk >= a.tiers[q].chunk0) tk = q;
    const ChunkTier& T = a.tiers[tk];
    const int c2 = static_cast<int>(chunk - T.chunk0);
    const int di = c2 / T.cpi;
    img = T.image0 + di;
    sub = c2 - di * T.cpi;
    t_begin = sub * T.chunk_tiles;
    t_end = min(t_begin + T.chunk_tiles, a.tt.tiles_per_image);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reduction of the chunk partials without a second launch, by the last CTA of the launch:
//   reduce_images   chunk partials -> per_image rows.  L = 1..32 lanes share an image so that no lane has more than 16
//                   partials to add (a fine tail tier has 300 chunks per image, the coarse tier ~38); the (tier, group
//                   of images) items are dealt to the eight warps, so the tiers' L2 round trips overlap instead of
//                   following each other; float64, fixed lane / shuffle order -> deterministic.
//   finalize_total  per_image rows -> total, and, when the communicator is attached, the exchange of the total with
//                   the peer ranks (dh_comm.cuh).
template <class P>
__device__ __forceinline__ void reduce_images(const LossArgs<P>& a, double (&tot)[4]) {
    constexpr int kInFlight = 16;  // partials one lane loads per image: all issued before the first add (one L2 round trip)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4* part = reinterpret_cast<const float4*>(a.partials);
    int item = 0;  // (tier, group of 32 / L images) items are dealt to the warps: the tiers' round trips overlap
#pragma unroll 1
    for (int tk = 0; tk < a.n_tiers; ++tk) {
        const ChunkTier& T = a.tiers[tk];
        const int img_end = tk + 1 < a.n_tiers ? a.tiers[tk + 1].image0 : a.tt.batch;
        int L = 1;
        while (L < 32 && L * kInFlight < T.cpi) L <<= 1;
        const int G = 32 / L, sub = lane & (L - 1);
#pragma unroll 1
        for (int b0 = T.image0; b0 < img_end; b0 += G, ++item) {
            if ((item & (DH_THREADS / 32 - 1)) != warp) continue;  // warp-uniform
            const int b = b0 + lane / L;
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            if (b < img_end) {
                const float4* p = part + T.chunk0 + static_cast<long long>(b - T.image0) * T.cpi;
#pragma unroll 1
                for (int t0 = sub; t0 < T.cpi; t0 += L * kInFlight) {  // one trip unless an image has > 512 chunks
                    float4 v[kInFlight];
#pragma unroll
                    for (int u = 0; u < kInFlight; ++u)
                        v[u] = t0 + u * L < T.cpi ? __ldcg(p + t0 + u * L) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < kInFlight; ++u) acc[0] += v[u].x, acc[1] += v[u].y, acc[2] += v[u].z, acc[3] += v[u].w;
                }
            }
            for (int o = L >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            }
            if (b < img_end && sub == 0) {
                const float4 o4 = make_float4(static_cast<float>(acc[0]), static_cast<float>(acc[1]), static_cast<float>(acc[2]),
                                              static_cast<float>(acc[3]));
                reinterpret_cast<float4*>(a.per_image)[b] = o4;
                tot[0] += o4.x, tot[1] += o4.y, tot[2] += o4.z, tot[3] += o4.w;  // the total is the sum of the float32 per-image values
            }
        }
    }
}

// `red` is [8][4] doubles followed by 4 floats (dynamic shared memory); tot = this thread's share of the total
template <class P>
__device__ __forceinline__ void finalize_total(const LossArgs<P>& a, double (&tot)[4], double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!a.out_total) return;
    const unsigned int prev_seq = (a.use_comm && warp == 0) ? peer_allreduce_seq(a.comm) : 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) tot[k] = warp_sum_d(tot[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) red[warp * 4 + k] = tot[k];
    }
    __syncthreads();
    if (warp == 0) {
        float* vals = reinterpret_cast<float*>(red + (DH_THREADS / 32) * 4);  // (a static __shared__ array would count
        if (lane < 4) {                                                      // against the 227 KB dynamic opt-in)
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < DH_THREADS / 32; ++w) v += red[w * 4 + lane];
            vals[lane] = static_cast<float>(v);
        }
        __syncwarp();
        if (a.use_comm) peer_allreduce_warp(a.comm, vals, 4, prev_seq);
        __syncwarp();
        if (lane < 4) a.out_total[lane] = vals[lane];
    }
}

template <class P, int kCls, bool kGrad>
__global__ void __launch_bounds__(DH_THREADS, 4) fused_loss_kernel(const __grid_constant__ LossArgs<P> ga) {
    extern __shared__ __align__(128) unsigned char smem[];
    const FusedSmemLayout lay = fused_smem_layout<P>(ga.box_cap);
    const LossArgs<P>& a = *reinterpret_cast<const LossArgs<P>*>(smem + lay.args_off);  // see encode_kernel
    copy_args_to_smem(ga, reinterpret_cast<LossArgs<P>*>(smem + lay.args_off));
    typename P::Rec* recs = reinterpret_cast<typename P::Rec*>(smem + lay.rec_off);
    float* raw = reinterpret_cast<float*>(smem + lay.raw_off);
    uint64_t* boxbar = reinterpret_cast<uint64_t*>(smem + lay.misc_off);
    long long* next_chunk = reinterpret_cast<long long*>(smem + lay.misc_off + 8);
    int* is_last = reinterpret_cast<int*>(smem + lay.misc_off + 16);
    int* next_img = reinterpret_cast<int*>(smem + lay.misc_off + 24);  // [0] image whose GT rows are on their way into `raw` (-1: none), [1] its box count
    float* wred = reinterpret_cast<float*>(smem + lay.misc_off + 64);  // [8][4] floats; [8][4] doubles in the finalize

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SegLists segs = seg_lists(smem + lay.seg_off, ga.box_cap);
    const bool vec = ga.allow_vec && (ga.tt.ch & 3) == 0 && ga.spec.cen_mode == 0 && ga.spec.reg_ch == 4;
    const long long n_chunks = ga.n_chunks;
    long long chunk = blockIdx.x;  // (the launchers clamp the grid to the chunk count)
    if (tid == 0) {
        mbar_init(boxbar, 1);
        mbar_init_fence();
    }
    if (tid <= kMaxSegments) segs.nmap[tid] = 0;  // nmap[], npairs
    for (int e = tid; e < kMaxChunkTiles * DH_THREADS / 2; e += DH_THREADS) segs.head[e] = 0xffffffffu;
    if (tid < kMaxChunkTiles * 8) segs.rowbits[tid] = 0u;
    if (tid == 0) *segs.span_ctr = 0, next_img[0] = -1;
    const int span_batch = 32 * (kGrad ? 5 : 7);  // items of one load batch of stream_vec
    if (chunk < n_chunks) {
        int img, sub, t_begin, t_end;
        chunk_span(ga, chunk, img, sub, t_begin, t_end);
        build_tile_tab<P>(ga, segs, img, t_begin, t_end);
        __syncwarp();
        build_span_tab<P>(ga, segs, t_end - t_begin, vec, span_batch);
    }
    __syncthreads();

    // Warps 0..kHelpers-1 pair boxes with rows and resolve the target rows of a chunk while the others already stream it;
    // they join the streaming afterwards (the span counter gives them what is left).  With the gradient a matched row
    // overwrites the zero-label gradient the streaming pass wrote for it, so there the rows are resolved after a barrier.
    constexpr int kHelpers = 4;
    uint32_t box_parity = 0;
    int cur_img = -1, n_boxes = 0;
    const bool trace = ga.trace != nullptr && tid == 0;
    int n_done = 0;
    long long ph[6] = {0, 0, 0, 0, 0, 0}, pt = 0;  // trace: time thread 0 spends per phase of a chunk
#define DH_TRACE_PHASE(i)                                    \
    if (trace) {                                            \
        const long long now = static_cast<long long>(global_ns()); \
        ph[i] += now - pt;                                  \
        pt = now;                                           \
    }
    if (trace) ga.trace[blockIdx.x * kTraceSlots + 0] = pt = static_cast<long long>(global_ns());
#pragma unroll 1
    for (; chunk < n_chunks;) {
        int img, sub, t_begin, t_end;
        chunk_span(a, chunk, img, sub, t_begin, t_end);
        if (img != cur_img) {  // block-uniform; the barrier at the end of the previous chunk protects recs/raw
            if (next_img[0] == img) {  // the previous chunk asked for these rows: they have long arrived
                mbar_wait(boxbar, box_parity);
                box_parity ^= 1u;
                n_boxes = next_img[1];
            } else {
                n_boxes = stage_boxes(a.boxes, a.nbox, img, a.max_boxes, a.box_cap, raw, boxbar, box_parity);
            }
            const float hi = a.img_dim[2 * img], wi = a.img_dim[2 * img + 1];
            if (tid < n_boxes) P::make_record(a.pp, raw + 5 * tid, hi, wi, tid, recs[tid]);
            __syncthreads();
            cur_img = img;
        }
        // One thread of a streaming warp takes the next chunk from the scheduler and, when that chunk belongs to another
        // image, starts the bulk copy of its GT rows into `raw` (free from here on: the records are built) -- the counter's
        // and the copy's round trips pass under the streaming of this chunk instead of opening the next one.  All
        // max_boxes rows are copied (the count is read meanwhile), which needs 16-byte-sized and -aligned images.
        if (tid == DH_THREADS - 1) {
            const long long nc = static_cast<long long>(atomicAdd(ga.sched, 1u)) + gridDim.x;
            int pre = -1;
            if (nc < n_chunks) {
                int img2, sub2, tb2, te2;
                chunk_span(a, nc, img2, sub2, tb2, te2);
                const int rows = min(a.max_boxes, a.box_cap);
                const float* src = a.boxes + static_cast<long long>(img2) * a.max_boxes * 5;
                if (img2 != img && (rows & 3) == 0 && (a.max_boxes & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
                    mbar_expect_tx(boxbar, static_cast<uint32_t>(rows) * 20u);
                    bulk_g2s(raw, src, static_cast<uint32_t>(rows) * 20u, boxbar);
                    const int n2 = a.nbox ? a.nbox[img2] : a.max_boxes;
                    next_img[1] = max(0, min(n2, rows));
                    pre = img2;
                }
            }
            next_img[0] = pre;
            *next_chunk = nc;
        }
        if (sub == 0) P::image_prologue(a.pp, recs, n_boxes, img);
        DH_TRACE_PHASE(0)
        if (trace && n_done++ == 0) ga.trace[blockIdx.x * kTraceSlots + 1] = static_cast<long long>(global_ns());

        LossAcc acc = {0.f, 0.f, 0.f, 0};  // corrections + regression, natural units
        int n_pairs = 0;
        auto resolve = [&]() {  // helper warps, after their barrier: the pair list is complete
            n_pairs = *segs.npairs;
            if (tid == 0) segs.npairs[2] = n_pairs > kMaxPairs;  // (read at the chunk end, where npairs itself is being zeroed)
            if (n_pairs > kMaxPairs) {
                acc = visit_dense<P>(a, recs, segs, img, t_begin, t_end, kHelpers);
            } else if (n_pairs > 0) {
                if (warp == 0) {
                    const int n_rows = scan_rows(segs);
                    if (lane == 0) segs.npairs[1] = n_rows;
                }
                helper_barrier(32 * kHelpers);
                acc = resolve_pass<P>(a, recs, segs, segs.npairs[1], img, reinterpret_cast<unsigned short*>(smem + lay.run_off), kHelpers);
            }
        };
        if (n_boxes > 0 && warp < kHelpers) {
            pair_pass<P>(a, recs, n_boxes, segs, img, t_begin, t_end, kHelpers);
            helper_barrier(32 * kHelpers);
            DH_TRACE_PHASE(1)
            if (!kGrad) resolve();
            DH_TRACE_PHASE(4)
        }
        if (vec) stream_spans<P, kCls, kGrad, 1>(a, segs, img);
        else if (ga.allow_vec) stream_spans<P, kCls, kGrad, 2>(a, segs, img);
        else stream_spans<P, kCls, kGrad, 0>(a, segs, img);
        DH_TRACE_PHASE(2)
        if (kGrad) {
            __syncthreads();  // every zero-label gradient of the chunk is written before a matched row overwrites its own
            if (n_boxes > 0 && warp < kHelpers) resolve();
        }
        DH_TRACE_PHASE(3)

        // ---- per-chunk reduction -> partials[chunk] -----------------------------------------------------------
        const float reg = warp_sum(acc.reg), ccls = warp_sum(acc.cls), ccen = warp_sum(acc.cen);
        const int npos = warp_sum_i(acc.npos);
        if (lane == 0) {
            wred[warp * 4 + 0] = ccls, wred[warp * 4 + 1] = reg, wred[warp * 4 + 2] = ccen;
            wred[warp * 4 + 3] = static_cast<float>(npos);
        }
        __syncthreads();  // also: every warp is done with recs and with the spans; the prefetched chunk id is visible
        if (warp == 0) {  // span sums (in span order) + wred [8][4] -> one float4 per chunk
            const int n_spans = segs.span_ctr[1];
            float scls = 0.f, scen = 0.f;
            for (int q = lane; q < n_spans; q += 32) scls += segs.span_sum[2 * q], scen += segs.span_sum[2 * q + 1];
            scls = warp_sum(scls) * ((kCls == 2 ? 1.0f : 1.0f - ga.spec.alpha) * kLn2), scen = warp_sum(scen);
            float v = 0.f;
            if (lane < 4) {
#pragma unroll
                for (int w = 0; w < DH_THREADS / 32; ++w) v += wred[w * 4 + lane];
            }
            const float4 o = make_float4(__shfl_sync(0xffffffffu, v, 0) + scls, __shfl_sync(0xffffffffu, v, 1), __shfl_sync(0xffffffffu, v, 2) + scen,
                                         __shfl_sync(0xffffffffu, v, 3));
            if (lane == 0) reinterpret_cast<float4*>(a.partials)[chunk] = o;
        }
        chunk = *next_chunk;
        if (n_boxes > 0) {  // (the correct pass is done with the lists) counters and row chains back to empty
            if (tid <= kMaxSegments) segs.nmap[tid] = 0;
            if (segs.npairs[2]) {  // the pair list overflowed (resolve_pass leaves both tables empty; the dense visit does not use them)
                for (int e = tid; e < kMaxChunkTiles * DH_THREADS / 2; e += DH_THREADS) segs.head[e] = 0xffffffffu;
                if (tid < kMaxChunkTiles * 8) segs.rowbits[tid] = 0u;
            }
        }
        if (tid == DH_THREADS - 1) *segs.span_ctr = 0;
        if (chunk < n_chunks) {  // the next chunk's tile and span tables, published by the barrier below
            int img2, sub2, tb2, te2;
            chunk_span(a, chunk, img2, sub2, tb2, te2);
            build_tile_tab<P>(a, segs, img2, tb2, te2);
            __syncwarp();
            build_span_tab<P>(a, segs, te2 - tb2, vec, span_batch);
        }
        __syncthreads();  // wred / next_chunk are free again; the lists are empty; the tile table is the next chunk's
        DH_TRACE_PHASE(5)
    }
#undef DH_TRACE_PHASE
    if (trace) {
        ga.trace[blockIdx.x * kTraceSlots + 2] = static_cast<long long>(global_ns());
        ga.trace[blockIdx.x * kTraceSlots + 3] = n_done;
        for (int k = 0; k < 6; ++k) ga.trace[blockIdx.x * kTraceSlots + 4 + k] = ph[k];
    }
    // ---- the last CTA to get here finalizes ------------------------------------------------------------------
    if (tid == 0) {
        __threadfence();  // this CTA's partials and per-image rows are visible device-wide before it counts itself done
        const unsigned done = atomicAdd(ga.sched + 1, 1u);
        const int last = done == gridDim.x - 1;
        if (last) ga.sched[0] = 0u, ga.sched[1] = 0u;  // the scheduler line is ready for the next launch that gets it (kernel
        *is_last = last;                               // completion publishes the stores: nothing waits for them here)
    }
    if (!ga.fold_finalize) return;
    __syncthreads();
    if (*is_last) {
        double tot[4] = {0.0, 0.0, 0.0, 0.0};
        reduce_images<P>(a, tot);
        finalize_total<P>(a, tot, reinterpret_cast<double*>(wred));
        if (trace) ga.trace[static_cast<long long>(gridDim.x) * kTraceSlots] = static_cast<long long>(global_ns());
    }
}

}  // namespace dh
