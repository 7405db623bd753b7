// Encoder policies: the per-detector target rules, written as an order-free per-row gather.
//
// The reference paints boxes sequentially into NumPy maps ("last painter wins channels 0..4,
// class channels OR").  Each policy restates that as: (1) `make_record` -- per-GT quantities
// computed once per image by one thread per box, in the reference's float32 operation order;
// (2) `tile_hit` -- can this GT touch this tile at all; (3) `emit_row` -- for one output row, scan
// the tile's candidates, OR the class channels, pick the winner (= the painter the reference would
// have painted last) and write the row into the zero-initialised shared-memory tile.
// All citations are relative to /root/reference.
#pragma once
#include "dh_tile.cuh"

namespace dh {

enum FcosMode { FCOS_FOOTPRINT = 0, FCOS_CENTER3X3 = 1, FCOS_CENTER_ONLY = 2, FCOS_CENTER_V1 = 3, FCOS_MIN_AREA = 4 };
enum CenterNetMode { CN_ONEHOT_SCALES = 0, CN_HOURGLASS = 1, CN_POWER_FALLOFF = 2, CN_HOURGLASS4 = 3, CN_GAUSSIAN = 4 };
// FCOS_MIN_AREA and CN_GAUSSIAN are the canonical variants BASELINE's north_star names and the reference does not have
// (SURVEY.md section 0); their specifications are fcos_format_data(order="min_area") and
// centernet_gaussian_format_data of the CPU oracle.

struct Corners {
    float y0, x0, y1, x1;
};
// FCOS/fcos.py:211-215 (same expression in every encoder): (c -/+ 0.5*d) * img_dim
__device__ __forceinline__ Corners pixel_corners(const float* g, float hi, float wi) {
    Corners c;
    const float hh = fmul(0.5f, g[2]), hw = fmul(0.5f, g[3]);
    c.y0 = fmul(fsub(g[0], hh), hi);
    c.x0 = fmul(fsub(g[1], hw), wi);
    c.y1 = fmul(fadd(g[0], hh), hi);
    c.x1 = fmul(fadd(g[1], hw), wi);
    return c;
}
// A GT class outside [0, num_classes) would index past the row's class channels (the reference raises IndexError,
// FCOS/fcos.py:281-283): the box is dropped and bit 1 of the status word is set; the host wrappers raise.
__device__ __forceinline__ bool bad_class(int cls, int num_classes, int* status) {
    const bool bad = cls < 0 || cls >= num_classes;
    if (bad && status) atomicOr(status, 2);
    return bad;
}
// the reference's paint order is ascending area (stable): the later painter is the larger
// (area, index) pair.
__device__ __forceinline__ bool paints_later(float area, int k, float best_area, int best_k) {
    return best_k < 0 || area > best_area || (area == best_area && k > best_k);
}
// FCOS/fcos.py:262-271 in float64
__device__ __forceinline__ double ratio64(float a, float b) {
    const double lo = static_cast<double>(fminf(a, b)), hi = static_cast<double>(fmaxf(a, b));
    return ddiv(dadd(lo, 1.0e-8), dadd(hi, 1.0e-8));
}

// Where a policy writes the row it matched.  DenseSink is a zero-initialised row of the float32 target map
// (encoders, and the shared-memory tile of the unfused-compatible loss kernel); CompactSink keeps the row in
// registers as <= 5 regression values + a class bitmask (fused loss: targets never exist in memory at all).
struct DenseSink {
    static constexpr bool kExact = true;  // the targets themselves are the product: the reference's float64 arithmetic
    float* dst;
    __device__ __forceinline__ void cls(int first_class_ch, int c) { dst[first_class_ch + c] = 1.0f; }
    __device__ __forceinline__ void reg(int k, float v) { dst[k] = v; }
};
constexpr int kCompactClassWords = 4;  // fused loss fast path: up to 128 classes
struct CompactSink {
    static constexpr bool kExact = false;  // targets feed a loss with a 1e-5 bar: float32 arithmetic is enough
    float r[5];
    uint32_t m[kCompactClassWords];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < 5; ++k) r[k] = 0.f;
#pragma unroll
        for (int w = 0; w < kCompactClassWords; ++w) m[w] = 0u;
    }
    __device__ __forceinline__ void cls(int, int c) {
        const uint32_t bit = 1u << (c & 31);
        const int word = c >> 5;
#pragma unroll
        for (int w = 0; w < kCompactClassWords; ++w) m[w] |= (w == word) ? bit : 0u;
    }
    __device__ __forceinline__ void reg(int k, float v) {
#pragma unroll
        for (int q = 0; q < 5; ++q)
            if (q == k) r[q] = v;
    }
};

// =====================================================================================
// FCOS family: FCOS/fcos.py:136-378, fcos_center.py:149-279, fcos_center_v1.py:149-258
// =====================================================================================
struct FcosPolicy {
    struct Params {
        int n_levels, num_classes, mode;
        int stride[DH_MAX_LEVELS];
        float stride_f[DH_MAX_LEVELS];
        float b_dim[DH_MAX_LEVELS];  // n_levels-1 thresholds
        int hl[DH_MAX_LEVELS], wl[DH_MAX_LEVELS];
        int* num_targets;  // [B, n_levels] or null
        int* status;       // bit1: a GT class outside [0, num_classes) was dropped (the reference raises IndexError)
    };
    struct Rec {
        float y0s, x0s, y1s, x1s;  // corners / stride  (CENTER_V1: the four regression values)
        float area;
        int level;
        int ry0, ry1, rx0, rx1;  // half-open footprint in cells
        int ycen, xcen;
        int cls;
        int flags;  // bit0 live_y, bit1 live_x, bit2 valid
    };
    static constexpr int kRegCh = 5;  // t, b, l, r, centerness
    static constexpr bool kScatter = false;

    __device__ static void make_record(const Params& p, const float* g, float hi, float wi, int k, Rec& r) {
        const float gh = fmul(g[2], hi), gw = fmul(g[3], wi);  // fcos.py:152-153
        const float d = fmaxf(gw, gh);
        int l = p.n_levels - 1;  // fcos.py:168-179
        for (int n = 0; n < p.n_levels - 1; ++n)
            if (d < p.b_dim[n]) {
                l = n;
                break;
            }
        r.level = l;
        r.area = fmul(gh, gw);  // fcos.py:202-204
        r.cls = trunc_i(g[4]);
        const bool bad_cls = bad_class(r.cls, p.num_classes, p.status);
        const float s = p.stride_f[l];
        const int hl = p.hl[l], wl = p.wl[l];
        const Corners c = pixel_corners(g, hi, wi);
        int flags = 4;
        if (p.mode == FCOS_CENTER_V1) {  // fcos_center_v1.py:229-246
            const float box_sc = (l < p.n_levels - 1) ? p.b_dim[l] : fmaxf(hi, wi);
            const float raw_y = fmul(g[0], hi), raw_x = fmul(g[1], wi);
            const int i = trunc_i(fdiv(raw_y, s)), j = trunc_i(fdiv(raw_x, s));
            r.y0s = fdiv(fsub(raw_y, static_cast<float>(i * p.stride[l])), s);
            r.x0s = fdiv(fsub(raw_x, static_cast<float>(j * p.stride[l])), s);
            r.y1s = fdiv(gh, box_sc);
            r.x1s = fdiv(gw, box_sc);
            r.ry0 = i, r.ry1 = i + 1, r.rx0 = j, r.rx1 = j + 1;
            r.ycen = i, r.xcen = j;
            if (i < 0 || j < 0 || i >= hl || j >= wl || bad_cls) flags = 0;
            r.flags = flags;
            return;
        }
        r.y0s = fdiv(c.y0, s), r.x0s = fdiv(c.x0, s), r.y1s = fdiv(c.y1, s), r.x1s = fdiv(c.x1, s);
        const float h_ratio = fdiv(hi, s), w_ratio = fdiv(wi, s);  // fcos.py:162-163
        if (p.mode == FCOS_FOOTPRINT || p.mode == FCOS_MIN_AREA) {
            const float half_h = fdiv(g[2], 2.0f), half_w = fdiv(g[3], 2.0f);
            const int y_low = max(0, trunc_i(fmul(fsub(g[0], half_h), h_ratio)) + 1);  // fcos.py:217-225
            const int x_low = max(0, trunc_i(fmul(fsub(g[1], half_w), w_ratio)) + 1);
            const int y_upp = min(trunc_i(fmul(fadd(g[0], half_h), h_ratio)) + 1, hl);
            const int x_upp = min(trunc_i(fmul(fadd(g[1], half_w), w_ratio)) + 1, wl);
            r.ycen = min((y_low + y_upp) / 2, hl - 1);  // fcos.py:227-230 (C division truncates like int())
            r.xcen = min((x_low + x_upp) / 2, wl - 1);
            if (y_upp > y_low) {
                flags |= 1;
                r.ry0 = y_low, r.ry1 = y_upp;
            } else {
                r.ry0 = r.ycen, r.ry1 = r.ycen + 1;
            }
            if (x_upp > x_low) {
                flags |= 2;
                r.rx0 = x_low, r.rx1 = x_upp;
            } else {
                r.rx0 = r.xcen, r.rx1 = r.xcen + 1;
            }
            if (r.ry0 < 0 || r.rx0 < 0) flags = 0;  // box outside the image: out of contract
        } else {  // fcos_center.py:231-246
            const int rad = (p.mode == FCOS_CENTER_ONLY) ? 0 : 1;
            r.ycen = trunc_i(fadd(fmul(g[0], h_ratio), 0.5f));
            r.xcen = trunc_i(fadd(fmul(g[1], w_ratio), 0.5f));
            r.ry0 = r.ycen - rad, r.ry1 = r.ycen + rad + 1;
            r.rx0 = r.xcen - rad, r.rx1 = r.xcen + rad + 1;
        }
        r.flags = bad_cls ? 0 : flags;
    }

    __device__ static void image_prologue(const Params& p, const Rec* recs, int n, int b) {
        // num_targets[b][l] = GT count per level (fcos.py:376); written once per image
        if (p.num_targets && threadIdx.x < p.n_levels) {
            int c = 0;
            for (int k = 0; k < n; ++k) c += (recs[k].level == static_cast<int>(threadIdx.x));
            p.num_targets[b * p.n_levels + threadIdx.x] = c;
        }
    }

    __device__ static void tile_epilogue(const Params&, const TileInfo&, int) {}

    __device__ static bool tile_hit(const Params& p, const Rec& r, const TileInfo& ti, const MapDesc& md) {
        if (!(r.flags & 4) || r.level != ti.level) return false;
        const int i0 = static_cast<int>(fdiv_u32(ti.r0, md.div_width));
        const int i1 = static_cast<int>(fdiv_u32(ti.r0 + ti.nrows - 1, md.div_width));
        return r.ry1 > i0 && r.ry0 <= i1;
    }
    // two-level variant for small tiles (fused loss): map_hit is the tile-independent part, range_hit adds the
    // row range and, when the tile lies inside one map row, the column range
    __device__ static bool map_hit(const Params&, const Rec& r, int level, int) { return (r.flags & 4) && r.level == level; }
    __device__ static bool range_hit(const Params&, const Rec& r, const TileInfo& ti, const MapDesc& md) {
        const int i0 = static_cast<int>(fdiv_u32(ti.r0, md.div_width));
        const int last = ti.r0 + ti.nrows - 1;
        const int i1 = static_cast<int>(fdiv_u32(last, md.div_width));
        if (!(r.ry1 > i0 && r.ry0 <= i1)) return false;
        if (i0 != i1) return true;
        const int j0 = ti.r0 - i0 * md.width, j1 = last - i0 * md.width;
        return r.rx1 > j0 && r.rx0 <= j1;
    }

    // cells [ilo, ihi] x [jlo, jhi] of the map that hold every row this box can touch; false: none
    __device__ static bool cell_bounds(const Params&, const Rec& r, const MapDesc& md, int, int, int& ilo, int& ihi, int& jlo, int& jhi) {
        ilo = max(r.ry0, 0), ihi = min(r.ry1, md.height) - 1, jlo = max(r.rx0, 0), jhi = min(r.rx1, md.width) - 1;
        return ilo <= ihi && jlo <= jhi;
    }

    // may this box paint row `row` = cell (i, j)?  A superset is enough (match_row decides); here the rectangle is exact.
    __device__ static bool pair_hit(const Params&, const Rec&, const MapDesc&, int, int, int, int, int) { return true; }

    // returns the number of painters that touched the row (0 = row stays zero)
    __device__ static int emit_row(const Params& p, const TileInfo& ti, const MapDesc& md, int row, float* dst,
                                   const Rec* recs, const unsigned short* cand, int ncand) {
        DenseSink sink{dst};
        return match_row(p, ti, md, row, sink, recs, cand, ncand);
    }
    template <class Sink>
    __device__ static int match_row(const Params& p, const TileInfo& ti, const MapDesc& md, int row, Sink& dst,
                                    const Rec* recs, const unsigned short* cand, int ncand) {
        const int i = static_cast<int>(fdiv_u32(row, md.div_width));
        const int j = row - i * ti.width;
        int best = -1, hits = 0;
        float best_area = 0.f, best_score = 0.f;
        for (int q = 0; q < ncand; ++q) {
            const int k = cand[q];
            const Rec& r = recs[k];
            if (i < r.ry0 || i >= r.ry1 || j < r.rx0 || j >= r.rx1) continue;
            ++hits;
            dst.cls(kRegCh, r.cls);
            if (p.mode == FCOS_CENTER3X3 || p.mode == FCOS_CENTER_ONLY) {  // fcos_center.py:253-265
                const int dy = r.ycen - i, dx = r.xcen - j;
                const float sc = (dy == 0 && dx == 0) ? 1.0f : ((dy != 0 && dx != 0) ? 0.25f : 0.5f);
                best_score = fmaxf(best_score, sc);
            }
            // the reference paints in ascending-area order, so the largest covering box supplies channels 0..4
            // (fcos.py:202-209); FCOS_MIN_AREA is what its comment (:185-188) and the FCOS paper ask for: the smallest
            if (p.mode == FCOS_MIN_AREA ? (best < 0 || r.area < best_area || (r.area == best_area && k > best))
                                        : paints_later(r.area, k, best_area, best))
                best = k, best_area = r.area;
        }
        if (best < 0) return 0;
        const Rec& r = recs[best];
        if (p.mode == FCOS_CENTER_V1) {
            dst.reg(0, r.y0s), dst.reg(1, r.x0s), dst.reg(2, r.y1s), dst.reg(3, r.x1s), dst.reg(4, 1.0f);
            return hits;
        }
        const float fi = static_cast<float>(i) + 0.5f, fj = static_cast<float>(j) + 0.5f;
        if (p.mode != FCOS_FOOTPRINT && p.mode != FCOS_MIN_AREA) {  // fcos_center.py:267-273 (unclipped)
            dst.reg(0, fsub(fi, r.y0s));
            dst.reg(1, fsub(fsub(r.y1s, static_cast<float>(i)), 0.5f));
            dst.reg(2, fsub(fj, r.x0s));
            dst.reg(3, fsub(fsub(r.x1s, static_cast<float>(j)), 0.5f));
            dst.reg(4, best_score);
            return hits;
        }
        const bool live_y = r.flags & 1, live_x = r.flags & 2;
        // fcos.py:241-257 on a live axis; :289-301 / :325-337 / :357-369 on a collapsed axis (the
        // reference subtracts the integer centre and 0.5 separately there)
        const float t = fmaxf(0.f, fsub(fi, r.y0s));
        const float bt = live_y ? fmaxf(0.f, fsub(r.y1s, fi)) : fmaxf(0.f, fsub(fsub(r.y1s, static_cast<float>(i)), 0.5f));
        const float l = fmaxf(0.f, fsub(fj, r.x0s));
        const float rt = live_x ? fmaxf(0.f, fsub(r.x1s, fj)) : fmaxf(0.f, fsub(fsub(r.x1s, static_cast<float>(j)), 0.5f));
        dst.reg(0, t), dst.reg(1, bt), dst.reg(2, l), dst.reg(3, rt);
        float cen = 1.0f;  // fcos.py:371-372 when both axes collapsed
        if (i == r.ycen && j == r.xcen) {
            cen = 1.0f;  // fcos.py:279-280
        } else if (live_y || live_x) {
            const double qy = live_y ? ratio64(t, bt) : 1.0, qx = live_x ? ratio64(l, rt) : 1.0;
            cen = static_cast<float>(sqrt(dmul(qy, qx)));  // fcos.py:273-274
        }
        dst.reg(4, cen);
        return hits;
    }
};

// =====================================================================================
// RetinaNet: RetinaNet/retinanet_module.py:251-365 + RetinaNet/utils.py:42-83
// =====================================================================================
struct RetinaPolicy {
    struct Params {
        int n_levels, n_anchors, num_classes;
        int stride[DH_MAX_LEVELS];
        float anchor_h[DH_MAX_LEVELS][12], anchor_w[DH_MAX_LEVELS][12];
        float thr;
        int* num_pairs;  // [B] (zeroed by the launcher) or null
        int* status;     // see FcosPolicy::Params
    };
    struct Rec {
        float gy, gx, gh, gw;
        float lo_y, lo_x, hi_y, hi_x;
        float area;
        int cls;
    };
    static constexpr int kRegCh = 4;
    static constexpr bool kScatter = false;

    __device__ static void make_record(const Params& p, const float* g, float hi, float wi, int k, Rec& r) {
        r.gy = fmul(g[0], hi), r.gx = fmul(g[1], wi), r.gh = fmul(g[2], hi), r.gw = fmul(g[3], wi);  // :274-278
        const float hh = fdiv(r.gh, 2.0f), hw = fdiv(r.gw, 2.0f);                                    // utils.py:61-63
        r.lo_y = fsub(r.gy, hh), r.lo_x = fsub(r.gx, hw), r.hi_y = fadd(r.gy, hh), r.hi_x = fadd(r.gx, hw);
        r.area = fmul(r.gh, r.gw);
        r.cls = trunc_i(g[4]);
        if (bad_class(r.cls, p.num_classes, p.status)) r.cls = -1;  // never matches (tile_hit / map_hit)
    }
    __device__ static void image_prologue(const Params&, const Rec*, int, int) {}
    // positive (gt, anchor) pairs of the tile -> num_pairs[b] (integer atomics: order-independent)
    __device__ static void tile_epilogue(const Params& p, const TileInfo& ti, int pairs) {
        if (!p.num_pairs) return;
        const int s = warp_sum_i(pairs);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(p.num_pairs + ti.b, s);
    }

    // Necessary conditions for IoU > thr (thr > 0), used as cheap rejects before any division:
    //   IoU <= min(area)/max(area),  IoU <= overlap_y / max(ah, gh),  IoU <= overlap_x / max(aw, gw)
    // (inter <= oy * min(aw, gw) and union >= max(ah, gh) * min(aw, gw)).  0.999 covers float rounding.
    __device__ static bool tile_hit(const Params& p, const Rec& r, const TileInfo& ti, const MapDesc& md) {
        if (r.cls < 0) return false;
        if (p.thr < 0.f) return true;
        const float ah = p.anchor_h[ti.level][ti.anchor], aw = p.anchor_w[ti.level][ti.anchor];
        const float s = static_cast<float>(p.stride[ti.level]);
        const float i0 = static_cast<float>(fdiv_u32(ti.r0, md.div_width));
        const float i1 = static_cast<float>(fdiv_u32(ti.r0 + ti.nrows - 1, md.div_width));
        float need = 0.f;
        if (p.thr > 0.f) {
            const float aa = ah * aw;
            if (fminf(aa, r.area) < 0.999f * p.thr * fmaxf(aa, r.area)) return false;
            if (fminf(aw, r.gw) < 0.999f * p.thr * fmaxf(aw, r.gw)) return false;
            need = 0.999f * p.thr * fmaxf(ah, r.gh);
        }
        // some anchor row i in [i0, i1] must overlap the GT by more than `need` in y
        return i1 * s + 0.5f * ah > r.lo_y + need - 0.01f && i0 * s - 0.5f * ah < r.hi_y - need + 0.01f;
    }

    __device__ static bool map_hit(const Params& p, const Rec& r, int level, int anchor) {
        if (r.cls < 0) return false;
        if (!(p.thr > 0.f)) return true;
        const float ah = p.anchor_h[level][anchor], aw = p.anchor_w[level][anchor];
        const float aa = ah * aw, k = 0.999f * p.thr;
        return !(fminf(aa, r.area) < k * fmaxf(aa, r.area)) && !(fminf(aw, r.gw) < k * fmaxf(aw, r.gw)) &&
               !(fminf(ah, r.gh) < k * fmaxf(ah, r.gh));
    }
    __device__ static bool range_hit(const Params& p, const Rec& r, const TileInfo& ti, const MapDesc& md) {
        if (p.thr < 0.f) return true;
        const float ah = p.anchor_h[ti.level][ti.anchor], aw = p.anchor_w[ti.level][ti.anchor];
        const float s = static_cast<float>(p.stride[ti.level]);
        const int i0 = static_cast<int>(fdiv_u32(ti.r0, md.div_width));
        const int last = ti.r0 + ti.nrows - 1;
        const int i1 = static_cast<int>(fdiv_u32(last, md.div_width));
        const float k = p.thr > 0.f ? 0.999f * p.thr : 0.f;
        const float need_y = k * fmaxf(ah, r.gh);
        if (!(static_cast<float>(i1) * s + 0.5f * ah > r.lo_y + need_y - 0.01f &&
              static_cast<float>(i0) * s - 0.5f * ah < r.hi_y - need_y + 0.01f))
            return false;
        if (i0 != i1) return true;
        const float j0 = static_cast<float>(ti.r0 - i0 * md.width), j1 = static_cast<float>(last - i0 * md.width);
        const float need_x = k * fmaxf(aw, r.gw);
        return j1 * s + 0.5f * aw > r.lo_x + need_x - 0.01f && j0 * s - 0.5f * aw < r.hi_x - need_x + 0.01f;
    }

    // cells [ilo, ihi] x [jlo, jhi] of the (level, anchor) map that hold every anchor this box can match: the inverse of
    // range_hit's tests (anchor row i needs i*s + ah/2 > lo_y + need - 0.01 and i*s - ah/2 < hi_y - need + 0.01), widened
    // by a cell on each side; false: none
    __device__ static bool cell_bounds(const Params& p, const Rec& r, const MapDesc& md, int level, int anchor, int& ilo, int& ihi,
                                       int& jlo, int& jhi) {
        if (p.thr < 0.f) {
            ilo = 0, ihi = md.height - 1, jlo = 0, jhi = md.width - 1;
            return true;
        }
        const float ah = p.anchor_h[level][anchor], aw = p.anchor_w[level][anchor];
        const float inv_s = 1.0f / static_cast<float>(p.stride[level]);
        const float k = p.thr > 0.f ? 0.999f * p.thr : 0.f;
        const float need_y = k * fmaxf(ah, r.gh) - 0.01f, need_x = k * fmaxf(aw, r.gw) - 0.01f;
        const float hmax = static_cast<float>(md.height), wmax = static_cast<float>(md.width);
        ilo = static_cast<int>(fminf(fmaxf(floorf((r.lo_y + need_y - 0.5f * ah) * inv_s), 0.f), hmax));
        ihi = static_cast<int>(fminf(fmaxf(ceilf((r.hi_y - need_y + 0.5f * ah) * inv_s), -1.f), hmax - 1.f));
        jlo = static_cast<int>(fminf(fmaxf(floorf((r.lo_x + need_x - 0.5f * aw) * inv_s), 0.f), wmax));
        jhi = static_cast<int>(fminf(fmaxf(ceilf((r.hi_x - need_x + 0.5f * aw) * inv_s), -1.f), wmax - 1.f));
        return ilo <= ihi && jlo <= jhi;
    }

    // may this box match the anchor at cell (i, j)?  A superset is enough (match_row decides exactly): the IoU test without
    // the division and with a margin under the threshold
    __device__ static bool pair_hit(const Params& p, const Rec& r, const MapDesc&, int level, int anchor, int, int i, int j) {
        if (p.thr < 0.f) return true;
        const float ah = p.anchor_h[level][anchor], aw = p.anchor_w[level][anchor];
        const float s = static_cast<float>(p.stride[level]);
        const float ay = static_cast<float>(i) * s, ax = static_cast<float>(j) * s;
        const float dy = fmaxf(0.f, fminf(r.hi_y, ay + 0.5f * ah) - fmaxf(r.lo_y, ay - 0.5f * ah));
        const float dx = fmaxf(0.f, fminf(r.hi_x, ax + 0.5f * aw) - fmaxf(r.lo_x, ax - 0.5f * aw));
        const float inter = dy * dx;
        return inter > (p.thr - 1.0e-3f) * (r.area + ah * aw - inter) - 1.0e-6f;
    }

    // returns the number of (gt, anchor) pairs above the threshold at this row (:302-317)
    __device__ static int emit_row(const Params& p, const TileInfo& ti, const MapDesc& md, int row, float* dst,
                                   const Rec* recs, const unsigned short* cand, int ncand) {
        DenseSink sink{dst};
        return match_row(p, ti, md, row, sink, recs, cand, ncand);
    }
    template <class Sink>
    __device__ static int match_row(const Params& p, const TileInfo& ti, const MapDesc& md, int row, Sink& dst,
                                    const Rec* recs, const unsigned short* cand, int ncand) {
        const int i = static_cast<int>(fdiv_u32(row, md.div_width));
        const int j = row - i * ti.width;
        const int s = p.stride[ti.level];
        const float ah = p.anchor_h[ti.level][ti.anchor], aw = p.anchor_w[ti.level][ti.anchor];
        const float ay = static_cast<float>(i * s), ax = static_cast<float>(j * s);  // no half-cell offset (:293-294)
        const float hh = fdiv(ah, 2.0f), hw = fdiv(aw, 2.0f);
        const float alo_y = fsub(ay, hh), ahi_y = fadd(ay, hh), alo_x = fsub(ax, hw), ahi_x = fadd(ax, hw);
        const float a_area = fmul(ah, aw);
        const float pre = 0.999f * p.thr;
        int best = -1, pairs = 0;
        for (int q = 0; q < ncand; ++q) {  // candidates are in ascending GT order
            const int k = cand[q];
            const Rec& r = recs[k];
            const float dy = fmaxf(0.f, fsub(fminf(r.hi_y, ahi_y), fmaxf(r.lo_y, alo_y)));  // utils.py:66-72
            const float dx = fmaxf(0.f, fsub(fminf(r.hi_x, ahi_x), fmaxf(r.lo_x, alo_x)));
            if (p.thr > 0.f && (dy < pre * fmaxf(ah, r.gh) || dx < pre * fmaxf(aw, r.gw))) continue;  // see tile_hit
            const float inter = fmul(dy, dx);
            if (!(inter > 0.f) && !(p.thr < 0.f)) continue;  // IoU == 0 cannot exceed thr >= 0
            const float uni = fmaxf(fsub(fadd(r.area, a_area), inter), 1e-8f);  // utils.py:77-80
            const float iou = fminf(fmaxf(fdiv(inter, uni), 0.f), 1.f);
            if (iou > p.thr) {  // :302 strict
                ++pairs;
                dst.cls(kRegCh, r.cls);
                best = max(best, k);  // highest GT index wins the regression (:357), whatever the order of the list
            }
        }
        if (best >= 0) {  // :337-353, float64 like the reference's containers
            const Rec& r = recs[best];
            if constexpr (Sink::kExact) {
                const double dah = static_cast<double>(ah), daw = static_cast<double>(aw);
                dst.reg(0, static_cast<float>(ddiv(dsub(static_cast<double>(i * s), static_cast<double>(r.gy)), dah)));
                dst.reg(1, static_cast<float>(ddiv(dsub(static_cast<double>(j * s), static_cast<double>(r.gx)), daw)));
                dst.reg(2, static_cast<float>(ddiv(static_cast<double>(r.gh), dah)));
                dst.reg(3, static_cast<float>(ddiv(static_cast<double>(r.gw), daw)));
            } else {  // (four float64 divisions are ~250 instructions on the warp that holds up its CTA's chunk)
                const float iah = 1.0f / ah, iaw = 1.0f / aw;
                dst.reg(0, (ay - r.gy) * iah), dst.reg(1, (ax - r.gx) * iaw), dst.reg(2, r.gh * iah), dst.reg(3, r.gw * iaw);
            }
        }
        return pairs;
    }
};

// =====================================================================================
// CenterNet: tf_centernet_resnet_s8.py:243-330, tf_centernet_hourglass.py:379-456,
//            tf_centernet.py:152-342
// =====================================================================================
struct CenterNetPolicy {
    struct Params {
        int mode, num_classes, stride, n_scales;
        float stride_f, sigma;
        float scales[8];
        int pad0, pad1;  // img_pad[0], img_pad[1]
        int* status;     // bit0: a box was not below the largest scale (reference raises ValueError); bit1: bad class
    };
    struct Rec {
        float r0, r1, r2, r3;  // ONEHOT/HOURGLASS: regression values; FALLOFF: y0s, x0s, y1s, x1s
        float area;
        int row;                 // ONEHOT/HOURGLASS: target row, -1 if invalid
        int ry0, ry1, rx0, rx1;  // FALLOFF footprint
        int muy, mux;
        int cls;
        int flags;  // bit0 live_y, bit1 live_x, bit2 valid
        float gstd;  // CN_GAUSSIAN: max(1, sqrt(box area in cells)), tf_centernet.py:203-205
        float ginv_hi, ginv_lo;  // 1 / (2 gstd^2), a float64 value split into two floats (gauss_heat)
    };
    static constexpr int kRegChOnehot = 4;
    static constexpr bool kScatter = true;  // centre-cell modes: one thread per box writes its single row
    __device__ static bool footprint(const Params& p) { return p.mode == CN_POWER_FALLOFF || p.mode == CN_GAUSSIAN; }
    __device__ static bool use_scatter(const Params& p) { return !footprint(p); }

    __device__ static int reg_ch(const Params& p) { return (footprint(p) || p.mode == CN_HOURGLASS4) ? 5 : 4; }

    __device__ static void make_record(const Params& p, const float* g, float hi, float wi, int k, Rec& r) {
        const Corners c = pixel_corners(g, hi, wi);
        const float s = p.stride_f;
        r.area = fmul(fmul(g[2], hi), fmul(g[3], wi));
        r.cls = trunc_i(g[4]);
        r.flags = 4;
        r.row = -1;
        if (bad_class(r.cls, p.num_classes, p.status)) {
            r.flags = 0;
            return;
        }
        if (footprint(p)) {  // tf_centernet.py:165-223
            const float h_ratio = fdiv(hi, s), w_ratio = fdiv(wi, s);
            r.gstd = fmaxf(1.0f, __fsqrt_rn(fmul(fmul(fmul(g[2], g[3]), h_ratio), w_ratio)));  // :203-205 (before the 8.0 override)
            const double sd = static_cast<double>(r.gstd), ginv = 1.0 / (2.0 * sd * sd);
            r.ginv_hi = static_cast<float>(ginv), r.ginv_lo = static_cast<float>(ginv - static_cast<double>(r.ginv_hi));
            const int hl = static_cast<int>(static_cast<double>(p.pad0) / p.stride);
            const int wl = static_cast<int>(static_cast<double>(p.pad1) / p.stride);
            const int clip_h = trunc_i(fdiv(hi, s)), clip_w = trunc_i(fdiv(wi, s));  // clip by img_dim (:222-223)
            r.r0 = fdiv(c.y0, s), r.r1 = fdiv(c.x0, s), r.r2 = fdiv(c.y1, s), r.r3 = fdiv(c.x1, s);
            const int y_cen = trunc_i(fmul(g[0], h_ratio)), x_cen = trunc_i(fmul(g[1], w_ratio));
            const float sh = fdiv(fmul(p.sigma, g[2]), 2.0f), sw = fdiv(fmul(p.sigma, g[3]), 2.0f);
            const int y_low = max(0, 1 + trunc_i(fmul(fsub(g[0], sh), h_ratio)));
            const int x_low = max(0, 1 + trunc_i(fmul(fsub(g[1], sw), w_ratio)));
            const int y_upp = min(1 + trunc_i(fmul(fadd(g[0], sh), h_ratio)), clip_h);
            const int x_upp = min(1 + trunc_i(fmul(fadd(g[1], sw), w_ratio)), clip_w);
            int flags = 4;
            if (y_upp > y_low) {
                flags |= 1;
                r.ry0 = y_low, r.ry1 = y_upp, r.muy = (y_low + y_upp) / 2;
            } else {
                r.ry0 = y_cen, r.ry1 = y_cen + 1, r.muy = y_cen;
            }
            if (x_upp > x_low) {
                flags |= 2;
                r.rx0 = x_low, r.rx1 = x_upp, r.mux = (x_low + x_upp) / 2;
            } else {
                r.rx0 = x_cen, r.rx1 = x_cen + 1, r.mux = x_cen;
            }
            if (r.ry0 < 0 || r.rx0 < 0 || r.ry0 >= hl || r.rx0 >= wl) flags = 0;
            r.flags = flags;
            return;
        }
        if (p.mode == CN_HOURGLASS4) {  // the inline encoder of CenterNet/train_hourglass_voc.py:99-153
            const int n_map_y = static_cast<int>(static_cast<double>(p.pad0) / p.stride);
            const int n_map_x = static_cast<int>(static_cast<double>(p.pad1) / p.stride);
            const float pad_y = static_cast<float>(trunc_i(fdiv(fsub(static_cast<float>(p.pad0), hi), 2.0f)));  // :93
            const float pad_x = static_cast<float>(trunc_i(fdiv(fsub(static_cast<float>(p.pad1), wi), 2.0f)));
            const float y_cen = fadd(pad_y, fmul(g[0], hi)), x_cen = fadd(pad_x, fmul(g[1], wi));  // :118-119
            const float bh = fmul(g[2], hi), bw = fmul(g[3], wi);                                   // :120-121
            r.area = fmul(fmul(g[3], g[2]), 100.0f);                                                // :110-111
            if (bw < 0.f || bh < 0.f) {  // :123-124
                r.flags = 0;
                return;
            }
            int sc = p.n_scales - 1;  // :125-138: the first scale that exceeds BOTH sides, else the last
            for (int n = 0; n < p.n_scales - 1; ++n)
                if (bw < p.scales[n] && bh < p.scales[n]) {
                    sc = n;
                    break;
                }
            const int i = trunc_i(fdiv(y_cen, s)), j = trunc_i(fdiv(x_cen, s));  // :142-143
            if (i < 0 || j < 0 || i >= n_map_y || j >= n_map_x) {
                r.flags = 0;
                return;
            }
            r.r0 = fdiv(fsub(y_cen, static_cast<float>(i * p.stride)), s);  // :144-145
            r.r1 = fdiv(fsub(x_cen, static_cast<float>(j * p.stride)), s);
            r.r2 = fdiv(bh, p.scales[sc]);                                   // :140-141
            r.r3 = fdiv(bw, p.scales[sc]);
            r.row = (i * n_map_x + j) * p.n_scales + sc;
            return;
        }
        // centre-cell encoders; note the reference's swapped pad indices (tf_centernet_resnet_s8.py:259-262)
        const int h_max = static_cast<int>(static_cast<double>(p.pad1) / p.stride);
        const int w_max = static_cast<int>(static_cast<double>(p.pad0) / p.stride);
        const float pad_y = static_cast<float>(trunc_i(fdiv(fsub(static_cast<float>(p.pad1), wi), 2.0f)));
        const float pad_x = static_cast<float>(trunc_i(fdiv(fsub(static_cast<float>(p.pad0), hi), 2.0f)));
        const float yc = fdiv(fadd(c.y0, c.y1), 2.0f), xc = fdiv(fadd(c.x0, c.x1), 2.0f);
        const float py = fadd(pad_y, yc), px = fadd(pad_x, xc);
        const int i = trunc_i(fdiv(py, s)), j = trunc_i(fdiv(px, s));  // :310-313
        if (i < 0 || j < 0 || i >= h_max || j >= w_max) {
            r.flags = 0;
            return;
        }
        if (p.mode == CN_ONEHOT_SCALES) {
            const float bh = fsub(c.y1, c.y0), bw = fsub(c.x1, c.x0);
            const float d = fmaxf(bh, bw);
            int sc = -1;
            for (int n = 0; n < p.n_scales; ++n)
                if (d < p.scales[n]) {
                    sc = n;
                    break;
                }
            if (sc < 0) {  // reference: min([]) -> ValueError (:306-307); the host wrapper raises
                if (p.status) atomicOr(p.status, 1);
                r.flags = 0;
                return;
            }
            r.r0 = fdiv(fsub(py, static_cast<float>(i * p.stride)), s);  // :316-322
            r.r1 = fdiv(fsub(px, static_cast<float>(j * p.stride)), s);
            r.r2 = fdiv(bh, p.scales[sc]);
            r.r3 = fdiv(bw, p.scales[sc]);
            r.row = (i * w_max + j) * p.n_scales + sc;
        } else {  // tf_centernet_hourglass.py:445-449
            r.r0 = fsub(static_cast<float>(i) + 0.5f, fdiv(fadd(pad_y, c.y0), s));
            r.r1 = fsub(fsub(fdiv(fadd(pad_y, c.y1), s), static_cast<float>(i)), 0.5f);
            r.r2 = fsub(static_cast<float>(j) + 0.5f, fdiv(fadd(pad_x, c.x0), s));
            r.r3 = fsub(fsub(fdiv(fadd(pad_x, c.x1), s), static_cast<float>(j)), 0.5f);
            r.row = i * w_max + j;
        }
    }
    __device__ static void image_prologue(const Params&, const Rec*, int, int) {}
    __device__ static void tile_epilogue(const Params&, const TileInfo&, int) {}

    __device__ static bool tile_hit(const Params& p, const Rec& r, const TileInfo& ti, const MapDesc& md) {
        if (!(r.flags & 4)) return false;
        if (!footprint(p)) return r.row >= ti.r0 && r.row < ti.r0 + ti.nrows;
        const int i0 = static_cast<int>(fdiv_u32(ti.r0, md.div_width));
        const int i1 = static_cast<int>(fdiv_u32(ti.r0 + ti.nrows - 1, md.div_width));
        return r.ry1 > i0 && r.ry0 <= i1;
    }
    __device__ static bool map_hit(const Params&, const Rec& r, int, int) { return (r.flags & 4) != 0; }
    __device__ static bool range_hit(const Params& p, const Rec& r, const TileInfo& ti, const MapDesc& md) {
        if (!footprint(p)) return r.row >= ti.r0 && r.row < ti.r0 + ti.nrows;
        const int i0 = static_cast<int>(fdiv_u32(ti.r0, md.div_width));
        const int last = ti.r0 + ti.nrows - 1;
        const int i1 = static_cast<int>(fdiv_u32(last, md.div_width));
        if (!(r.ry1 > i0 && r.ry0 <= i1)) return false;
        if (i0 != i1) return true;
        const int j0 = ti.r0 - i0 * md.width, j1 = last - i0 * md.width;
        return r.rx1 > j0 && r.rx0 <= j1;
    }

    __device__ static bool cell_bounds(const Params& p, const Rec& r, const MapDesc& md, int, int, int& ilo, int& ihi, int& jlo, int& jhi) {
        if (!footprint(p)) {  // one row: its cell
            const int cell = static_cast<int>(fdiv_u32(static_cast<uint32_t>(r.row), md.div_sub));
            ilo = ihi = static_cast<int>(fdiv_u32(static_cast<uint32_t>(cell), md.div_width));
            jlo = jhi = cell - ilo * md.width;
            return r.row >= 0;
        }
        ilo = max(r.ry0, 0), ihi = min(r.ry1, md.height) - 1, jlo = max(r.rx0, 0), jhi = min(r.rx1, md.width) - 1;
        return ilo <= ihi && jlo <= jhi;
    }

    __device__ static bool pair_hit(const Params& p, const Rec& r, const MapDesc&, int, int, int row, int, int) {
        return footprint(p) || r.row == row;
    }

    // Scatter emission for the centre-cell modes: thread q owns candidate q, whose target is a single
    // row.  Class bits are idempotent stores; the regression values are written only by the box no
    // other candidate of the same row paints after (pairwise check instead of atomics).
    __device__ static void emit_tile(const Params& p, const TileInfo& ti, const MapDesc& md, float* tile, int ch,
                                     const Rec* recs, const unsigned short* cand, int ncand, unsigned short* dirty_rows) {
        for (int q = threadIdx.x; q < ncand; q += blockDim.x) {
            const int k = cand[q];
            const Rec& r = recs[k];
            const int local = r.row - ti.r0;
            float* dst = tile + local * ch;
            if (p.mode == CN_HOURGLASS4) dst[4] = 1.0f, dst[5 + r.cls] = 1.0f;  // objectness + class (:148-150)
            else dst[4 + r.cls] = 1.0f;
            bool wins = true;
            for (int q2 = 0; q2 < ncand; ++q2) {
                const int k2 = cand[q2];
                if (k2 != k && recs[k2].row == r.row && paints_later(recs[k2].area, k2, r.area, k)) wins = false;
            }
            if (wins) dst[0] = r.r0, dst[1] = r.r1, dst[2] = r.r2, dst[3] = r.r3;
            dirty_rows[q] = static_cast<unsigned short>(local);
        }
    }

    // gaussian_dist_2d of tf_centernet.py:30-40 on the footprint grid (cells at z + 0.5, integer mean): exp(-d^2 / (2 std^2))
    // over the live axes, divided by its maximum over the footprint (reached 0.5 cells from the mean on every live axis);
    // the centre cell is forced to 1 like the reference does for its fall-off (:261-262).  The specification
    // (centernet_gaussian_format_data of the CPU oracle) evaluates the exponent and exp in float64 and rounds once; here the exponent
    // is carried as an unevaluated sum of two floats -- d^2 is exact in float32, 1 / (2 std^2) is a float64 split in two,
    // the product's rounding error comes out of an fma -- and exp(hi + lo) = expf(hi) * (1 + lo): within 2 ulp of float32
    // of the specified value (the parity tests allow 2e-6 relative), for a dozen float32 instructions instead of a
    // float64 division and a float64 exp per covering box per cell.
    __device__ static float gauss_heat(const Rec& c, int i, int j) {
        const bool live_y = c.flags & 1, live_x = c.flags & 2;
        if ((i == c.muy && j == c.mux) || !(live_y || live_x)) return 1.0f;
        float d2 = 0.0f;  // (integers and halves below 2^11: every operation is exact)
        if (live_y) {
            const float dy = (static_cast<float>(i) + 0.5f) - static_cast<float>(c.muy);
            d2 += dy * dy - 0.25f;
        }
        if (live_x) {
            const float dx = (static_cast<float>(j) + 0.5f) - static_cast<float>(c.mux);
            d2 += dx * dx - 0.25f;
        }
        const float hi = fmul(-d2, c.ginv_hi);
        const float lo = fmaf(-d2, c.ginv_hi, -hi) - d2 * c.ginv_lo;
        return expf(hi) * (1.0f + lo);
    }

    __device__ static double inv_pow8(double d) {  // 1/(d^8): tf_centernet.py:6-19 with spread forced to 8 (:207)
        const double d2 = dmul(d, d), d4 = dmul(d2, d2);
        return ddiv(1.0, dmul(d4, d4));
    }

    // Encoders: the 32 rows of a warp are neighbouring cells, so the warp first ballots which candidates' footprints meet
    // the rectangle its cells span (one candidate per lane and round) and each lane then walks only those -- a tile's
    // candidate list holds every box that touches its two image rows, a cell is covered by a handful of them.
    static constexpr int kCandWords = (DH_MAX_BOXES + 31) / 32;
    __device__ static int emit_row(const Params& p, const TileInfo& ti, const MapDesc& md, int row, float* dst,
                                   const Rec* recs, const unsigned short* cand, int ncand) {
        DenseSink sink{dst};
        if (footprint(p) && __activemask() == 0xffffffffu && ncand <= 32 * kCandWords) {
            const int lane = threadIdx.x & 31;
            const int i = static_cast<int>(fdiv_u32(row, md.div_width)), j = row - i * ti.width;
            const int i0 = __reduce_min_sync(0xffffffffu, i), i1 = __reduce_max_sync(0xffffffffu, i);
            const int j0 = __reduce_min_sync(0xffffffffu, j), j1 = __reduce_max_sync(0xffffffffu, j);
            unsigned near[kCandWords];
#pragma unroll
            for (int w = 0; w < kCandWords; ++w) {
                bool hit = false;
                if (32 * w + lane < ncand) {
                    const Rec& r = recs[cand[32 * w + lane]];
                    hit = i1 >= r.ry0 && i0 < r.ry1 && j1 >= r.rx0 && j0 < r.rx1;
                }
                near[w] = 32 * w < ncand ? __ballot_sync(0xffffffffu, hit) : 0u;
            }
            return match_row<DenseSink, true>(p, ti, md, row, sink, recs, cand, ncand, near);
        }
        return match_row(p, ti, md, row, sink, recs, cand, ncand);
    }
    template <class Sink, bool kNear = false>
    __device__ static int match_row(const Params& p, const TileInfo& ti, const MapDesc& md, int row, Sink& dst,
                                    const Rec* recs, const unsigned short* cand, int ncand, const unsigned* near = nullptr) {
        int best = -1, hits = 0;
        float best_area = 0.f;
        if (!footprint(p)) {
            for (int q = 0; q < ncand; ++q) {
                const int k = cand[q];
                const Rec& r = recs[k];
                if (r.row != row) continue;
                ++hits;
                if (p.mode == CN_HOURGLASS4) dst.cls(4, 0), dst.cls(4, r.cls + 1);  // ch 4 = objectness, then the classes
                else dst.cls(4, r.cls);
                if (paints_later(r.area, k, best_area, best)) best = k, best_area = r.area;
            }
            if (best < 0) return 0;
            const Rec& r = recs[best];
            dst.reg(0, r.r0), dst.reg(1, r.r1), dst.reg(2, r.r2), dst.reg(3, r.r3);
            return hits;
        }
        const int i = static_cast<int>(fdiv_u32(row, md.div_width));
        const int j = row - i * ti.width;
        float heat = 0.f;  // CN_GAUSSIAN: the maximum over the covering boxes (see below), gathered in the same walk
        auto visit = [&](int k) {
            const Rec& r = recs[k];
            if (i < r.ry0 || i >= r.ry1 || j < r.rx0 || j >= r.rx1) return;
            ++hits;
            dst.cls(5, r.cls);
            if (paints_later(r.area, k, best_area, best)) best = k, best_area = r.area;
            if (p.mode == CN_GAUSSIAN) heat = fmaxf(heat, gauss_heat(r, i, j));
        };
        if constexpr (kNear) {
#pragma unroll
            for (int w = 0; w < kCandWords; ++w)
                for (unsigned m = near[w]; m; m &= m - 1) visit(cand[32 * w + __ffs(m) - 1]);
        } else {
            for (int q = 0; q < ncand; ++q) visit(cand[q]);
        }
        if (best < 0) return 0;
        const Rec& r = recs[best];
        const bool live_y = r.flags & 1, live_x = r.flags & 2;
        const float fi = static_cast<float>(i) + 0.5f, fj = static_cast<float>(j) + 0.5f;
        dst.reg(0, fmaxf(0.f, fsub(fi, r.r0)));
        dst.reg(1, live_y ? fmaxf(0.f, fsub(r.r2, fi)) : fmaxf(0.f, fsub(fsub(r.r2, static_cast<float>(i)), 0.5f)));
        dst.reg(2, fmaxf(0.f, fsub(fj, r.r1)));
        dst.reg(3, live_x ? fmaxf(0.f, fsub(r.r3, fj)) : fmaxf(0.f, fsub(fsub(r.r3, static_cast<float>(j)), 0.5f)));
        // CN_GAUSSIAN, the canonical CenterNet splat: the heat of a cell is the MAXIMUM over the boxes whose footprint covers it
        // of gaussian_dist_2d (the reference's commented-out code, tf_centernet.py:30-40) -- order-free, no atomics; gathered above
        if (p.mode != CN_GAUSSIAN) {
            heat = 1.0f;
            if (!(i == r.muy && j == r.mux) && (live_y || live_x)) {
                // max over the footprint is reached at the centre cell: 1/0.5^8 = 256 per live axis
                double v = 1.0;
                if (live_y) v = dmul(v, inv_pow8(static_cast<double>(fi) - static_cast<double>(r.muy)) * (1.0 / 256.0));
                if (live_x) v = dmul(v, inv_pow8(static_cast<double>(fj) - static_cast<double>(r.mux)) * (1.0 / 256.0));
                heat = static_cast<float>(v);
            }
        }
        dst.reg(4, heat);
        return hits;
    }
};

}  // namespace dh
