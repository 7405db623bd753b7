// Dense-head losses as a streaming kernel: focal (classes / centerness), smooth-L1 or -log(IoU)
// (boxes), smooth-L1 of sigmoid (centerness) over (prediction, target) maps.
//
// The kernel is bound by one read of the predictions (and, unfused, one of the targets), with about
// 3 MUFU + 15 ALU instructions of transcendental work per element, so the structure is:
//   * predictions are read straight from HBM with 128-bit loads, four independent loads in flight per
//     thread (no shared-memory staging: every element is used exactly once);
//   * targets come either from HBM the same way (unfused: `format_data` output -> `model_loss`) or are
//     produced on the fly by the encoder policy into a small zeroed shared-memory tile (fused
//     encode+loss: targets never touch HBM).  A tile that no GT box can touch -- the common case --
//     takes a label-free fast path and never looks at shared memory;
//   * tiles are handed out in image-aligned chunks by a device-side counter; a CTA keeps its partial
//     sums in registers for a whole chunk and reduces once per chunk, so a tile costs one block barrier;
//   * per-chunk partials {cls, reg, cen, n_pos} are summed per image by a second tiny kernel in a fixed
//     order with float64 accumulation: results are deterministic run to run.
//
// Reference formulas: FCOS/fcos.py:380-496 (identical copies in the other modules).
#pragma once
#include "dh_comm.cuh"
#include "dh_encode_kernel.cuh"  // build_candidates / total_candidates

namespace dh {

struct LossSpec {
    int reg_ch;    // 0 or 4: channels [0, reg_ch) are box regression
    int cen_mode;  // 0 no centerness channel, 1 smooth-L1(sigmoid(pred)) over all rows, 2 focal, 3 present but unused
    int reg_mode;  // 0 smooth-L1, 1 -log(IoU) on the integer grid, 2 1 - GIoU on the same boxes
    int pos_rule;  // 0 max(class) >= 1, 1 max(class) > 0, 2 external per-row mask
    int cls_mode;  // 0 focal, 1 sigmoid cross-entropy (alpha / gamma unused)
    float alpha, gamma, delta;
    float w_cls, w_reg, w_cen;  // gradient weights: grad = d(w_cls*cls + w_reg*reg + w_cen*cen) / d pred
};

struct NoPolicy {
    struct Params {
        int unused;
    };
    struct Rec {
        int unused;
    };
};

constexpr int kMaxChunkTiers = 8;
struct ChunkTier {
    long long chunk0;  // id of the tier's first chunk
    int image0;        // first image of the tier
    int chunk_tiles;   // tiles per chunk
    int cpi;           // chunks per image
    int pad_;
};

template <class P>
struct LossArgs {
    TileTable tt;  // maps[m].pred = predictions, maps[m].out = targets (unfused)
    typename P::Params pp;
    LossSpec spec;
    const float* boxes;
    const int* nbox;
    const float* img_dim;
    int max_boxes;
    int box_cap;         // shared-memory capacity in boxes (fused): max_boxes rounded up to 32
    int tile_buf_bytes;  // fused: bytes of the shared-memory target tile
    int chunk_tiles;     // tiles per scheduler chunk; an image is cut into chunks_per_image chunks
    int chunks_per_image;
    int allow_vec;  // every map pointer is 16-byte aligned
    unsigned int* sched;
    float* partials;  // [n_chunks, 4]
    // fused stream + correct kernel only: the batch is cut into up to four runs of images ("tiers") with ever finer
    // chunks, so that the launch ends on short chunks (tail = one fine chunk, not one coarse one); the last CTA to
    // finish reduces the partials (per image, then the total) and, when asked, exchanges the total with the peer ranks
    int n_tiers;
    ChunkTier tiers[kMaxChunkTiers];
    long long n_chunks;
    float* per_image;  // [batch, 4] (caller's buffer or scratch)
    float* out_total;  // [4] or null
    int fold_finalize;  // 1: the last CTA finalizes (no finalize kernels follow)
    int span_fine;      // log2 of the spans the LAST tile of a chunk is cut into (fused kernel, see spans_of_tile)
    int use_comm;       // 1: out_total is summed over the ranks of `comm` (peer mailboxes)
    CommDev comm;
    long long* trace;  // profiling aid (dh_set_trace): per CTA {start, first chunk ready, loop end (globaltimer ns), chunks}
    const float* mask_maps[DH_MAX_MAPS];
    float* grad_maps[DH_MAX_MAPS];  // null: forward only; else d loss / d pred, same layout as the predictions
};

struct LossSmemLayout {
    int tgt_off, rowpos_off, rec_off, raw_off, cand_off, cand2_off, misc_off, args_off, total;
};
template <class P, bool kFused>
__host__ __device__ inline LossSmemLayout loss_smem_layout(int tile_buf_bytes, int rows_per_tile, int box_cap) {
    LossSmemLayout l;
    l.tgt_off = 0;
    l.rowpos_off = kFused ? tile_buf_bytes : 0;
    l.rec_off = l.rowpos_off + ((rows_per_tile * 4 + 127) & ~127);
    l.raw_off = l.rec_off + (kFused ? ((static_cast<int>(sizeof(typename P::Rec)) * box_cap + 127) & ~127) : 0);
    l.cand_off = l.raw_off + (kFused ? ((box_cap * 20 + 127) & ~127) : 0);
    l.cand2_off = l.cand_off + (kFused ? 2 * DH_THREADS * 2 : 0);  // per-warp segments, double-buffered by tile parity
    l.misc_off = l.cand2_off + (kFused ? ((box_cap * 2 + 127) & ~127) : 0);
    l.args_off = l.misc_off + 256;
    l.total = l.args_off + ((static_cast<int>(sizeof(LossArgs<P>)) + 127) & ~127);
    return l;
}

// ---- element formulas (float32) ---------------------------------------------------------------
struct ExpParts {
    float e, inv, soft;  // exp(-|x|), 1/(1+e), log(1+e)
};
__device__ __forceinline__ ExpParts exp_parts(float x) {
    ExpParts p;
    p.e = __expf(-fabsf(x));
    const float w = 1.0f + p.e;
    p.inv = __fdividef(1.0f, w);
    p.soft = __logf(w);  // log(1 + exp(-|x|)), as the reference writes it
    return p;
}
// focal: y*a*(1-s)^g*softplus(-x) + (1-y)*(1-a)*s^g*softplus(x)   (FCOS/fcos.py:443-462, stable form)
__device__ __forceinline__ float focal_term(float y, float x, float alpha, float gamma) {
    const ExpParts p = exp_parts(x);
    const float s = x >= 0.f ? p.inv : p.e * p.inv;   // sigmoid(x)
    const float om = x >= 0.f ? p.e * p.inv : p.inv;  // 1 - sigmoid(x)
    float p_pos, p_neg;
    if (gamma == 2.0f) {
        p_pos = om * om, p_neg = s * s;
    } else {
        p_pos = __powf(om, gamma), p_neg = __powf(s, gamma);
    }
    const float sp_pos = p.soft + fmaxf(x, 0.f);  // softplus(x)
    const float sp_neg = p.soft - fminf(x, 0.f);  // softplus(-x)
    return y * alpha * p_pos * sp_neg + (1.0f - y) * (1.0f - alpha) * p_neg * sp_pos;
}
// the same with y == 0: (1-a) * s^g * softplus(x)
__device__ __forceinline__ float focal_neg(float x, float one_minus_alpha, float gamma) {
    const ExpParts p = exp_parts(x);
    const float s = x >= 0.f ? p.inv : p.e * p.inv;
    const float pw = gamma == 2.0f ? s * s : __powf(s, gamma);
    return one_minus_alpha * pw * (p.soft + fmaxf(x, 0.f));
}
__device__ __forceinline__ float sigmoid_f(float x) {
    const float e = __expf(-fabsf(x));
    const float inv = __fdividef(1.0f, 1.0f + e);
    return x >= 0.f ? inv : e * inv;
}
__device__ __forceinline__ float smooth_l1_term(float y, float x, float delta) {
    const float d = y - x, ad = fabsf(d);
    return ad < delta ? 0.5f * d * d : ad;  // no -delta/2 (FCOS/fcos.py:386-388)
}
// Box loss of two tblr boxes anchored at the integer grid point (gx, gy).  mode 1: -log(IoU), FCOS/fcos.py:393-441.
// mode 2 (DH_REG_GIOU, an extension: BASELINE's north_star names it, the reference has no GIoU): 1 - GIoU on the same
// box construction, GIoU = IoU - (C - union) / (C + 1e-12) with C the area of the smallest enclosing box.
__device__ __forceinline__ float iou_loss_term(const float* t, const float* p, float gy, float gx, int mode = 1) {
    const float ty0 = gy - t[0], ty1 = gy + t[1], tx0 = gx - t[2], tx1 = gx + t[3];
    const float py0 = gy - p[0], py1 = gy + p[1], px0 = gx - p[2], px1 = gx + p[3];
    const float ih = fmaxf(0.f, fminf(ty1, py1) - fmaxf(ty0, py0));
    const float iw = fmaxf(0.f, fminf(tx1, px1) - fmaxf(tx0, px0));
    const float inter = iw * ih;
    const float uni = ((ty1 - ty0) * (tx1 - tx0) + (py1 - py0) * (px1 - px0)) - inter;
    const float iou = inter / (uni + 1.0e-12f);
    if (mode == 2) {
        const float ac = (fmaxf(ty1, py1) - fminf(ty0, py0)) * (fmaxf(tx1, px1) - fminf(tx0, px0));
        return 1.0f - (iou - (ac - uni) / (ac + 1.0e-12f));
    }
    return -logf(iou + 1.0e-12f);
}

// ---- derivatives with respect to the prediction (what TF autodiff yields for the formulas above) -----------------
// d/dx [ y*a*(1-s)^g*softplus(-x) + (1-y)*(1-a)*s^g*softplus(x) ]
__device__ __forceinline__ float focal_grad(float y, float x, float alpha, float gamma) {
    const ExpParts p = exp_parts(x);
    const float s = x >= 0.f ? p.inv : p.e * p.inv;
    const float om = x >= 0.f ? p.e * p.inv : p.inv;
    float p_pos, p_neg;
    if (gamma == 2.0f) {
        p_pos = om * om, p_neg = s * s;
    } else {
        p_pos = __powf(om, gamma), p_neg = __powf(s, gamma);
    }
    const float sp_pos = p.soft + fmaxf(x, 0.f);  // softplus(x)
    const float sp_neg = p.soft - fminf(x, 0.f);  // softplus(-x)
    const float d_pos = -(gamma * s * p_pos * sp_neg + p_pos * om);
    const float d_neg = gamma * p_neg * om * sp_pos + p_neg * s;
    return y * alpha * d_pos + (1.0f - y) * (1.0f - alpha) * d_neg;
}
// d/dx where(|y-x| < delta, (y-x)^2/2, |y-x|)
__device__ __forceinline__ float smooth_l1_grad(float y, float x, float delta) {
    const float d = x - y, ad = fabsf(d);
    return ad < delta ? d : (d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.f));
}
// d/dx smooth_l1(y, sigmoid(x))
__device__ __forceinline__ float cen_l1_grad(float y, float x, float delta) {
    const float s = sigmoid_f(x);
    return smooth_l1_grad(y, s, delta) * s * (1.0f - s);
}
// gradient of iou_loss_term with respect to the four predicted distances (t, b, l, r)
__device__ __forceinline__ void iou_loss_grad(const float* t, const float* p, float gy, float gx, float* g, int mode = 1) {
    const float ty0 = gy - t[0], ty1 = gy + t[1], tx0 = gx - t[2], tx1 = gx + t[3];
    const float py0 = gy - p[0], py1 = gy + p[1], px0 = gx - p[2], px1 = gx + p[3];
    const float ih_raw = fminf(ty1, py1) - fmaxf(ty0, py0), iw_raw = fminf(tx1, px1) - fmaxf(tx0, px0);
    const float ih = fmaxf(0.f, ih_raw), iw = fmaxf(0.f, iw_raw);
    const float inter = iw * ih;
    const float ph = py1 - py0, pw = px1 - px0;
    const float uni = ((ty1 - ty0) * (tx1 - tx0) + ph * pw) - inter;
    const float den = uni + 1.0e-12f;
    const float iou = inter / den;
    const float live_h = ih_raw > 0.f ? 1.0f : 0.f, live_w = iw_raw > 0.f ? 1.0f : 0.f;
    float di[4], da[4];  // d inter / d p_k, d area_pred / d p_k
    di[0] = iw * live_h * (py0 > ty0 ? 1.0f : 0.f);  // the top edge of the intersection is the prediction's
    di[1] = iw * live_h * (py1 < ty1 ? 1.0f : 0.f);
    di[2] = ih * live_w * (px0 > tx0 ? 1.0f : 0.f);
    di[3] = ih * live_w * (px1 < tx1 ? 1.0f : 0.f);
    da[0] = da[1] = pw, da[2] = da[3] = ph;
    if (mode == 2) {  // d (1 - GIoU) = -(d IoU - d ((C - union) / (C + eps)))
        const float eh = fmaxf(ty1, py1) - fminf(ty0, py0), ew = fmaxf(tx1, px1) - fminf(tx0, px0);
        const float ac = eh * ew, dc = ac + 1.0e-12f;
        float de[4];  // d C / d p_k: an edge of the enclosing box moves with the prediction where the prediction is the outer one
        de[0] = ew * (py0 < ty0 ? 1.0f : 0.f);
        de[1] = ew * (py1 > ty1 ? 1.0f : 0.f);
        de[2] = eh * (px0 < tx0 ? 1.0f : 0.f);
        de[3] = eh * (px1 > tx1 ? 1.0f : 0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float d_uni = da[q] - di[q];
            const float d_iou = (di[q] * den - inter * d_uni) / (den * den);
            const float d_gap = ((de[q] - d_uni) * dc - (ac - uni) * de[q]) / (dc * dc);
            g[q] = -(d_iou - d_gap);
        }
        return;
    }
    const float k = -1.0f / (iou + 1.0e-12f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float d_uni = da[q] - di[q];
        g[q] = k * (di[q] * den - inter * d_uni) / (den * den);
    }
}

// class-channel term / derivative for either classification loss
__device__ __forceinline__ float cls_term(const LossSpec& sp, float y, float x) {
    if (sp.cls_mode == 0) return focal_term(y, x, sp.alpha, sp.gamma);
    return fmaxf(x, 0.f) - x * y + __logf(1.0f + __expf(-fabsf(x)));  // max(x, 0) - x z + log(1 + exp(-|x|))
}
__device__ __forceinline__ float cls_grad(const LossSpec& sp, float y, float x) {
    if (sp.cls_mode == 0) return focal_grad(y, x, sp.alpha, sp.gamma);
    return sigmoid_f(x) - y;
}

struct LossAcc {
    float cls, reg, cen;
    int npos;
};

// One element of a row whose targets are known (`y`), `m` = positive-row weight (0, 1 or the caller's mask).
__device__ __forceinline__ void accumulate_element(const LossSpec& sp, int cls0, int c, float x, float y, float m,
                                                   LossAcc& acc) {
    if (c >= cls0) {
        acc.cls += cls_term(sp, y, x);
    } else if (c < sp.reg_ch) {
        if (sp.reg_mode == 0 && m != 0.f) acc.reg += m * smooth_l1_term(y, x, sp.delta);
    } else if (sp.cen_mode == 1) {
        acc.cen += smooth_l1_term(y, sigmoid_f(x), sp.delta);
    } else if (sp.cen_mode == 2) {
        acc.cen += focal_term(y, x, sp.alpha, sp.gamma);
    }
}

template <class P, bool kFused, bool kGrad = false>
__global__ void __launch_bounds__(DH_THREADS) loss_kernel(const __grid_constant__ LossArgs<P> ga) {
    extern __shared__ __align__(128) unsigned char smem[];
    const LossSmemLayout lay = loss_smem_layout<P, kFused>(ga.tile_buf_bytes, ga.tt.rows_per_tile, ga.box_cap);
    const LossArgs<P>& a = *reinterpret_cast<const LossArgs<P>*>(smem + lay.args_off);  // see encode_kernel
    copy_args_to_smem(ga, reinterpret_cast<LossArgs<P>*>(smem + lay.args_off));
    float* st = reinterpret_cast<float*>(smem + lay.tgt_off);  // fused: the on-the-fly target tile
    int* rowpos = reinterpret_cast<int*>(smem + lay.rowpos_off);
    typename P::Rec* recs = reinterpret_cast<typename P::Rec*>(smem + lay.rec_off);
    float* raw = reinterpret_cast<float*>(smem + lay.raw_off);
    unsigned short* cand = reinterpret_cast<unsigned short*>(smem + lay.cand_off);
    unsigned short* cand_dense = reinterpret_cast<unsigned short*>(smem + lay.cand2_off);
    uint64_t* boxbar = reinterpret_cast<uint64_t*>(smem + lay.misc_off);
    long long* next_chunk = reinterpret_cast<long long*>(smem + lay.misc_off + 8);
    int* wcnt = reinterpret_cast<int*>(smem + lay.misc_off + 32);      // [2][8]
    float* wred = reinterpret_cast<float*>(smem + lay.misc_off + 96);  // [8][4]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ch = ga.tt.ch;
    const LossSpec sp = ga.spec;
    const int cls0 = sp.reg_ch + (sp.cen_mode != 0 ? 1 : 0);
    const bool vec = ga.allow_vec && (ch & 3) == 0 && sp.cen_mode == 0;  // rows are whole float4s: [regs][classes...]
    const int vpr = ch >> 2;                            // float4s per row (vector path)
    const FastDiv div_ch = make_fastdiv(static_cast<uint32_t>(ch));
    const FastDiv div_vpr = make_fastdiv(static_cast<uint32_t>(vpr > 0 ? vpr : 1));
    const float oma = 1.0f - sp.alpha;
    const int tpi = ga.tt.tiles_per_image;
    const long long n_chunks = static_cast<long long>(ga.tt.batch) * ga.chunks_per_image;
    long long chunk = blockIdx.x;
    if (chunk >= n_chunks) {  // (the launchers clamp the grid to the chunk count; kept for safety)
        if (threadIdx.x == 0) sched_release(ga.sched);
        return;
    }

    {  // shared-memory state that must start at zero
        float4* z = reinterpret_cast<float4*>(smem + lay.tgt_off);
        const int n4 = (kFused ? ga.tile_buf_bytes : 0) / 16;
        for (int e = tid; e < n4; e += DH_THREADS) z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = tid; r < ga.tt.rows_per_tile; r += DH_THREADS) rowpos[r] = 0;
    }
    if (tid == 0) {
        mbar_init(boxbar, 1);
        mbar_init_fence();
    }
    __syncthreads();

    uint32_t box_parity = 0;
    int cur_img = -1, n_boxes = 0;
    int it = 0;
    for (; chunk < n_chunks;) {
        const long long img = chunk / ga.chunks_per_image;
        const int sub = static_cast<int>(chunk - img * ga.chunks_per_image);
        const long long t_begin = img * tpi + static_cast<long long>(sub) * ga.chunk_tiles;
        const long long t_end = min(t_begin + ga.chunk_tiles, (img + 1) * tpi);
        if (tid == 0) *next_chunk = static_cast<long long>(atomicAdd(ga.sched, 1u)) + gridDim.x;
        LossAcc acc = {0.f, 0.f, 0.f, 0};
        TileCursor cur;
        cursor_init(a.tt, t_begin, cur);
        for (long long tile = t_begin; tile < t_end; ++tile, ++it, cursor_next(a.tt, cur)) {
            TileInfo ti;
            cursor_info(a.tt, cur, ti);
            const MapDesc& md = a.tt.maps[ti.m];
            const long long off = static_cast<long long>(ti.b) * md.image_stride + static_cast<long long>(ti.r0) * ch;
            const float* __restrict__ gp = md.pred + off;
            const float* __restrict__ gt = kFused ? nullptr : md.out + off;
            float* __restrict__ gg = (kGrad && !kFused && a.grad_maps[ti.m]) ? a.grad_maps[ti.m] + off : nullptr;
            const int nfl = ti.nrows * ch;
            int ncand = 0;
            uint32_t dmask = 0u;

            if constexpr (kFused) {
                if (ti.b != cur_img) {
                    __syncthreads();
                    n_boxes = stage_boxes(a.boxes, a.nbox, ti.b, a.max_boxes, a.box_cap, raw, boxbar, box_parity);
                    const float hi = a.img_dim[2 * ti.b], wi = a.img_dim[2 * ti.b + 1];
                    if (tid < n_boxes) P::make_record(a.pp, raw + 5 * tid, hi, wi, tid, recs[tid]);
                    __syncthreads();
                    if (tile == static_cast<long long>(ti.b) * tpi) P::image_prologue(a.pp, recs, n_boxes, ti.b);
                    cur_img = ti.b;
                }
                unsigned short* cseg = cand + (it & 1) * DH_THREADS;  // parity double-buffer: one barrier per tile
                int* wc = wcnt + (it & 1) * 8;
                build_candidates<P>(a.pp, recs, n_boxes, ti, md, cseg, wc, warp, lane);
                __syncthreads();  // barrier A
                ncand = total_candidates(wc);
                if (ncand > 0) {  // uniform: materialise this tile's targets in shared memory
                    int base = 0;
                    for (int w = 0; w < warp; ++w) base += wc[w];
                    if (lane < wc[warp]) cand_dense[base + lane] = cseg[warp * 32 + lane];
                    __syncthreads();
                    int painted = 0;
                    for (int k = 0, r = tid; r < ti.nrows; r += DH_THREADS, ++k) {
                        const int n = P::emit_row(a.pp, ti, md, ti.r0 + r, st + r * ch, recs, cand_dense, ncand);
                        if (n > 0) dmask |= (1u << k), rowpos[r] = 1, ++acc.npos;
                        painted += n;
                    }
                    P::tile_epilogue(a.pp, ti, painted);
                    __syncthreads();
                }
            } else {
                // unfused: positive rows come from the materialised targets (or the caller's mask)
                if (sp.reg_ch > 0) {
                    if (sp.pos_rule == 2) {
                        const float* gm = a.mask_maps[ti.m] + static_cast<long long>(ti.b) * md.rows + ti.r0;
                        for (int r = tid; r < ti.nrows; r += DH_THREADS) rowpos[r] = __float_as_int(gm[r]);
                    } else if (vec) {
                        const float4* __restrict__ gt4 = reinterpret_cast<const float4*>(gt);
                        for (int q = tid; q < ti.nrows * vpr; q += DH_THREADS) {
                            const int r = static_cast<int>(fdiv_u32(q, div_vpr));
                            if ((q - r * vpr) * 4 >= cls0) {
                                const float4 y = __ldg(gt4 + q);
                                const float mx = fmaxf(fmaxf(y.x, y.y), fmaxf(y.z, y.w));
                                if (sp.pos_rule == 0 ? (mx >= 1.0f) : (mx > 0.0f)) rowpos[r] = 1;
                            }
                        }
                    } else {
                        for (int e = tid; e < nfl; e += DH_THREADS) {
                            const int r = static_cast<int>(fdiv_u32(e, div_ch));
                            if (e - r * ch >= cls0) {
                                const float y = __ldg(gt + e);
                                if (sp.pos_rule == 0 ? (y >= 1.0f) : (y > 0.0f)) rowpos[r] = 1;
                            }
                        }
                    }
                    __syncthreads();
                }
                ncand = 1;  // targets always have to be looked at
            }

            // ---- element pass --------------------------------------------------------------------
            if (vec) {
                const int nvec = ti.nrows * vpr;
                const float4* __restrict__ gp4 = reinterpret_cast<const float4*>(gp);
                const float4* __restrict__ gt4 = reinterpret_cast<const float4*>(gt);
                const float4* st4 = reinterpret_cast<const float4*>(st);
                for (int q0 = tid; q0 < nvec; q0 += 4 * DH_THREADS) {
                    float4 x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int q = q0 + u * DH_THREADS;
                        x[u] = q < nvec ? __ldcs(gp4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int q = q0 + u * DH_THREADS;
                        if (q >= nvec) break;
                        const int r = static_cast<int>(fdiv_u32(q, div_vpr));
                        const int c4 = q - r * vpr;
                        const bool is_reg = sp.reg_ch > 0 && c4 == 0;
                        if (kFused && (ncand == 0 || !rowpos[r])) {  // label-free fast path
                            if (!is_reg) {
                                if (sp.cls_mode == 0)
                                    acc.cls += focal_neg(x[u].x, oma, sp.gamma) + focal_neg(x[u].y, oma, sp.gamma) +
                                               focal_neg(x[u].z, oma, sp.gamma) + focal_neg(x[u].w, oma, sp.gamma);
                                else
                                    acc.cls += cls_term(sp, 0.f, x[u].x) + cls_term(sp, 0.f, x[u].y) + cls_term(sp, 0.f, x[u].z) + cls_term(sp, 0.f, x[u].w);
                            }
                            continue;
                        }
                        const float4 y = kFused ? st4[q] : __ldcs(gt4 + q);
                        float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (is_reg) {
                            const float m = sp.pos_rule == 2 ? __int_as_float(rowpos[r]) : (rowpos[r] ? 1.0f : 0.0f);
                            if (m != 0.f) {
                                if (!kFused) ++acc.npos;
                                if (sp.reg_mode == 0) {
                                    acc.reg += m * (smooth_l1_term(y.x, x[u].x, sp.delta) + smooth_l1_term(y.y, x[u].y, sp.delta) +
                                                    smooth_l1_term(y.z, x[u].z, sp.delta) + smooth_l1_term(y.w, x[u].w, sp.delta));
                                    if (kGrad && gg) {
                                        const float k = m * sp.w_reg;
                                        gv = make_float4(k * smooth_l1_grad(y.x, x[u].x, sp.delta), k * smooth_l1_grad(y.y, x[u].y, sp.delta),
                                                         k * smooth_l1_grad(y.z, x[u].z, sp.delta), k * smooth_l1_grad(y.w, x[u].w, sp.delta));
                                    }
                                } else {
                                    const int row = ti.r0 + r;
                                    const int cell = static_cast<int>(fdiv_u32(row, md.div_sub));
                                    const int i = static_cast<int>(fdiv_u32(cell, md.div_width));
                                    const float tv[4] = {y.x, y.y, y.z, y.w}, pv[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
                                    acc.reg += m * iou_loss_term(tv, pv, static_cast<float>(i), static_cast<float>(cell - i * md.width), sp.reg_mode);
                                    if (kGrad && gg) {
                                        float g4[4];
                                        iou_loss_grad(tv, pv, static_cast<float>(i), static_cast<float>(cell - i * md.width), g4, sp.reg_mode);
                                        const float k = m * sp.w_reg;
                                        gv = make_float4(k * g4[0], k * g4[1], k * g4[2], k * g4[3]);
                                    }
                                }
                            }
                        } else {
                            acc.cls += cls_term(sp, y.x, x[u].x) + cls_term(sp, y.y, x[u].y) + cls_term(sp, y.z, x[u].z) + cls_term(sp, y.w, x[u].w);
                            if (kGrad && gg)
                                gv = make_float4(sp.w_cls * cls_grad(sp, y.x, x[u].x), sp.w_cls * cls_grad(sp, y.y, x[u].y),
                                                 sp.w_cls * cls_grad(sp, y.z, x[u].z), sp.w_cls * cls_grad(sp, y.w, x[u].w));
                        }
                        if (kGrad && gg) __stcs(reinterpret_cast<float4*>(gg) + q, gv);
                    }
                }
            } else {
                for (int e0 = tid; e0 < nfl; e0 += 4 * DH_THREADS) {
                    float x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = e0 + u * DH_THREADS;
                        x[u] = e < nfl ? __ldcs(gp + e) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = e0 + u * DH_THREADS;
                        if (e >= nfl) break;
                        const int r = static_cast<int>(fdiv_u32(e, div_ch));
                        const int c = e - r * ch;
                        const bool pos_row = rowpos[r] != 0;
                        float y = 0.f;
                        if (!kFused) y = __ldcs(gt + e);
                        else if (ncand > 0 && pos_row) y = st[e];
                        const float m = sp.pos_rule == 2 ? __int_as_float(rowpos[r]) : (pos_row ? 1.0f : 0.0f);
                        if (c == 0 && sp.reg_ch > 0 && m != 0.f) {
                            if (!kFused) ++acc.npos;
                            if (sp.reg_mode != 0) {
                                const int row = ti.r0 + r;
                                const int cell = static_cast<int>(fdiv_u32(row, md.div_sub));
                                const int i = static_cast<int>(fdiv_u32(cell, md.div_width));
                                float tv[4], pv[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    pv[k] = gp[e + k];
                                    tv[k] = kFused ? st[e + k] : gt[e + k];
                                }
                                acc.reg += m * iou_loss_term(tv, pv, static_cast<float>(i), static_cast<float>(cell - i * md.width), sp.reg_mode);
                            }
                        }
                        accumulate_element(sp, cls0, c, x[u], y, m, acc);
                        if (kGrad && gg) {
                            float g = 0.f;
                            if (c >= cls0) {
                                g = sp.w_cls * cls_grad(sp, y, x[u]);
                            } else if (c < sp.reg_ch) {
                                if (m != 0.f) {
                                    if (sp.reg_mode == 0) {
                                        g = m * sp.w_reg * smooth_l1_grad(y, x[u], sp.delta);
                                    } else {  // each of the row's four regression elements recomputes the row's IoU gradient
                                        const int row = ti.r0 + r;
                                        const int cell = static_cast<int>(fdiv_u32(row, md.div_sub));
                                        const int i = static_cast<int>(fdiv_u32(cell, md.div_width));
                                        float tv[4], pv[4], g4[4];
#pragma unroll
                                        for (int k = 0; k < 4; ++k) pv[k] = gp[e - c + k], tv[k] = gt[e - c + k];
                                        iou_loss_grad(tv, pv, static_cast<float>(i), static_cast<float>(cell - i * md.width), g4, sp.reg_mode);
                                        g = m * sp.w_reg * (c == 0 ? g4[0] : (c == 1 ? g4[1] : (c == 2 ? g4[2] : g4[3])));
                                    }
                                }
                            } else if (sp.cen_mode == 1) {
                                g = sp.w_cen * cen_l1_grad(y, x[u], sp.delta);
                            } else if (sp.cen_mode == 2) {
                                g = sp.w_cen * focal_grad(y, x[u], sp.alpha, sp.gamma);
                            }
                            __stcs(gg + e, g);
                        }
                    }
                }
            }

            if constexpr (kFused) {
                if (ncand > 0) {  // uniform: re-zero the rows this thread emitted once everybody has read them
                    __syncthreads();
                    for (int k = 0; dmask; ++k, dmask >>= 1)
                        if (dmask & 1u) {
                            const int r = tid + k * DH_THREADS;
                            float* row = st + r * ch;
                            for (int c = 0; c < ch; ++c) row[c] = 0.f;
                            rowpos[r] = 0;
                        }
                }
            } else if (sp.reg_ch > 0) {
                __syncthreads();  // everybody has consumed rowpos
                for (int r = tid; r < ti.nrows; r += DH_THREADS) rowpos[r] = 0;
            }
        }
        // ---- per-chunk reduction -> partials[chunk] ----------------------------------------------------
        acc.cls = warp_sum(acc.cls), acc.reg = warp_sum(acc.reg), acc.cen = warp_sum(acc.cen);
        acc.npos = warp_sum_i(acc.npos);
        if (lane == 0) {
            wred[warp * 4 + 0] = acc.cls, wred[warp * 4 + 1] = acc.reg, wred[warp * 4 + 2] = acc.cen;
            wred[warp * 4 + 3] = static_cast<float>(acc.npos);
        }
        __syncthreads();  // also: the prefetched chunk id is visible
        if (tid < 4) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < DH_THREADS / 32; ++w) v += wred[w * 4 + tid];
            a.partials[chunk * 4 + tid] = v;
        }
        chunk = *next_chunk;
        __syncthreads();  // wred / next_chunk are free again
    }
    if (tid == 0) sched_release(ga.sched);
}

// partials [B * chunks_per_image, 4] -> per_image [B, 4]; one CTA per image, fixed summation order
__global__ void loss_finalize_images(const float* __restrict__ partials, int chunks_per_image, float* __restrict__ per_image) {
    __shared__ double red[128][4];
    const int b = blockIdx.x, tid = threadIdx.x;
    double acc[4] = {0, 0, 0, 0};
    const float4* p = reinterpret_cast<const float4*>(partials) + static_cast<long long>(b) * chunks_per_image;
    for (int t = tid; t < chunks_per_image; t += 128) {
        const float4 v = p[t];
        acc[0] += v.x, acc[1] += v.y, acc[2] += v.z, acc[3] += v.w;
    }
    for (int k = 0; k < 4; ++k) red[tid][k] = acc[k];
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (tid < o)
            for (int k = 0; k < 4; ++k) red[tid][k] += red[tid + o][k];
        __syncthreads();
    }
    if (tid < 4) per_image[b * 4 + tid] = static_cast<float>(red[0][tid]);
}

// per_image [B, 4] -> total [4]
__global__ void loss_finalize_total(const float* __restrict__ per_image, int batch, float* __restrict__ total) {
    __shared__ double red[256][4];
    const int tid = threadIdx.x;
    double acc[4] = {0, 0, 0, 0};
    const float4* p = reinterpret_cast<const float4*>(per_image);
    for (int b = tid; b < batch; b += 256) {
        const float4 v = p[b];
        acc[0] += v.x, acc[1] += v.y, acc[2] += v.z, acc[3] += v.w;
    }
    for (int k = 0; k < 4; ++k) red[tid][k] = acc[k];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o)
            for (int k = 0; k < 4; ++k) red[tid][k] += red[tid + o][k];
        __syncthreads();
    }
    if (tid < 4) total[tid] = static_cast<float>(red[0][tid]);
}

}  // namespace dh
