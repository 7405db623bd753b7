// Dense-head losses as a tile streamer: focal (classes / centerness), smooth-L1 or -log(IoU) (boxes),
// smooth-L1 of sigmoid (centerness) over (prediction, target) tiles.
//
// Predictions always arrive through the TMA: one elected thread issues a 1-D bulk load per tile
// into a two-stage shared-memory ring (mbarrier complete_tx), so the next tile is in flight while
// the CTA does the transcendental work on the current one.  Targets come either from HBM the same
// way (unfused: `format_data` output -> `model_loss`) or are produced on the fly by the encoder
// policy into a zeroed shared-memory tile (fused encode+loss: targets never touch HBM, algorithmic
// bytes drop from 3x to 1x the map size).
// Per tile the CTA reduces {cls, reg, cen, n_pos} and writes one partial; a second tiny kernel sums
// the partials per image in a fixed order (deterministic, float64 accumulation).
//
// Reference formulas: FCOS/fcos.py:380-496 (identical copies in the other modules).
#pragma once
#include "dh_policies.cuh"

namespace dh {

struct LossSpec {
    int reg_ch;    // 0 or 4: channels [0, reg_ch) are box regression
    int cen_mode;  // 0 no centerness channel, 1 smooth-L1(sigmoid(pred)) over all rows, 2 focal, 3 present but unused
    int reg_mode;  // 0 smooth-L1, 1 -log(IoU) on the integer grid
    int pos_rule;  // 0 max(class) >= 1, 1 max(class) > 0, 2 external per-row mask
    float alpha, gamma, delta;
};

struct NoPolicy {
    struct Params {
        int unused;
    };
    struct Rec {
        int unused;
    };
};

template <class P>
struct LossArgs {
    TileTable tt;  // maps[m].pred = predictions, maps[m].out = targets (unfused), maps[m].mask optional
    typename P::Params pp;
    LossSpec spec;
    const float* boxes;
    const int* nbox;
    const float* img_dim;
    int max_boxes;
    int box_cap;  // shared-memory capacity in boxes (fused): max_boxes rounded up to 32
    int tile_buf_bytes;
    float* partials;  // [batch * tiles_per_image, 4]
    const float* mask_maps[DH_MAX_MAPS];
};

struct LossSmemLayout {
    int pred_off, tgt_off, rowpos_off, rec_off, raw_off, cand_off, misc_off, args_off, total;
};
template <class P, bool kFused>
__host__ __device__ inline LossSmemLayout loss_smem_layout(int tile_buf_bytes, int rows_per_tile, int box_cap) {
    LossSmemLayout l;
    l.pred_off = 0;
    l.tgt_off = 2 * tile_buf_bytes;
    l.rowpos_off = l.tgt_off + (kFused ? 1 : 2) * tile_buf_bytes;
    l.rec_off = l.rowpos_off + ((rows_per_tile * 4 + 127) & ~127);
    l.raw_off = l.rec_off + (kFused ? ((static_cast<int>(sizeof(typename P::Rec)) * box_cap + 127) & ~127) : 0);
    l.cand_off = l.raw_off + (kFused ? ((box_cap * 20 + 127) & ~127) : 0);
    l.misc_off = l.cand_off + (kFused ? ((box_cap * 2 + 127) & ~127) : 0);
    l.args_off = l.misc_off + 256;
    l.total = l.args_off + ((static_cast<int>(sizeof(LossArgs<P>)) + 127) & ~127);
    return l;
}

// ---- element formulas (float32) ---------------------------------------------------------------
// focal: y*a*(1-s)^g*softplus(-x) + (1-y)*(1-a)*s^g*softplus(x)   (FCOS/fcos.py:443-462, stable form)
__device__ __forceinline__ float focal_term(float y, float x, float alpha, float gamma) {
    const float ax = fabsf(x);
    const float e = __expf(-ax);                 // exp(-|x|)
    const float soft = __logf(1.0f + e);         // log(1 + exp(-|x|)), as the reference writes it
    const float inv = __fdividef(1.0f, 1.0f + e);
    const float s = x >= 0.f ? inv : e * inv;    // sigmoid(x)
    const float om = x >= 0.f ? e * inv : inv;   // 1 - sigmoid(x)
    float p_pos, p_neg;
    if (gamma == 2.0f) {
        p_pos = om * om, p_neg = s * s;
    } else {
        p_pos = __powf(om, gamma), p_neg = __powf(s, gamma);
    }
    const float sp_pos = soft + fmaxf(x, 0.f);   // softplus(x)
    const float sp_neg = soft - fminf(x, 0.f);   // softplus(-x)
    return y * alpha * p_pos * sp_neg + (1.0f - y) * (1.0f - alpha) * p_neg * sp_pos;
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
// -log(IoU) of two tblr boxes anchored at the integer grid point (gx, gy) (FCOS/fcos.py:393-441)
__device__ __forceinline__ float iou_loss_term(const float* t, const float* p, float gy, float gx) {
    const float ty0 = gy - t[0], ty1 = gy + t[1], tx0 = gx - t[2], tx1 = gx + t[3];
    const float py0 = gy - p[0], py1 = gy + p[1], px0 = gx - p[2], px1 = gx + p[3];
    const float ih = fmaxf(0.f, fminf(ty1, py1) - fmaxf(ty0, py0));
    const float iw = fmaxf(0.f, fminf(tx1, px1) - fmaxf(tx0, px0));
    const float inter = iw * ih;
    const float uni = ((ty1 - ty0) * (tx1 - tx0) + (py1 - py0) * (px1 - px0)) - inter;
    const float iou = inter / (uni + 1.0e-12f);
    return -logf(iou + 1.0e-12f);
}

// Bring `nfl` floats starting at `g` into shared memory at `s` (same 16-byte phase as `g`):
// the 16-byte aligned body goes through the TMA (returns its byte count for expect_tx), the ragged
// edges (<= 3 floats each) are copied by the caller's threads with `edge_copy`.
__device__ __forceinline__ uint32_t bulk_body(const float* g, int nfl, int& head, int& body) {
    const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(g) >> 2) & 3u);
    head = min((4 - mis) & 3, nfl);
    body = (nfl - head) & ~3;
    return static_cast<uint32_t>(body) * 4u;
}

template <class P, bool kFused>
__global__ void __launch_bounds__(DH_THREADS) loss_kernel(const __grid_constant__ LossArgs<P> ga) {
    extern __shared__ __align__(128) unsigned char smem[];
    const LossSmemLayout lay = loss_smem_layout<P, kFused>(ga.tile_buf_bytes, ga.tt.rows_per_tile, ga.box_cap);
    const LossArgs<P>& a = *reinterpret_cast<const LossArgs<P>*>(smem + lay.args_off);  // see encode_kernel
    copy_args_to_smem(ga, reinterpret_cast<LossArgs<P>*>(smem + lay.args_off));
    int* rowpos = reinterpret_cast<int*>(smem + lay.rowpos_off);
    typename P::Rec* recs = reinterpret_cast<typename P::Rec*>(smem + lay.rec_off);
    float* raw = reinterpret_cast<float*>(smem + lay.raw_off);
    unsigned short* cand = reinterpret_cast<unsigned short*>(smem + lay.cand_off);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + lay.misc_off);     // [2] tile-landed barriers
    uint64_t* boxbar = reinterpret_cast<uint64_t*>(smem + lay.misc_off + 16);
    int* wcount = reinterpret_cast<int*>(smem + lay.misc_off + 32);        // [8]
    float* wred = reinterpret_cast<float*>(smem + lay.misc_off + 64);      // [8][4]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ch = ga.tt.ch;
    const FastDiv div_ch = make_fastdiv(static_cast<uint32_t>(ch));
    const long long total_tiles = static_cast<long long>(ga.tt.batch) * ga.tt.tiles_per_image;
    const long long t_begin = total_tiles * blockIdx.x / gridDim.x;
    const long long t_end = total_tiles * (blockIdx.x + 1) / gridDim.x;
    if (t_begin >= t_end) return;

    if (kFused) {  // the on-the-fly target tile starts zeroed; owners re-zero what they dirty
        float4* z = reinterpret_cast<float4*>(smem + lay.tgt_off);
        for (int e = tid; e < ga.tile_buf_bytes / 16; e += DH_THREADS) z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(boxbar, 1);
        mbar_init_fence();
    }
    __syncthreads();

    // producer: one thread issues the bulk loads of a tile into stage s
    auto issue = [&](const TileCursor& c, int s) {
        TileInfo ti;
        cursor_info(a.tt, c, ti);
        const MapDesc& md = a.tt.maps[ti.m];
        const long long off = static_cast<long long>(ti.b) * md.image_stride + static_cast<long long>(ti.r0) * ch;
        const int nfl = ti.nrows * ch;
        int head, body;
        const float* gp = md.pred + off;
        uint32_t bytes = bulk_body(gp, nfl, head, body);
        const int misp = static_cast<int>((reinterpret_cast<uintptr_t>(gp) >> 2) & 3u);
        float* sp = reinterpret_cast<float*>(smem + lay.pred_off + s * a.tile_buf_bytes) + misp;
        uint32_t tot = bytes;
        const float* gt = nullptr;
        float* st = nullptr;
        int head_t = 0, body_t = 0;
        if (!kFused) {
            gt = md.out + off;
            tot += bulk_body(gt, nfl, head_t, body_t);
            const int mist = static_cast<int>((reinterpret_cast<uintptr_t>(gt) >> 2) & 3u);
            st = reinterpret_cast<float*>(smem + lay.tgt_off + s * a.tile_buf_bytes) + mist;
        }
        mbar_expect_tx(&full[s], tot);  // tot may be 0: the arrival alone completes the phase
        if (body > 0) bulk_g2s(sp + head, gp + head, static_cast<uint32_t>(body) * 4u, &full[s]);
        if (!kFused && body_t > 0) bulk_g2s(st + head_t, gt + head_t, static_cast<uint32_t>(body_t) * 4u, &full[s]);
    };

    TileCursor cur, nxt;
    cursor_init(a.tt, t_begin, cur);
    nxt = cur;
    if (tid == 0) issue(cur, 0);
    uint32_t parity0 = 0, parity1 = 0, box_parity = 0;
    int cur_img = -1, n_boxes = 0;
    int it = 0;
    for (long long tile = t_begin; tile < t_end; ++tile, ++it, cur = nxt) {
        const int s = it & 1;
        TileInfo ti;
        cursor_info(a.tt, cur, ti);
        const MapDesc& md = a.tt.maps[ti.m];
        cursor_next(a.tt, nxt);
        const long long off = static_cast<long long>(ti.b) * md.image_stride + static_cast<long long>(ti.r0) * ch;
        const int nfl = ti.nrows * ch;
        // stage s^1 was fully consumed before the __syncthreads that ended the previous iteration
        if (tid == 0 && tile + 1 < t_end) issue(nxt, s ^ 1);

        const float* gp = md.pred + off;
        const int misp = static_cast<int>((reinterpret_cast<uintptr_t>(gp) >> 2) & 3u);
        float* sp = reinterpret_cast<float*>(smem + lay.pred_off + s * a.tile_buf_bytes) + misp;
        float* st;
        {   // ragged edges of the prediction (and target) tile: plain loads into the same buffers
            int head, body;
            bulk_body(gp, nfl, head, body);
            if (tid < head) sp[tid] = gp[tid];
            const int tail0 = head + body;
            if (tid >= 32 && tid - 32 < nfl - tail0) sp[tail0 + tid - 32] = gp[tail0 + tid - 32];
        }
        if (!kFused) {
            const float* gt = md.out + off;
            const int mist = static_cast<int>((reinterpret_cast<uintptr_t>(gt) >> 2) & 3u);
            st = reinterpret_cast<float*>(smem + lay.tgt_off + s * a.tile_buf_bytes) + mist;
            int head, body;
            bulk_body(gt, nfl, head, body);
            if (tid >= 64 && tid - 64 < head) st[tid - 64] = gt[tid - 64];
            const int tail0 = head + body;
            if (tid >= 96 && tid - 96 < nfl - tail0) st[tail0 + tid - 96] = gt[tail0 + tid - 96];
        } else {
            st = reinterpret_cast<float*>(smem + lay.tgt_off);
        }

        // ---- targets ------------------------------------------------------------------------
        for (int r = tid; r < ti.nrows; r += DH_THREADS) rowpos[r] = 0;
        uint32_t dmask = 0u;
        if constexpr (kFused) {
            if (ti.b != cur_img) {
                __syncthreads();
                n_boxes = stage_boxes(a.boxes, a.nbox, ti.b, a.max_boxes, a.box_cap, raw, boxbar, box_parity);
                const float hi = a.img_dim[2 * ti.b], wi = a.img_dim[2 * ti.b + 1];
                if (tid < n_boxes) P::make_record(a.pp, raw + 5 * tid, hi, wi, tid, recs[tid]);
                __syncthreads();
                if (tile == static_cast<long long>(ti.b) * a.tt.tiles_per_image) P::image_prologue(a.pp, recs, n_boxes, ti.b);
                cur_img = ti.b;
            }
            const bool hit = tid < n_boxes && P::tile_hit(a.pp, recs[tid], ti, md);
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) wcount[warp] = __popc(bal);
            __syncthreads();
            int base = 0, ncand = 0;
#pragma unroll
            for (int w = 0; w < DH_THREADS / 32; ++w) {
                const int c = wcount[w];
                base += (w < warp) ? c : 0;
                ncand += c;
            }
            if (hit) cand[base + __popc(bal & ((1u << lane) - 1u))] = static_cast<unsigned short>(tid);
            __syncthreads();
            if (ncand > 0) {
                int painted = 0;
                for (int k = 0, r = tid; r < ti.nrows; r += DH_THREADS, ++k) {
                    const int n = P::emit_row(a.pp, ti, md, ti.r0 + r, st + r * ch, recs, cand, ncand);
                    if (n > 0) dmask |= (1u << k), rowpos[r] = 1;
                    painted += n;
                }
                P::tile_epilogue(a.pp, ti, painted);
            }
        }
        // wait for the TMA bytes of this stage
        mbar_wait(&full[s], s ? parity1 : parity0);
        if (s) parity1 ^= 1u; else parity0 ^= 1u;
        __syncthreads();  // edge loads + emitted rows + rowpos zeroing visible

        const int cls0 = a.spec.reg_ch + (a.spec.cen_mode != 0 ? 1 : 0);
        if (!kFused) {  // positive rows from the materialised targets (or the caller's mask)
            if (a.spec.pos_rule == 2) {
                const float* gm = a.mask_maps[ti.m] + static_cast<long long>(ti.b) * md.rows + ti.r0;
                for (int r = tid; r < ti.nrows; r += DH_THREADS) rowpos[r] = __float_as_int(gm[r]);
            } else if (a.spec.reg_ch > 0) {
                for (int e = tid; e < nfl; e += DH_THREADS) {
                    const int r = static_cast<int>(fdiv_u32(e, div_ch));
                    const int c = e - r * ch;
                    if (c >= cls0) {
                        const float y = st[e];
                        if (a.spec.pos_rule == 0 ? (y >= 1.0f) : (y > 0.0f)) rowpos[r] = 1;
                    }
                }
            }
            __syncthreads();
        }

        // ---- element pass -------------------------------------------------------------------
        float acc_cls = 0.f, acc_reg = 0.f, acc_cen = 0.f;
        int npos = 0;
        const float alpha = a.spec.alpha, gamma = a.spec.gamma, delta = a.spec.delta;
        for (int e = tid; e < nfl; e += DH_THREADS) {
            const int r = static_cast<int>(fdiv_u32(e, div_ch));
            const int c = e - r * ch;
            const float x = sp[e], y = st[e];
            if (c >= cls0) {
                acc_cls += focal_term(y, x, alpha, gamma);
            } else if (c < a.spec.reg_ch) {
                float m;
                if (a.spec.pos_rule == 2) {
                    m = __int_as_float(rowpos[r]);
                } else {
                    m = rowpos[r] ? 1.0f : 0.0f;
                }
                if (c == 0 && m != 0.f) ++npos;
                if (m != 0.f) {
                    if (a.spec.reg_mode == 0) {
                        acc_reg += m * smooth_l1_term(y, x, delta);
                    } else if (c == 0) {
                        const int row = ti.r0 + r;
                        const int cell = static_cast<int>(fdiv_u32(row, md.div_sub));
                        const int i = static_cast<int>(fdiv_u32(cell, md.div_width));
                        const int j = cell - i * md.width;
                        acc_reg += m * iou_loss_term(st + e, sp + e, static_cast<float>(i), static_cast<float>(j));
                    }
                }
            } else {  // the centerness channel
                if (a.spec.cen_mode == 1)
                    acc_cen += smooth_l1_term(y, sigmoid_f(x), delta);
                else if (a.spec.cen_mode == 2)
                    acc_cen += focal_term(y, x, alpha, gamma);
            }
        }
        // ---- per-tile reduction -> partials[tile] -------------------------------------------
        acc_cls = warp_sum(acc_cls), acc_reg = warp_sum(acc_reg), acc_cen = warp_sum(acc_cen);
        npos = warp_sum_i(npos);
        if (lane == 0) {
            wred[warp * 4 + 0] = acc_cls, wred[warp * 4 + 1] = acc_reg, wred[warp * 4 + 2] = acc_cen;
            wred[warp * 4 + 3] = static_cast<float>(npos);
        }
        __syncthreads();
        if (tid < 4) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < DH_THREADS / 32; ++w) v += wred[w * 4 + tid];
            a.partials[tile * 4 + tid] = v;
        }
        if constexpr (kFused) {  // re-zero the rows this thread emitted
            for (int k = 0; dmask; ++k, dmask >>= 1)
                if (dmask & 1u) {
                    float* row = st + (tid + k * DH_THREADS) * ch;
                    for (int c = 0; c < ch; ++c) row[c] = 0.f;
                }
        }
        __syncthreads();  // stage s, wred and the target tile are free again
    }
}

// partials [B * tiles_per_image, 4] -> per_image [B, 4]; one CTA per image, fixed summation order
__global__ void loss_finalize_images(const float* __restrict__ partials, int tiles_per_image, float* __restrict__ per_image) {
    __shared__ double red[128][4];
    const int b = blockIdx.x, tid = threadIdx.x;
    double acc[4] = {0, 0, 0, 0};
    const float4* p = reinterpret_cast<const float4*>(partials) + static_cast<long long>(b) * tiles_per_image;
    for (int t = tid; t < tiles_per_image; t += 128) {
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
