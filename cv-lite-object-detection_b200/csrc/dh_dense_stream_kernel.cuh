// Losses over MATERIALISED targets (format_data -> model_loss, FCOS/fcos.py:464-496 and its copies) in the same
// "stream + row pass" shape as the fused kernel: a warp owns 32-row tiles, reads predictions and targets once with
// coalesced streaming loads, takes the label-free fast path wherever the four targets of a load are all zero (> 99 % of
// the class channels) and the exact per-element term otherwise; rows whose class targets mark them positive are recorded
// in a per-warp bit mask, and after the stream the owner lane of each positive row adds the box-regression loss (and
// overwrites the row's regression gradient).  No block barrier inside a chunk, deterministic per-chunk partials.
#pragma once
#include "dh_fused_loss_kernel.cuh"

namespace dh {

// one class-channel element with a known target
template <int kCls, bool kNewton, bool kGrad>
__device__ __forceinline__ float labelled_term(const LossSpec& sp, float y, float x, float gscale, float& fast_acc, float& slow_acc) {
    if (y == 0.f) return stream_term<kCls, kNewton, kGrad>(x, sp.gamma, fast_acc) * gscale;
    slow_acc += cls_term(sp, y, x);
    return kGrad ? sp.w_cls * cls_grad(sp, y, x) : 0.f;
}

struct DenseAcc {
    StreamAcc s;    // label-free class terms (log2 units) + centerness
    float cls_nat;  // labelled class terms, natural units
};

template <int kCls, bool kGrad, int U>
__device__ __forceinline__ void dense_vec(const LossSpec& sp, const float* __restrict__ p, const float* __restrict__ t, float* __restrict__ gout,
                                          int nrows, int vpr, int c4, int step, int lane, float gscale, const FastDiv& div_vpr,
                                          unsigned* pos_bits, DenseAcc& a) {
    const float4* __restrict__ pp = reinterpret_cast<const float4*>(p) + lane;
    const float4* __restrict__ tp = reinterpret_cast<const float4*>(t) + lane;
    float4* __restrict__ gp = reinterpret_cast<float4*>(gout) + lane;
    const int n_mine = (nrows * vpr - lane + 31) >> 5;
    auto item = [&](const float4& x, const float4& y, int k) {
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 != 0) {
            if (y.x == 0.f && y.y == 0.f && y.z == 0.f && y.w == 0.f) {
                if (kCls == 1) {
                    stream_pair_g2<kGrad>(x.x, x.y, a.s.p0, d.x, d.y);
                    stream_pair_g2<kGrad>(x.z, x.w, a.s.p1, d.z, d.w);
                } else {
                    d.x = stream_term<kCls, false, kGrad>(x.x, sp.gamma, a.s.c0);
                    d.y = stream_term<kCls, true, kGrad>(x.y, sp.gamma, a.s.c1);
                    d.z = stream_term<kCls, true, kGrad>(x.z, sp.gamma, a.s.c2);
                    d.w = stream_term<kCls, false, kGrad>(x.w, sp.gamma, a.s.c3);
                }
                d.x *= gscale, d.y *= gscale, d.z *= gscale, d.w *= gscale;
            } else {  // rare: a labelled class channel
                d.x = labelled_term<kCls, false, kGrad>(sp, y.x, x.x, gscale, a.s.c0, a.cls_nat);
                d.y = labelled_term<kCls, false, kGrad>(sp, y.y, x.y, gscale, a.s.c1, a.cls_nat);
                d.z = labelled_term<kCls, false, kGrad>(sp, y.z, x.z, gscale, a.s.c2, a.cls_nat);
                d.w = labelled_term<kCls, false, kGrad>(sp, y.w, x.w, gscale, a.s.c3, a.cls_nat);
                const float mx = fmaxf(fmaxf(y.x, y.y), fmaxf(y.z, y.w));
                if (sp.pos_rule == 0 ? (mx >= 1.0f) : (mx > 0.0f))
                    atomicOr(pos_bits, 1u << fdiv_u32(static_cast<uint32_t>(lane + 32 * k), div_vpr));
            }
        }
        if (kGrad) __stcs(gp + 32 * k, d);
        c4 += step;
        if (c4 >= vpr) c4 -= vpr;
    };
    int k = 0;
#pragma unroll 1
    for (; k + U <= n_mine; k += U) {
        float4 x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = __ldcs(pp + 32 * (k + u)), y[u] = __ldcs(tp + 32 * (k + u));
#pragma unroll
        for (int u = 0; u < U; ++u) item(x[u], y[u], k + u);
    }
#pragma unroll 1
    for (; k < n_mine; ++k) item(__ldcs(pp + 32 * k), __ldcs(tp + 32 * k), k);
}

template <int kCls, bool kGrad, int U>
__device__ __forceinline__ void dense_scalar(const LossSpec& sp, const float* __restrict__ p, const float* __restrict__ t,
                                             float* __restrict__ gout, int nrows, int ch, int c, int step, int lane, float gscale,
                                             const FastDiv& div_ch, unsigned* pos_bits, DenseAcc& a) {
    const int cls0 = sp.reg_ch + (sp.cen_mode != 0 ? 1 : 0);
    const int n_mine = (nrows * ch - lane + 31) >> 5;
    auto item = [&](float x, float y, int k) {
        float d = 0.f;
        if (c >= cls0) {
            d = labelled_term<kCls, false, kGrad>(sp, y, x, gscale, a.s.c0, a.cls_nat);
            if (y != 0.f && (sp.pos_rule == 0 ? (y >= 1.0f) : (y > 0.0f)))
                atomicOr(pos_bits, 1u << fdiv_u32(static_cast<uint32_t>(lane + 32 * k), div_ch));
        } else if (c >= sp.reg_ch) {  // the centerness channel (cen_mode != 0)
            if (sp.cen_mode == 1) {
                a.s.cen += smooth_l1_term(y, sigmoid_f(x), sp.delta);
                if (kGrad) d = sp.w_cen * cen_l1_grad(y, x, sp.delta);
            } else if (sp.cen_mode == 2) {
                a.s.cen += focal_term(y, x, sp.alpha, sp.gamma);
                if (kGrad) d = sp.w_cen * focal_grad(y, x, sp.alpha, sp.gamma);
            }
        }
        if (kGrad) __stcs(gout + lane + 32 * k, d);
        c += step;
        if (c >= ch) c -= ch;
    };
    int k = 0;
#pragma unroll 1
    for (; k + U <= n_mine; k += U) {
        float x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = __ldcs(p + lane + 32 * (k + u)), y[u] = __ldcs(t + lane + 32 * (k + u));
#pragma unroll
        for (int u = 0; u < U; ++u) item(x[u], y[u], k + u);
    }
#pragma unroll 1
    for (; k < n_mine; ++k) item(__ldcs(p + lane + 32 * k), __ldcs(t + lane + 32 * k), k);
}

// the owner lane of a positive row: box-regression loss (weight m) and its gradient
__device__ __noinline__ float dense_row_reg(const LossSpec& sp, const float* __restrict__ prow, const float* __restrict__ trow,
                                            float* __restrict__ grow, float m, float gy, float gx) {
    float x[4], y[4], acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = prow[k], y[k] = trow[k];
    if (sp.reg_mode == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += smooth_l1_term(y[k], x[k], sp.delta);
        if (grow) {
#pragma unroll
            for (int k = 0; k < 4; ++k) grow[k] = m * sp.w_reg * smooth_l1_grad(y[k], x[k], sp.delta);
        }
    } else {
        acc = iou_loss_term(y, x, gy, gx, sp.reg_mode);
        if (grow) {
            float g[4];
            iou_loss_grad(y, x, gy, gx, g, sp.reg_mode);
#pragma unroll
            for (int k = 0; k < 4; ++k) grow[k] = m * sp.w_reg * g[k];
        }
    }
    return m * acc;
}

template <int kCls, bool kGrad>
__global__ void __launch_bounds__(DH_THREADS, 4) dense_stream_kernel(const __grid_constant__ LossArgs<NoPolicy> ga) {
    extern __shared__ __align__(128) unsigned char smem[];
    LossArgs<NoPolicy>* sa_ptr = reinterpret_cast<LossArgs<NoPolicy>*>(smem);
    const LossArgs<NoPolicy>& a = *sa_ptr;
    copy_args_to_smem(ga, sa_ptr);
    __shared__ long long next_chunk;
    __shared__ float wred[DH_THREADS / 32][4];
    __shared__ unsigned pos_bits[DH_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ch = ga.tt.ch;
    const LossSpec sp = ga.spec;
    const bool vec = ga.allow_vec && (ch & 3) == 0 && sp.cen_mode == 0 && sp.reg_ch == 4;
    const int per = vec ? (ch >> 2) : ch;
    const int step = 32 % per, c_lane = lane % per;
    const FastDiv div_per = make_fastdiv(static_cast<uint32_t>(per));
    const float gscale = (kCls == 2 ? 1.0f : 1.0f - sp.alpha) * sp.w_cls;
    const int tpi = ga.tt.tiles_per_image;
    const long long n_chunks = static_cast<long long>(ga.tt.batch) * ga.chunks_per_image;
    long long chunk = blockIdx.x;
    __syncthreads();
    if (chunk >= n_chunks) {
        if (tid == 0) sched_release(ga.sched);
        return;
    }
#pragma unroll 1
    for (; chunk < n_chunks;) {
        const int img = static_cast<int>(chunk / ga.chunks_per_image);
        const int sub = static_cast<int>(chunk - static_cast<long long>(img) * ga.chunks_per_image);
        const int t_begin = sub * ga.chunk_tiles, t_end = min(t_begin + ga.chunk_tiles, tpi);
        if (tid == 0) next_chunk = static_cast<long long>(atomicAdd(ga.sched, 1u)) + gridDim.x;
        DenseAcc acc = {{0.f, 0.f, 0.f, 0.f, 0.f, 0ull, 0ull}, 0.f};
        float reg = 0.f;
        int npos = 0;
        TileCursor cur;
        cursor_init(a.tt, static_cast<long long>(img) * tpi + t_begin, cur);
#pragma unroll 1
        for (int tile = t_begin; tile < t_end; ++tile, cursor_next(a.tt, cur)) {
            TileInfo ti;
            cursor_info(a.tt, cur, ti);
            const int nrw = min(32, ti.nrows - 32 * warp);
            if (nrw <= 0) continue;
            const int r0 = ti.r0 + 32 * warp;
            const MapDesc& md = a.tt.maps[ti.m];
            const long long off = static_cast<long long>(img) * md.image_stride + static_cast<long long>(r0) * ch;
            const float* __restrict__ gp = md.pred + off;
            const float* __restrict__ gt = md.out + off;
            float* __restrict__ gg = (kGrad && a.grad_maps[ti.m]) ? a.grad_maps[ti.m] + off : nullptr;
            if (lane == 0) pos_bits[warp] = 0u;
            __syncwarp();
            if (vec) dense_vec<kCls, kGrad, 4>(sp, gp, gt, gg, nrw, per, c_lane, step, lane, gscale, div_per, &pos_bits[warp], acc);
            else dense_scalar<kCls, kGrad, 4>(sp, gp, gt, gg, nrw, ch, c_lane, step, lane, gscale, div_per, &pos_bits[warp], acc);
            __syncwarp();
            if (sp.reg_ch > 0 && lane < nrw) {
                float m = 0.f;
                if (sp.pos_rule == 2) m = a.mask_maps[ti.m][static_cast<long long>(img) * md.rows + r0 + lane];
                else m = ((pos_bits[warp] >> lane) & 1u) ? 1.0f : 0.f;
                if (m != 0.f) {
                    const int row = r0 + lane;
                    const int cell = static_cast<int>(fdiv_u32(row, md.div_sub));
                    const int i = static_cast<int>(fdiv_u32(cell, md.div_width));
                    reg += dense_row_reg(sp, gp + lane * ch, gt + lane * ch, gg ? gg + lane * ch : nullptr, m, static_cast<float>(i),
                                         static_cast<float>(cell - i * md.width));
                    ++npos;
                }
            }
            __syncwarp();
        }
        float q0, q1, q2, q3;
        unpack2(acc.s.p0, q0, q1);
        unpack2(acc.s.p1, q2, q3);
        float cls = (((acc.s.c0 + acc.s.c1) + (acc.s.c2 + acc.s.c3)) + ((q0 + q1) + (q2 + q3))) * ((kCls == 2 ? 1.0f : 1.0f - sp.alpha) * kLn2) + acc.cls_nat;
        cls = warp_sum(cls);
        reg = warp_sum(reg);
        const float cen = warp_sum(acc.s.cen);
        npos = warp_sum_i(npos);
        if (lane == 0) wred[warp][0] = cls, wred[warp][1] = reg, wred[warp][2] = cen, wred[warp][3] = static_cast<float>(npos);
        __syncthreads();
        if (tid < 4) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < DH_THREADS / 32; ++w) v += wred[w][tid];
            a.partials[chunk * 4 + tid] = v;
        }
        chunk = next_chunk;
        __syncthreads();
    }
    if (tid == 0) sched_release(ga.sched);
}

}  // namespace dh
