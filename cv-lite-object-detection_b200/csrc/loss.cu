// C-ABI launchers for the dense-head losses, unfused (targets in HBM) and fused encode+loss.
#include <cstring>

#include "dh_host.h"
#include "dh_launch.h"
#include "dh_dense_stream_kernel.cuh"

namespace dh {

const CommDev* comm_dev_if_fused(dh_handle_s* h);  // comm.cu

static int check_spec(const LossSpec& s, const char* who) {
    if (!(s.reg_ch == 0 || s.reg_ch == 4)) return set_error(DH_ERR_BAD_ARG, "%s: reg_ch must be 0 or 4", who);
    if (s.cen_mode < 0 || s.cen_mode > 3) return set_error(DH_ERR_BAD_ARG, "%s: cen_mode %d", who, s.cen_mode);
    if (s.reg_mode < 0 || s.reg_mode > 2) return set_error(DH_ERR_BAD_ARG, "%s: reg_mode %d", who, s.reg_mode);
    if (s.pos_rule < 0 || s.pos_rule > 2) return set_error(DH_ERR_BAD_ARG, "%s: pos_rule %d", who, s.pos_rule);
    if (s.cls_mode < 0 || s.cls_mode > 1) return set_error(DH_ERR_BAD_ARG, "%s: cls_mode %d", who, s.cls_mode);
    return DH_OK;
}

template <class P, bool kFused>
static int launch_loss(dh_handle_s* h, LossArgs<P>& a, float* out_per_image, float* out_total, cudaStream_t st,
                       const char* who) {
    const long long total = static_cast<long long>(a.tt.batch) * a.tt.tiles_per_image;
    if (a.tt.batch == 0) return DH_OK;
    a.box_cap = ((a.max_boxes > 0 ? a.max_boxes : 1) + 31) & ~31;
    const LossSmemLayout lay = loss_smem_layout<P, kFused>(a.tile_buf_bytes, a.tt.rows_per_tile, a.box_cap);
    if (lay.total > 227 * 1024)
        return set_error(DH_ERR_CAPACITY, "%s: needs %d bytes of shared memory", who, lay.total);
    DH_ONCE_PER_DEVICE(h) {
        DH_CUDA(cudaFuncSetAttribute(loss_kernel<P, kFused>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DH_CUDA(cudaFuncSetAttribute(loss_kernel<P, kFused, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    int per_sm = 1;  // resident CTAs per SM (registers and shared memory both count)
    DH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, loss_kernel<P, kFused>, DH_THREADS, lay.total));
    if (per_sm < 1) per_sm = 1;
    long long grid = static_cast<long long>(h->sm_count) * per_sm;
    // image-aligned chunks (a chunk never mixes images, so per-image sums stay separable): aim at >= 8 chunks
    // per CTA, 4..64 tiles each
    const int tpi = a.tt.tiles_per_image;
    long long want = total / (grid * 8);
    want = want < 4 ? 4 : (want > 64 ? 64 : want);
    const int n_sub = tpi > 0 ? static_cast<int>((tpi + want - 1) / want) : 1;
    a.chunk_tiles = tpi > 0 ? (tpi + n_sub - 1) / n_sub : 1;
    a.chunks_per_image = tpi > 0 ? (tpi + a.chunk_tiles - 1) / a.chunk_tiles : 1;
    const long long n_chunks = static_cast<long long>(a.tt.batch) * a.chunks_per_image;
    if (grid > n_chunks) grid = n_chunks;
    a.allow_vec = 1;
    for (int m = 0; m < a.tt.n_maps; ++m) {
        if ((reinterpret_cast<uintptr_t>(a.tt.maps[m].pred) & 15u) || (!kFused && (reinterpret_cast<uintptr_t>(a.tt.maps[m].out) & 15u)))
            a.allow_vec = 0;
        if ((a.tt.maps[m].image_stride & 3) != 0) a.allow_vec = 0;
        if (reinterpret_cast<uintptr_t>(a.grad_maps[m]) & 15u) a.allow_vec = 0;
    }
    // scratch: chunk partials + (optional) per-image sums when the caller only wants the total
    const size_t part_bytes = static_cast<size_t>(n_chunks) * 16;
    const size_t img_bytes = static_cast<size_t>(a.tt.batch) * 16;
    char* sc = static_cast<char*>(scratch(h, part_bytes + img_bytes + 512));
    if (!sc) return DH_ERR_CUDA;
    a.partials = reinterpret_cast<float*>(sc);
    float* per_image = out_per_image ? out_per_image : reinterpret_cast<float*>(sc + ((part_bytes + 255) & ~size_t(255)));
    if (total > 0) {
        a.sched = next_sched_counter(h, st);
        if (!a.sched) return DH_ERR_CUDA;
        bool grad = false;
        for (int m = 0; m < a.tt.n_maps; ++m) grad = grad || a.grad_maps[m] != nullptr;
        if (grad && !kFused) loss_kernel<P, kFused, true><<<static_cast<unsigned>(grid), DH_THREADS, lay.total, st>>>(a);
        else loss_kernel<P, kFused, false><<<static_cast<unsigned>(grid), DH_THREADS, lay.total, st>>>(a);
        DH_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    loss_finalize_images<<<a.tt.batch, 128, 0, st>>>(a.partials, total > 0 ? a.chunks_per_image : 0, per_image);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    if (out_total) {
        loss_finalize_total<<<1, 256, 0, st>>>(per_image, a.tt.batch, out_total);
        DH_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    return DH_OK;
}


static int loss_tile_bytes(const dh_handle_s* h) {
    (void)h;
    return 32768;  // work granule of the loss kernels (and the fused path's shared-memory target tile)
}

// Finalize kernels shared by both loss paths: chunk partials -> per-image sums -> total.
static int finalize_loss(dh_handle_s* h, const float* partials, int batch, int chunks_per_image, float* per_image,
                         float* out_total, cudaStream_t st) {
    loss_finalize_images<<<batch, 128, 0, st>>>(partials, chunks_per_image, per_image);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    if (out_total) {
        loss_finalize_total<<<1, 256, 0, st>>>(per_image, batch, out_total);
        DH_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    return DH_OK;
}

// The tiered chunk plan of the fused kernel (LossArgs::tiers): runs of images, each cut into chunks of its own size.
// Tier 0 holds most images in chunks of about `ct0` tiles; with `tail` the last images are cut into chunks of ct0/2,
// ct0/4, ... 1 tiles, each tier sized so that its tiles absorb the stagger of the tier before it (CTAs leave a tier spread
// over one of its chunk times: grid * ct / 2 tiles).  (Tried and not kept for small batches: a first wave whose chunks
// differ in size between the CTAs that share an SM, so that they would not stage / stream / resolve in lock step --
// 194 us against 188 us for 32 COCO images, 348 against 318 for 64.)
struct TierSpec {
    int chunk_tiles, images;
};
template <class P>
static void emit_tiers(LossArgs<P>& a, const TierSpec* spec, int n) {
    const int tpi = a.tt.tiles_per_image, batch = a.tt.batch;
    a.n_tiers = 0;
    long long chunk0 = 0;
    int image0 = 0;
    for (int k = 0; k < n && a.n_tiers < kMaxChunkTiers; ++k) {
        if (spec[k].images <= 0) continue;
        ChunkTier& T = a.tiers[a.n_tiers++];
        const int ct = spec[k].chunk_tiles < 1 ? 1 : spec[k].chunk_tiles;
        const int n_sub = tpi > 0 ? (tpi + ct - 1) / ct : 1;
        T.chunk_tiles = tpi > 0 ? (tpi + n_sub - 1) / n_sub : 1;
        T.cpi = tpi > 0 ? (tpi + T.chunk_tiles - 1) / T.chunk_tiles : 1;
        T.chunk0 = chunk0, T.image0 = image0, T.pad_ = 0;
        chunk0 += static_cast<long long>(spec[k].images) * T.cpi;
        image0 += spec[k].images;
    }
    if (a.n_tiers == 0 || image0 != batch) {  // (cannot happen; keeps a bad plan from skipping images)
        a.n_tiers = 1;
        a.tiers[0] = ChunkTier{0, 0, 1, tpi > 0 ? tpi : 1, 0};
        chunk0 = static_cast<long long>(batch) * a.tiers[0].cpi;
    }
    a.n_chunks = chunk0;
    a.chunk_tiles = a.tiers[0].chunk_tiles, a.chunks_per_image = a.tiers[0].cpi;
}

template <class P>
static void plan_tiers(LossArgs<P>& a, long long grid, int ct0, bool tail, int min_tiles) {
    const int tpi = a.tt.tiles_per_image, batch = a.tt.batch;
    TierSpec spec[kMaxChunkTiers];
    int n = 0;
    int cts[kMaxChunkTiers];
    int nt = 0;
    cts[nt++] = ct0;
    if (tail && tpi > 0)
        for (int c = ct0 / 2; c >= min_tiles && nt < 5; c /= 2) cts[nt++] = c;
    long long need[kMaxChunkTiers] = {}, need_total = 0;
    for (int k = 1; k < nt; ++k) {
        need[k] = (grid * cts[k - 1] / 2 + tpi - 1) / tpi;
        if (need[k] < 1) need[k] = 1;
        need_total += need[k];
    }
    if (need_total > batch / 2) {  // a small batch: the fine tiers share half of it (or vanish)
        for (int k = 1; k < nt; ++k) need[k] = need_total > 0 ? need[k] * (batch / 2) / need_total : 0;
    }
    int left = batch;
    for (int k = nt - 1; k >= 1; --k) left -= static_cast<int>(need[k]);
    spec[n++] = TierSpec{cts[0], left};
    for (int k = 1; k < nt; ++k) spec[n++] = TierSpec{cts[k], static_cast<int>(need[k])};
    emit_tiers(a, spec, n);
}

struct PlanOpts {
    int tail, max_chunk, chunks_per_cta;  // DH_OPT_FUSED_TAIL, DH_OPT_FUSED_MAX_CHUNK, DH_OPT_FUSED_CHUNKS_PER_CTA
};
// The chunk plan of one fused launch (host only; also behind dh_plan_fused_chunks for the CPU tests): fills a.tiers,
// a.n_chunks, a.span_fine from the tile table in a.tt and the grid.
template <class P>
static void plan_fused(const PlanOpts& o, LossArgs<P>& a, long long grid, bool grad) {
    const long long total = static_cast<long long>(a.tt.batch) * a.tt.tiles_per_image;
    // Chunk size.  Uniform chunks (DH_OPT_FUSED_TAIL = 0): aim at DH_OPT_FUSED_CHUNKS_PER_CTA chunks per CTA, 4..8 tiles each
    // (smaller chunks pay the chunk-end barriers too often, larger ones leave a tail).  Tiered (default): the tail is
    // taken care of by the finer tiers, so the first tier can be coarse -- about half of a CTA's share of the tiles,
    // 4..16 tiles -- which matters for small batches, where a chunk's fixed cost (GT staging, barriers) is not hidden
    // behind other CTAs' streaming.
    long long want;
    a.span_fine = (o.tail >> 1) ? (o.tail >> 1) : 2;
    // A chunk costs about 7 us whatever it holds (GT staging, pairing, resolving, the barriers) and a CTA streams at about
    // 11 GB/s, so a chunk under ~150 KB spends more time on that than on its bytes: that is 2 tiles of RetinaNet-COCO
    // rows (336 B), 6 of FCOS-VOC rows (100 B), all 16 of CenterNet-s8 rows (20 B) -- no tier goes below it.
    const long long tile_bytes = static_cast<long long>(a.tt.rows_per_tile) * a.tt.ch * 4;
    long long min_tiles = (150 * 1024 + tile_bytes - 1) / tile_bytes;
    min_tiles = min_tiles < 1 ? 1 : (min_tiles > o.max_chunk ? o.max_chunk : min_tiles);
    if (o.tail & 1) {
        want = (total / grid) * 55 / 100;
        want = want < 4 ? 4 : want;
        want = want < min_tiles ? min_tiles : want;
        want = want > o.max_chunk ? o.max_chunk : want;
    } else {
        want = total / (grid * o.chunks_per_cta);
        want = want < 4 ? 4 : (want > 8 ? 8 : want);
    }
    // A small launch -- at most three chunks per CTA -- is better off with EQUAL chunks (of up to kMaxChunkTiles = 18 tiles) when
    // their number fills whole waves of the grid: every CTA then does the same k chunks and nothing is left to balance.
    // 32 COCO images: 576 chunks of 18 tiles for 592 CTAs, 154 us against 175-182 us with tiers of 9, 4, 2 tiles; 64 images:
    // 1 152 chunks, 292 us against 306.  The largest chunk whose last wave is at least 85 % full wins (the fill of the last
    // wave IS the efficiency of such a plan); four waves measured slower than tiers (128 images: 555 us against 537), and
    // without a fitting chunk size the tiers take care of the tail.
    int uniform = 0;
    // (Not with the gradient: twice the traffic per tile halves the weight of a chunk's fixed cost, and its resolve pass sits
    // behind a block barrier -- 64 COCO images: 643 us with two 18-tile chunks per CTA against 612-622 us with tiers.)
    if (!grad && (o.tail & 1) && a.tt.tiles_per_image > 0) {
        double best = 0.0;
        const int tpi = a.tt.tiles_per_image;
        for (int c = (o.max_chunk == 16 ? kMaxChunkTiles : o.max_chunk); c >= (min_tiles > 4 ? min_tiles : 4); --c) {
            const int cpi = (tpi + c - 1) / c;
            const long long n = static_cast<long long>(cpi) * a.tt.batch, waves = (n + grid - 1) / grid;
            if (waves > 3) break;
            const double fill = static_cast<double>(n) / static_cast<double>(waves * grid);
            if (fill > best + 0.02) best = fill, uniform = (tpi + cpi - 1) / cpi;
        }
        if (best < 0.85) uniform = 0;
    }
    if (uniform) plan_tiers(a, grid, uniform, false, 1);
    else plan_tiers(a, grid, static_cast<int>(want), (o.tail & 1) != 0, static_cast<int>(min_tiles));
}

// Fused encode+loss, stream + correct formulation (dh_fused_loss_kernel.cuh): 256-row tiles, 32 rows per warp.  One
// launch: the kernel's last CTA reduces the chunk partials to the per-image sums and the total (and exchanges the
// total with the peer ranks when DH_OPT_LOSS_ALLREDUCE is on).
template <class P, int kCls, bool kGrad>
static int launch_fused_g(dh_handle_s* h, LossArgs<P>& a, float* out_per_image, float* out_total, cudaStream_t st,
                          const char* who) {
    const long long total = static_cast<long long>(a.tt.batch) * a.tt.tiles_per_image;
    const FusedSmemLayout lay = fused_smem_layout<P>(a.box_cap);
    DH_ONCE_PER_DEVICE(h) {
        DH_CUDA(cudaFuncSetAttribute(fused_loss_kernel<P, kCls, kGrad>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    int per_sm = 1;
    DH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_loss_kernel<P, kCls, kGrad>, DH_THREADS, lay.total));
    if (per_sm < 1) per_sm = 1;
    long long grid = static_cast<long long>(h->sm_count) * per_sm;
    const PlanOpts po{h->fused_tail, h->fused_max_chunk, h->fused_chunks_per_cta};
    plan_fused(po, a, grid, kGrad);
    const long long n_chunks = a.n_chunks;
    if (grid > n_chunks) grid = n_chunks;
    if (grid < 1) grid = 1;
    const size_t part_bytes = static_cast<size_t>(n_chunks) * 16;
    const size_t img_bytes = static_cast<size_t>(a.tt.batch) * 16;
    char* sc = static_cast<char*>(scratch(h, part_bytes + img_bytes + 512));
    if (!sc) return DH_ERR_CUDA;
    a.partials = reinterpret_cast<float*>(sc);
    a.per_image = out_per_image ? out_per_image : reinterpret_cast<float*>(sc + ((part_bytes + 255) & ~size_t(255)));
    a.out_total = out_total;
    a.fold_finalize = (out_per_image || out_total) ? 1 : 0;
    a.use_comm = 0;
    if (out_total) {
        if (const CommDev* cd = comm_dev_if_fused(h)) a.comm = *cd, a.use_comm = 1;
    }
    if (total == 0) {  // no tiles at all (maps of zero cells): the sums are zero
        if (a.fold_finalize) DH_CUDA(cudaMemsetAsync(a.per_image, 0, img_bytes, st));
        if (out_total) {
            DH_CUDA(cudaMemsetAsync(out_total, 0, 16, st));
            if (a.use_comm) return dh_allreduce_loss(h, out_total, 4, st);
        }
        return DH_OK;
    }
    a.sched = next_sched_counter(h, st);
    if (!a.sched) return DH_ERR_CUDA;
    a.trace = (h->trace && h->trace_bytes >= (grid * kTraceSlots + 1) * 8) ? h->trace : nullptr;
    fused_loss_kernel<P, kCls, kGrad><<<static_cast<unsigned>(grid), DH_THREADS, lay.total, st>>>(a);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    (void)who;
    return DH_OK;
}

// Picks the fused kernel: stream + correct when the class bitmask fits (<= 128 classes), else the shared-memory
// target-tile kernel (also selectable with DH_OPT_FUSED_LOSS_KERNEL = 1 for A/B checks).
template <class P>
static int launch_fused(dh_handle_s* h, LossArgs<P>& a, int num_classes, float* out_per_image, float* out_total,
                        cudaStream_t st, const char* who) {
    if (a.tt.batch == 0) return DH_OK;
    const int ch = a.tt.ch;
    if (h->fused_loss_kernel == 1 || num_classes > 32 * kCompactClassWords) {
        for (int m = 0; m < a.tt.n_maps; ++m)
            if (a.grad_maps[m])
                return set_error(DH_ERR_CAPACITY, "%s: the gradient needs the stream+correct kernel (<= %d classes, DH_OPT_FUSED_LOSS_KERNEL 0)",
                                 who, 32 * kCompactClassWords);
        a.tile_buf_bytes = finish_table(a.tt, ch, a.tt.batch, loss_tile_bytes(h));
        return launch_loss<P, true>(h, a, out_per_image, out_total, st, who);
    }
    finish_table(a.tt, ch, a.tt.batch, DH_THREADS * ch * 4);  // one row per thread
    a.box_cap = ((a.max_boxes > 0 ? a.max_boxes : 1) + 31) & ~31;
    a.allow_vec = 1;
    for (int m = 0; m < a.tt.n_maps; ++m)
        if ((reinterpret_cast<uintptr_t>(a.tt.maps[m].pred) & 15u) || (a.tt.maps[m].image_stride & 3)) a.allow_vec = 0;
    bool grad = false;
    for (int m = 0; m < a.tt.n_maps; ++m) {
        if (a.grad_maps[m]) grad = true;
        if (reinterpret_cast<uintptr_t>(a.grad_maps[m]) & 15u) a.allow_vec = 0;
    }
    const int kind = a.spec.cls_mode == 1 ? 2 : (a.spec.gamma == 2.0f ? 1 : 0);
    if (grad) {
        if (kind == 2) return launch_fused_g<P, 2, true>(h, a, out_per_image, out_total, st, who);
        if (kind == 1) return launch_fused_g<P, 1, true>(h, a, out_per_image, out_total, st, who);
        return launch_fused_g<P, 0, true>(h, a, out_per_image, out_total, st, who);
    }
    if (kind == 2) return launch_fused_g<P, 2, false>(h, a, out_per_image, out_total, st, who);
    if (kind == 1) return launch_fused_g<P, 1, false>(h, a, out_per_image, out_total, st, who);
    return launch_fused_g<P, 0, false>(h, a, out_per_image, out_total, st, who);
}

// Loss over materialised targets, stream + row-pass kernel (dh_dense_stream_kernel.cuh): 256-row tiles, 32 rows per warp.
template <int kCls, bool kGrad>
static int launch_dense_g(dh_handle_s* h, LossArgs<NoPolicy>& a, float* out_per_image, float* out_total, cudaStream_t st) {
    const long long total = static_cast<long long>(a.tt.batch) * a.tt.tiles_per_image;
    const size_t smem = (sizeof(LossArgs<NoPolicy>) + 127) & ~size_t(127);
    int per_sm = 1;
    DH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dense_stream_kernel<kCls, kGrad>, DH_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = static_cast<long long>(h->sm_count) * per_sm;
    const int tpi = a.tt.tiles_per_image;
    long long want = total / (grid * h->fused_chunks_per_cta);
    want = want < 4 ? 4 : (want > 8 ? 8 : want);
    const int n_sub = tpi > 0 ? static_cast<int>((tpi + want - 1) / want) : 1;
    a.chunk_tiles = tpi > 0 ? (tpi + n_sub - 1) / n_sub : 1;
    a.chunks_per_image = tpi > 0 ? (tpi + a.chunk_tiles - 1) / a.chunk_tiles : 1;
    const long long n_chunks = static_cast<long long>(a.tt.batch) * a.chunks_per_image;
    if (grid > n_chunks) grid = n_chunks;
    const size_t part_bytes = static_cast<size_t>(n_chunks) * 16;
    const size_t img_bytes = static_cast<size_t>(a.tt.batch) * 16;
    char* sc = static_cast<char*>(scratch(h, part_bytes + img_bytes + 512));
    if (!sc) return DH_ERR_CUDA;
    a.partials = reinterpret_cast<float*>(sc);
    float* per_image = out_per_image ? out_per_image : reinterpret_cast<float*>(sc + ((part_bytes + 255) & ~size_t(255)));
    if (total > 0) {
        a.sched = next_sched_counter(h, st);
        if (!a.sched) return DH_ERR_CUDA;
        dense_stream_kernel<kCls, kGrad><<<static_cast<unsigned>(grid), DH_THREADS, smem, st>>>(a);
        DH_CUDA(cudaGetLastError());
        h->launches += 1;
    }
    if (!out_per_image && !out_total) return DH_OK;  // gradient only
    return finalize_loss(h, a.partials, a.tt.batch, total > 0 ? a.chunks_per_image : 0, per_image, out_total, st);
}

static int launch_dense(dh_handle_s* h, LossArgs<NoPolicy>& a, float* out_per_image, float* out_total, cudaStream_t st) {
    if (a.tt.batch == 0) return DH_OK;
    const int ch = a.tt.ch;
    if (h->fused_loss_kernel == 1) {  // A/B: the shared-memory tile kernel
        a.tile_buf_bytes = finish_table(a.tt, ch, a.tt.batch, loss_tile_bytes(h));
        return launch_loss<NoPolicy, false>(h, a, out_per_image, out_total, st, "dh_dense_loss");
    }
    finish_table(a.tt, ch, a.tt.batch, DH_THREADS * ch * 4);
    a.allow_vec = 1;
    bool grad = false;
    for (int m = 0; m < a.tt.n_maps; ++m) {
        const MapDesc& md = a.tt.maps[m];
        if ((reinterpret_cast<uintptr_t>(md.pred) | reinterpret_cast<uintptr_t>(md.out) | reinterpret_cast<uintptr_t>(a.grad_maps[m])) & 15u)
            a.allow_vec = 0;
        if (md.image_stride & 3) a.allow_vec = 0;
        grad = grad || a.grad_maps[m] != nullptr;
    }
    const int kind = a.spec.cls_mode == 1 ? 2 : (a.spec.gamma == 2.0f ? 1 : 0);
    if (grad) {
        if (kind == 2) return launch_dense_g<2, true>(h, a, out_per_image, out_total, st);
        if (kind == 1) return launch_dense_g<1, true>(h, a, out_per_image, out_total, st);
        return launch_dense_g<0, true>(h, a, out_per_image, out_total, st);
    }
    if (kind == 2) return launch_dense_g<2, false>(h, a, out_per_image, out_total, st);
    if (kind == 1) return launch_dense_g<1, false>(h, a, out_per_image, out_total, st);
    return launch_dense_g<0, false>(h, a, out_per_image, out_total, st);
}

}  // namespace dh

using namespace dh;

struct GradOut {
    float w_cls, w_reg, w_cen;
    float* const* grad;  // per level (fused) / per map (dense)
};


static int dense_loss_impl(const GradOut* go, dh_handle_t h, int n_maps, const float* const* target_maps, const float* const* pred_maps,
                  const float* const* mask_maps, const int32_t* map_height, const int32_t* map_width,
                  const int32_t* map_sub, int batch, int ch, int reg_ch, int cen_mode, int reg_mode, int pos_rule, int cls_mode,
                  float alpha, float gamma, float delta, float* out_per_image, float* out_total, void* stream) {
    DH_CHECK_ARG(h && target_maps && pred_maps && map_height && map_width, "dh_dense_loss: NULL argument");
    DH_CHECK_ARG(n_maps >= 1 && n_maps <= DH_MAX_MAPS, "dh_dense_loss: n_maps %d not in [1,%d]", n_maps, DH_MAX_MAPS);
    DH_CHECK_ARG(batch >= 0 && ch >= 1, "dh_dense_loss: bad sizes");
    DH_CHECK_ARG(go || out_per_image || out_total, "dh_dense_loss: no output requested");
    LossArgs<NoPolicy> a;
    memset(&a, 0, sizeof(a));
    a.spec.reg_ch = reg_ch, a.spec.cen_mode = cen_mode, a.spec.reg_mode = reg_mode, a.spec.pos_rule = pos_rule;
    a.spec.cls_mode = cls_mode;
    a.spec.alpha = alpha, a.spec.gamma = gamma, a.spec.delta = delta;
    int rc = check_spec(a.spec, "dh_dense_loss");
    if (rc) return rc;
    DH_CHECK_ARG(ch >= reg_ch + (cen_mode != 0 ? 1 : 0), "dh_dense_loss: ch %d too small for the channel layout", ch);
    DH_CHECK_ARG(pos_rule != 2 || mask_maps, "dh_dense_loss: pos_rule 2 needs mask_maps");
    DeviceGuard guard(h);
    a.tt.n_maps = n_maps;
    for (int m = 0; m < n_maps; ++m) {
        DH_CHECK_ARG(target_maps[m] && pred_maps[m], "dh_dense_loss: map %d pointer is NULL", m);
        DH_CHECK_ARG(pos_rule != 2 || mask_maps[m], "dh_dense_loss: mask %d pointer is NULL", m);
        MapDesc& md = a.tt.maps[m];
        md.out = const_cast<float*>(target_maps[m]);
        md.pred = pred_maps[m];
        md.height = map_height[m], md.width = map_width[m], md.sub = map_sub ? map_sub[m] : 1;
        DH_CHECK_ARG(md.height >= 0 && md.width >= 0 && md.sub >= 1, "dh_dense_loss: map %d shape", m);
        md.rows = md.height * md.width * md.sub;
        md.image_stride = static_cast<long long>(md.rows) * ch;
        md.level = m, md.anchor = 0;
        a.mask_maps[m] = mask_maps ? mask_maps[m] : nullptr;
    }
    a.tt.ch = ch, a.tt.batch = batch;
    if (go) {
        a.spec.w_cls = go->w_cls, a.spec.w_reg = go->w_reg, a.spec.w_cen = go->w_cen;
        for (int m = 0; m < n_maps; ++m) {
            DH_CHECK_ARG(go->grad[m], "dh_dense_loss_grad: gradient map %d is NULL", m);
            a.grad_maps[m] = go->grad[m];
        }
    }
    return launch_dense(h, a, out_per_image, out_total, static_cast<cudaStream_t>(stream));
}

static int fcos_encode_loss_impl(const GradOut* go, dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                        int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, const float* b_dim,
                        int num_classes, int mode, const float* const* pred_levels, int reg_mode, int cen_mode,
                        float alpha, float gamma, float delta, float* out_per_image, float* out_total,
                        int32_t* num_targets, void* stream) {
    DH_CHECK_ARG(h && boxes && img_dim && strides && pred_levels, "dh_fcos_encode_loss: NULL argument");
    DH_CHECK_ARG(go || out_per_image || out_total, "dh_fcos_encode_loss: no output requested");
    DH_CHECK_ARG(batch >= 0 && max_boxes >= 0, "dh_fcos_encode_loss: bad sizes");
    if (max_boxes > DH_MAX_BOXES) return set_error(DH_ERR_CAPACITY, "dh_fcos_encode_loss: max_boxes %d > %d", max_boxes, DH_MAX_BOXES);
    DeviceGuard guard(h);
    LossArgs<FcosPolicy> a;
    memset(&a, 0, sizeof(a));
    int rc = fill_fcos(a.pp, a.tt, pad_h, pad_w, n_levels, strides, b_dim, num_classes, mode, nullptr, pred_levels,
                          num_targets, "dh_fcos_encode_loss");
    if (rc) return rc;
    a.spec.reg_ch = 4, a.spec.cen_mode = cen_mode, a.spec.reg_mode = reg_mode, a.spec.pos_rule = 0;
    a.spec.alpha = alpha, a.spec.gamma = gamma, a.spec.delta = delta;
    DH_CHECK_ARG(cen_mode >= 1 && cen_mode <= 3, "dh_fcos_encode_loss: cen_mode must be 1, 2 or 3");
    rc = check_spec(a.spec, "dh_fcos_encode_loss");
    if (rc) return rc;
    a.tt.ch = num_classes + 5, a.tt.batch = batch;
    a.boxes = boxes, a.nbox = nbox, a.img_dim = img_dim, a.max_boxes = max_boxes;
    a.pp.status = h->dev_status;
    if (go) {
        a.spec.w_cls = go->w_cls, a.spec.w_reg = go->w_reg, a.spec.w_cen = go->w_cen;
        for (int m = 0; m < a.tt.n_maps; ++m) {
            const int l = a.tt.maps[m].level;
            DH_CHECK_ARG(go->grad[l], "gradient pointer of level %d is NULL", l);
            a.grad_maps[m] = go->grad[l] + (a.tt.maps[m].pred - pred_levels[l]);
        }
    }
    return launch_fused<FcosPolicy>(h, a, num_classes, out_per_image, out_total, static_cast<cudaStream_t>(stream), "dh_fcos_encode_loss");
}

static int retina_encode_loss_impl(const GradOut* go, dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                          int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, int n_anchors,
                          const float* anchor_hw, float iou_thresh, int num_classes, const float* const* pred_levels,
                          float alpha, float gamma, float delta, float* out_per_image, float* out_total,
                          int32_t* num_pairs, void* stream) {
    DH_CHECK_ARG(h && boxes && img_dim && strides && anchor_hw && pred_levels, "dh_retina_encode_loss: NULL argument");
    DH_CHECK_ARG(go || out_per_image || out_total, "dh_retina_encode_loss: no output requested");
    DH_CHECK_ARG(batch >= 0 && max_boxes >= 0, "dh_retina_encode_loss: bad sizes");
    if (max_boxes > DH_MAX_BOXES) return set_error(DH_ERR_CAPACITY, "dh_retina_encode_loss: max_boxes %d > %d", max_boxes, DH_MAX_BOXES);
    DeviceGuard guard(h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LossArgs<RetinaPolicy> a;
    memset(&a, 0, sizeof(a));
    int rc = fill_retina(a.pp, a.tt, pad_h, pad_w, n_levels, strides, n_anchors, anchor_hw, iou_thresh, num_classes,
                            nullptr, pred_levels, num_pairs, "dh_retina_encode_loss");
    if (rc) return rc;
    a.spec.reg_ch = 4, a.spec.cen_mode = 0, a.spec.reg_mode = 0, a.spec.pos_rule = 1;
    a.spec.alpha = alpha, a.spec.gamma = gamma, a.spec.delta = delta;
    a.tt.ch = num_classes + 4, a.tt.batch = batch;
    a.boxes = boxes, a.nbox = nbox, a.img_dim = img_dim, a.max_boxes = max_boxes;
    a.pp.status = h->dev_status;
    if (num_pairs && batch > 0) DH_CUDA(cudaMemsetAsync(num_pairs, 0, sizeof(int32_t) * batch, st));
    if (go) {
        a.spec.w_cls = go->w_cls, a.spec.w_reg = go->w_reg, a.spec.w_cen = go->w_cen;
        for (int m = 0; m < a.tt.n_maps; ++m) {
            const int l = a.tt.maps[m].level;
            DH_CHECK_ARG(go->grad[l], "gradient pointer of level %d is NULL", l);
            a.grad_maps[m] = go->grad[l] + (a.tt.maps[m].pred - pred_levels[l]);
        }
    }
    return launch_fused<RetinaPolicy>(h, a, num_classes, out_per_image, out_total, st, "dh_retina_encode_loss");
}

static int centernet_encode_loss_impl(const GradOut* go, dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                             int max_boxes, int pad0, int pad1, int stride, int n_scales, const float* box_scales,
                             float sigma, int num_classes, int mode, const float* pred, int reg_mode, int cls_mode, float alpha,
                             float gamma, float delta, float* out_per_image, float* out_total, int32_t* status,
                             void* stream) {
    DH_CHECK_ARG(h && boxes && img_dim && pred, "dh_centernet_encode_loss: NULL argument");
    DH_CHECK_ARG(go || out_per_image || out_total, "dh_centernet_encode_loss: no output requested");
    DH_CHECK_ARG(batch >= 0 && max_boxes >= 0, "dh_centernet_encode_loss: bad sizes");
    if (max_boxes > DH_MAX_BOXES) return set_error(DH_ERR_CAPACITY, "dh_centernet_encode_loss: max_boxes %d > %d", max_boxes, DH_MAX_BOXES);
    DeviceGuard guard(h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LossArgs<CenterNetPolicy> a;
    memset(&a, 0, sizeof(a));
    int rc = fill_centernet(a.pp, a.tt, pad0, pad1, stride, n_scales, box_scales, sigma, num_classes, mode, nullptr,
                               pred, status, "dh_centernet_encode_loss");
    if (rc) return rc;
    const bool falloff = mode == DH_CENTERNET_POWER_FALLOFF || mode == DH_CENTERNET_GAUSSIAN;  // the [H,W,C+5] footprint layouts
    a.spec.reg_ch = 4, a.spec.cen_mode = falloff ? 1 : 0, a.spec.reg_mode = falloff ? reg_mode : 0;
    a.spec.pos_rule = falloff ? 0 : 1;  // tf_centernet.py:435 (>= 1) vs tf_centernet_resnet_s8.py:376 (> 0)
    a.spec.cls_mode = cls_mode;
    a.spec.alpha = alpha, a.spec.gamma = gamma, a.spec.delta = delta;
    rc = check_spec(a.spec, "dh_centernet_encode_loss");
    if (rc) return rc;
    a.tt.ch = num_classes + ((falloff || mode == DH_CENTERNET_HOURGLASS4) ? 5 : 4), a.tt.batch = batch;
    a.boxes = boxes, a.nbox = nbox, a.img_dim = img_dim, a.max_boxes = max_boxes;
    if (status) DH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    else a.pp.status = h->dev_status;
    if (go) {
        a.spec.w_cls = go->w_cls, a.spec.w_reg = go->w_reg, a.spec.w_cen = go->w_cen;
        DH_CHECK_ARG(go->grad[0], "gradient pointer is NULL");
        a.grad_maps[0] = go->grad[0];
    }
    return launch_fused<CenterNetPolicy>(h, a, num_classes, out_per_image, out_total, st, "dh_centernet_encode_loss");
}

extern "C" {

int dh_dense_loss(dh_handle_t h, int n_maps, const float* const* target_maps, const float* const* pred_maps,
                  const float* const* mask_maps, const int32_t* map_height, const int32_t* map_width,
                  const int32_t* map_sub, int batch, int ch, int reg_ch, int cen_mode, int reg_mode, int pos_rule, int cls_mode,
                  float alpha, float gamma, float delta, float* out_per_image, float* out_total, void* stream) {
    return dense_loss_impl(nullptr, h, n_maps, target_maps, pred_maps, mask_maps, map_height, map_width, map_sub, batch, ch, reg_ch,
                           cen_mode, reg_mode, pos_rule, cls_mode, alpha, gamma, delta, out_per_image, out_total, stream);
}
int dh_dense_loss_grad(dh_handle_t h, int n_maps, const float* const* target_maps, const float* const* pred_maps,
                       const float* const* mask_maps, const int32_t* map_height, const int32_t* map_width,
                       const int32_t* map_sub, int batch, int ch, int reg_ch, int cen_mode, int reg_mode, int pos_rule, int cls_mode,
                       float alpha, float gamma, float delta, float w_cls, float w_reg, float w_cen, float* const* grad_maps,
                       float* out_per_image, float* out_total, void* stream) {
    DH_CHECK_ARG(grad_maps, "dh_dense_loss_grad: grad_maps is NULL");
    const GradOut go = {w_cls, w_reg, w_cen, grad_maps};
    return dense_loss_impl(&go, h, n_maps, target_maps, pred_maps, mask_maps, map_height, map_width, map_sub, batch, ch, reg_ch,
                           cen_mode, reg_mode, pos_rule, cls_mode, alpha, gamma, delta, out_per_image, out_total, stream);
}

int dh_fcos_encode_loss(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                        int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, const float* b_dim,
                        int num_classes, int mode, const float* const* pred_levels, int reg_mode, int cen_mode,
                        float alpha, float gamma, float delta, float* out_per_image, float* out_total,
                        int32_t* num_targets, void* stream) {
    return fcos_encode_loss_impl(nullptr, h, boxes, nbox, img_dim, batch, max_boxes, pad_h, pad_w, n_levels, strides, b_dim, num_classes,
                                 mode, pred_levels, reg_mode, cen_mode, alpha, gamma, delta, out_per_image, out_total, num_targets, stream);
}
int dh_fcos_encode_loss_grad(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                             int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, const float* b_dim,
                             int num_classes, int mode, const float* const* pred_levels, int reg_mode, int cen_mode,
                             float alpha, float gamma, float delta, float w_cls, float w_reg, float w_cen,
                             float* const* grad_levels, float* out_per_image, float* out_total, int32_t* num_targets,
                             void* stream) {
    DH_CHECK_ARG(grad_levels, "dh_fcos_encode_loss_grad: grad_levels is NULL");
    const GradOut go = {w_cls, w_reg, w_cen, grad_levels};
    return fcos_encode_loss_impl(&go, h, boxes, nbox, img_dim, batch, max_boxes, pad_h, pad_w, n_levels, strides, b_dim, num_classes,
                                 mode, pred_levels, reg_mode, cen_mode, alpha, gamma, delta, out_per_image, out_total, num_targets, stream);
}

int dh_retina_encode_loss(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                          int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, int n_anchors,
                          const float* anchor_hw, float iou_thresh, int num_classes, const float* const* pred_levels,
                          float alpha, float gamma, float delta, float* out_per_image, float* out_total,
                          int32_t* num_pairs, void* stream) {
    return retina_encode_loss_impl(nullptr, h, boxes, nbox, img_dim, batch, max_boxes, pad_h, pad_w, n_levels, strides, n_anchors, anchor_hw,
                                   iou_thresh, num_classes, pred_levels, alpha, gamma, delta, out_per_image, out_total, num_pairs, stream);
}
int dh_retina_encode_loss_grad(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                               int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, int n_anchors,
                               const float* anchor_hw, float iou_thresh, int num_classes, const float* const* pred_levels,
                               float alpha, float gamma, float delta, float w_cls, float w_reg, float* const* grad_levels,
                               float* out_per_image, float* out_total, int32_t* num_pairs, void* stream) {
    DH_CHECK_ARG(grad_levels, "dh_retina_encode_loss_grad: grad_levels is NULL");
    const GradOut go = {w_cls, w_reg, 0.f, grad_levels};
    return retina_encode_loss_impl(&go, h, boxes, nbox, img_dim, batch, max_boxes, pad_h, pad_w, n_levels, strides, n_anchors, anchor_hw,
                                   iou_thresh, num_classes, pred_levels, alpha, gamma, delta, out_per_image, out_total, num_pairs, stream);
}

int dh_centernet_encode_loss(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                             int max_boxes, int pad0, int pad1, int stride, int n_scales, const float* box_scales,
                             float sigma, int num_classes, int mode, const float* pred, int reg_mode, int cls_mode, float alpha,
                             float gamma, float delta, float* out_per_image, float* out_total, int32_t* status,
                             void* stream) {
    return centernet_encode_loss_impl(nullptr, h, boxes, nbox, img_dim, batch, max_boxes, pad0, pad1, stride, n_scales, box_scales, sigma,
                                      num_classes, mode, pred, reg_mode, cls_mode, alpha, gamma, delta, out_per_image, out_total, status, stream);
}
int dh_centernet_encode_loss_grad(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                                  int max_boxes, int pad0, int pad1, int stride, int n_scales, const float* box_scales,
                                  float sigma, int num_classes, int mode, const float* pred, int reg_mode, int cls_mode, float alpha,
                                  float gamma, float delta, float w_cls, float w_reg, float w_cen, float* grad,
                                  float* out_per_image, float* out_total, int32_t* status, void* stream) {
    DH_CHECK_ARG(grad, "dh_centernet_encode_loss_grad: grad is NULL");
    float* const levels[1] = {grad};
    const GradOut go = {w_cls, w_reg, w_cen, levels};
    return centernet_encode_loss_impl(&go, h, boxes, nbox, img_dim, batch, max_boxes, pad0, pad1, stride, n_scales, box_scales, sigma,
                                      num_classes, mode, pred, reg_mode, cls_mode, alpha, gamma, delta, out_per_image, out_total, status, stream);
}

int dh_plan_fused_chunks(int batch, int tiles_per_image, int rows_per_tile, int ch, int grid, int with_grad, int tail, int max_chunk,
                         int32_t* out) {
    if (batch < 0 || tiles_per_image < 0 || rows_per_tile < 1 || ch < 1 || grid < 1 || !out || max_chunk < 4 || max_chunk > 16)
        return set_error(DH_ERR_BAD_ARG, "dh_plan_fused_chunks: bad argument");
    LossArgs<RetinaPolicy>* a = new LossArgs<RetinaPolicy>();  // (only the tile table's sizes and the plan fields are touched)
    memset(a, 0, sizeof(*a));
    a->tt.batch = batch, a->tt.tiles_per_image = tiles_per_image, a->tt.rows_per_tile = rows_per_tile, a->tt.ch = ch;
    const PlanOpts po{tail, max_chunk, 12};
    plan_fused(po, *a, grid, with_grad != 0);
    out[0] = a->n_tiers, out[1] = static_cast<int32_t>(a->n_chunks);
    for (int k = 0; k < 8; ++k) {
        const bool live = k < a->n_tiers;
        out[2 + 4 * k + 0] = live ? a->tiers[k].image0 : 0, out[2 + 4 * k + 1] = live ? a->tiers[k].chunk_tiles : 0;
        out[2 + 4 * k + 2] = live ? a->tiers[k].cpi : 0, out[2 + 4 * k + 3] = live ? static_cast<int32_t>(a->tiers[k].chunk0) : 0;
    }
    const int n = a->n_tiers;
    delete a;
    return n;
}

}  // extern "C"
