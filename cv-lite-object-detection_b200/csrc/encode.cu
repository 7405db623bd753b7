// C-ABI launchers for the target encoders (see include/densehead.h).
#include <cstring>

#include "dh_encode_direct_kernel.cuh"
#include "dh_host.h"
#include "dh_launch.h"

namespace dh {

int auto_tile_bytes(const TileTable& tt, int ch, int batch, int max_tile_bytes, int sm_count) {
    // big tiles amortise the per-tile bookkeeping (6.3 TB/s at 48 KB vs 3.5 TB/s at 16 KB on B200), but a small
    // problem still has to spread over all SMs: aim at >= 4 tiles per SM, never below 8 KB
    long long bytes = 0;
    for (int m = 0; m < tt.n_maps; ++m) bytes += static_cast<long long>(tt.maps[m].rows) * ch * 4;
    bytes *= batch;
    long long t = bytes / (static_cast<long long>(sm_count) * 4);
    if (t > max_tile_bytes) t = max_tile_bytes;
    if (t < 8192) t = 8192;
    return static_cast<int>(t) & ~1023;
}

int finish_table(TileTable& tt, int ch, int batch, int tile_bytes) {
    int rpt = (tile_bytes / (ch * 4)) & ~3;
    if (rpt < 4) rpt = 4;
    if (rpt > DH_THREADS * 32) rpt = DH_THREADS * 32;
    tt.ch = ch;
    tt.rows_per_tile = rpt;
    tt.batch = batch;
    int t = 0;
    for (int m = 0; m < tt.n_maps; ++m) {
        MapDesc& md = tt.maps[m];
        md.tile_begin = t;
        md.n_tiles = (md.rows + rpt - 1) / rpt;
        md.div_width = make_fastdiv(static_cast<uint32_t>(md.width > 0 ? md.width : 1));
        md.div_sub = make_fastdiv(static_cast<uint32_t>(md.sub > 0 ? md.sub : 1));
        t += md.n_tiles;
    }
    tt.tiles_per_image = t;
    return ((rpt * ch * 4 + 16) + 127) & ~127;  // shared-memory bytes per stage
}

// Output bytes up to which DH_OPT_ENCODE_KERNEL = 0 picks the direct-store kernel (measured on B200)
constexpr long long kDirectMaxBytes = 512ll << 20;

static long long map_bytes(const TileTable& tt, int ch, int batch) {
    long long bytes = 0;
    for (int m = 0; m < tt.n_maps; ++m) bytes += static_cast<long long>(tt.maps[m].rows) * ch * 4;
    return bytes * batch;
}

// The direct-store kernel (dh_encode_direct_kernel.cuh): one CTA per image-aligned chunk, cut statically.
template <class P>
static int launch_encode_direct(dh_handle_s* h, EncodeArgs<P>& a, cudaStream_t st, const char* who) {
    const int ch = a.tt.ch, batch = a.tt.batch;
    const long long bytes = map_bytes(a.tt, ch, batch);
    // the unit ("tile") is ~4 KB of rows; a chunk is a run of tiles of one image sized so that the problem makes about
    // 8 CTAs per SM, at least 8 KB and at most 128 KB each (small images: several whole images per chunk)
    a.tile_buf_bytes = finish_table(a.tt, ch, batch, 4096);
    const long long total = static_cast<long long>(batch) * a.tt.tiles_per_image;
    if (total == 0) return DH_OK;
    a.box_cap = ((a.max_boxes > 0 ? a.max_boxes : 1) + 31) & ~31;
    const DirectSmemLayout lay = direct_smem_layout<P>(a.box_cap);
    if (lay.total > 227 * 1024) return set_error(DH_ERR_CAPACITY, "%s: needs %d bytes of shared memory", who, lay.total);
    DH_ONCE_PER_DEVICE(h) {
        DH_CUDA(cudaFuncSetAttribute(encode_direct_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    int per_sm = 1;
    DH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_direct_kernel<P>, DH_THREADS, lay.total));
    if (per_sm < 1) per_sm = 1;
    // one wave of resident CTAs when the chunks then hold 8..128 KB each
    // Chunks per SM the problem is cut into (measured, tools/encode_ab.py with this value swept 2..16 on FCOS-VOC and
    // CenterNet-s8 shapes from 0.5 to 420 MB): a large output wants many medium chunks -- 14 per SM from 128 MB on: 140 MB in
    // 34.7 us against 38.4 with 4 per SM, 419 MB in 78.4 us against 93.8 (and 86.5 for the tile streamer) --, a small one few
    // CTAs (4 per SM below 24 MB: 8 images of 4.4 MB in 6.3 us against 8.8 with 8 per SM); images with many boxes want
    // large chunks whatever the size, because every chunk rebuilds the image's records (3 per SM: CenterNet, 150 boxes,
    // 52 MB in 16.4 us against 17.7).
    if (bytes >= (128ll << 20)) per_sm = 14;
    else if (a.max_boxes > 64) per_sm = 3;
    else if (bytes < (24ll << 20)) per_sm = per_sm < 4 ? per_sm : 4;
    else per_sm = 8;
    long long want_bytes = bytes / (static_cast<long long>(h->sm_count) * per_sm) + 1;
    const long long tile_b = static_cast<long long>(a.tt.rows_per_tile) * ch * 4;
    long long want = (want_bytes + tile_b - 1) / tile_b;
    if (want < 1) want = 1;
    const int tpi = a.tt.tiles_per_image;
    if (tpi <= want) {
        a.chunks_per_image = 1;
        a.images_per_chunk = static_cast<int>(want / tpi);
        if (a.images_per_chunk < 1) a.images_per_chunk = 1;
        a.chunk_tiles = a.images_per_chunk * tpi;
        a.n_chunks = (batch + a.images_per_chunk - 1) / a.images_per_chunk;
    } else {
        const int n_sub = static_cast<int>((tpi + want - 1) / want);
        a.chunk_tiles = (tpi + n_sub - 1) / n_sub;
        a.chunks_per_image = (tpi + a.chunk_tiles - 1) / a.chunk_tiles;
        a.images_per_chunk = 1;
        a.n_chunks = static_cast<long long>(batch) * a.chunks_per_image;
    }
    a.sched = nullptr;
    a.use_tma_store = 0;
    a.phase_cycles = nullptr;
    encode_direct_kernel<P><<<static_cast<unsigned>(a.n_chunks), DH_THREADS, lay.total, st>>>(a);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

template <class P>
static int launch_encode(dh_handle_s* h, EncodeArgs<P>& a, cudaStream_t st, const char* who) {
    if (h->encode_kernel == 2 || (h->encode_kernel == 0 && map_bytes(a.tt, a.tt.ch, a.tt.batch) <= kDirectMaxBytes))
        return launch_encode_direct<P>(h, a, st, who);
    const long long total = static_cast<long long>(a.tt.batch) * a.tt.tiles_per_image;
    if (total == 0) return DH_OK;
    a.box_cap = ((a.max_boxes > 0 ? a.max_boxes : 1) + 31) & ~31;
    const EncodeSmemLayout lay = encode_smem_layout<P>(a.tile_buf_bytes, a.box_cap);
    if (lay.total > 227 * 1024)
        return set_error(DH_ERR_CAPACITY, "%s: tile of %d bytes needs %d bytes of shared memory", who, a.tile_buf_bytes,
                         lay.total);
    DH_ONCE_PER_DEVICE(h) {  // (and per template instantiation)
        DH_CUDA(cudaFuncSetAttribute(encode_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    int per_sm = 1;  // resident CTAs per SM (registers and shared memory both count)
    DH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_kernel<P>, DH_THREADS, lay.total));
    // one CTA per SM leaves the TMA stores of a tile uncovered by another CTA's bookkeeping (CenterNet with 150 boxes:
    // 117 KB at 48 KB tiles, 3.5 TB/s): shrink the tile until two fit
    EncodeSmemLayout lay2 = lay;
    for (int tb = a.tile_buf_bytes; per_sm < 2 && h->ctas_per_sm >= 2 && tb > 20 * 1024;) {
        tb -= 4096;
        a.tile_buf_bytes = finish_table(a.tt, a.tt.ch, a.tt.batch, tb - 256);
        lay2 = encode_smem_layout<P>(a.tile_buf_bytes, a.box_cap);
        DH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_kernel<P>, DH_THREADS, lay2.total));
    }
    if (per_sm > h->ctas_per_sm) per_sm = h->ctas_per_sm;
    if (per_sm < 1) per_sm = 1;
    const long long total2 = static_cast<long long>(a.tt.batch) * a.tt.tiles_per_image;  // the table may have been re-cut
    long long grid = static_cast<long long>(h->sm_count) * per_sm;
    if (grid > total2) grid = total2;
    // image-aligned chunks for the dynamic scheduler: aim at ~8 chunks per CTA of up to 64 tiles each; a problem with
    // fewer tiles than that is cut into about one chunk per CTA (2..8 tiles) so that it still spreads over every SM --
    // per CTA the work is a chain of latencies, so a 4 MB output finishes sooner on 592 CTAs x 1-2 tiles than on 72 x 11
    {
        const int tpi = a.tt.tiles_per_image;
        long long want = total2 / (grid * 8);
        long long lo = total2 / grid;  // measured (tools/quick_encode_bench.py): about one chunk per CTA, 2..8 tiles each
        lo = lo < h->encode_min_chunk ? h->encode_min_chunk : (lo > 8 ? 8 : lo);
        want = want < lo ? lo : (want > 64 ? 64 : want);
        if (tpi <= want) {
            a.chunks_per_image = 1;
            a.images_per_chunk = static_cast<int>(want / tpi);
            if (a.images_per_chunk < 1) a.images_per_chunk = 1;
            a.chunk_tiles = a.images_per_chunk * tpi;
            a.n_chunks = (a.tt.batch + a.images_per_chunk - 1) / a.images_per_chunk;
        } else {
            const int n_sub = static_cast<int>((tpi + want - 1) / want);
            a.chunk_tiles = (tpi + n_sub - 1) / n_sub;
            a.chunks_per_image = (tpi + a.chunk_tiles - 1) / a.chunk_tiles;
            a.images_per_chunk = 1;
            a.n_chunks = static_cast<long long>(a.tt.batch) * a.chunks_per_image;
        }
        if (grid > a.n_chunks) grid = a.n_chunks;
    }
    a.sched = next_sched_counter(h, st);
    if (!a.sched) return DH_ERR_CUDA;
    a.use_tma_store = h->use_tma_store;
    a.phase_cycles = h->phase_cycles;
    encode_kernel<P><<<static_cast<unsigned>(grid), DH_THREADS, lay2.total, st>>>(a);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

}  // namespace dh

namespace dh {

#define DH_FILL_CHECK(cond, ...)                              \
    do {                                                      \
        if (!(cond)) return set_error(DH_ERR_BAD_ARG, __VA_ARGS__); \
    } while (0)

int fill_fcos(FcosPolicy::Params& p, TileTable& tt, int pad_h, int pad_w, int n_levels, const int32_t* strides,
              const float* b_dim, int num_classes, int mode, float* const* out_levels,
              const float* const* pred_levels, int32_t* num_targets, const char* who) {
    DH_FILL_CHECK(n_levels >= 1 && n_levels <= DH_MAX_LEVELS, "%s: n_levels %d not in [1,%d]", who, n_levels, DH_MAX_LEVELS);
    DH_FILL_CHECK(n_levels == 1 || b_dim, "%s: b_dim is NULL", who);
    DH_FILL_CHECK(pad_h > 0 && pad_w > 0, "%s: bad padded size", who);
    DH_FILL_CHECK(num_classes >= 1 && num_classes <= 4096, "%s: num_classes %d", who, num_classes);
    DH_FILL_CHECK(mode >= 0 && mode <= 4, "%s: mode %d", who, mode);
    p.n_levels = n_levels, p.num_classes = num_classes, p.mode = mode, p.num_targets = num_targets;
    tt.n_maps = n_levels;
    for (int l = 0; l < n_levels; ++l) {
        DH_FILL_CHECK(strides[l] > 0, "%s: stride of level %d", who, l);
        DH_FILL_CHECK((!out_levels || out_levels[l]) && (!pred_levels || pred_levels[l]), "%s: level %d pointer is NULL", who, l);
        p.stride[l] = strides[l];
        p.stride_f[l] = static_cast<float>(strides[l]);
        if (l < n_levels - 1) p.b_dim[l] = b_dim[l];
        p.hl[l] = static_cast<int>(static_cast<double>(pad_h) / strides[l]);  // int(img_pad[0] / stride)
        p.wl[l] = static_cast<int>(static_cast<double>(pad_w) / strides[l]);
        MapDesc& md = tt.maps[l];
        md.out = out_levels ? out_levels[l] : nullptr;
        md.pred = pred_levels ? pred_levels[l] : nullptr;
        md.rows = p.hl[l] * p.wl[l];
        md.height = p.hl[l], md.width = p.wl[l], md.sub = 1, md.level = l, md.anchor = 0;
        md.image_stride = static_cast<long long>(md.rows) * (num_classes + 5);
    }
    return DH_OK;
}

int fill_retina(RetinaPolicy::Params& p, TileTable& tt, int pad_h, int pad_w, int n_levels, const int32_t* strides,
                int n_anchors, const float* anchor_hw, float iou_thresh, int num_classes, float* const* out_levels,
                const float* const* pred_levels, int32_t* num_pairs, const char* who) {
    DH_FILL_CHECK(n_levels >= 1 && n_levels <= DH_MAX_LEVELS, "%s: n_levels %d", who, n_levels);
    DH_FILL_CHECK(n_anchors >= 1 && n_anchors <= 12, "%s: n_anchors %d not in [1,12]", who, n_anchors);
    if (n_levels * n_anchors > DH_MAX_MAPS)
        return set_error(DH_ERR_CAPACITY, "%s: %d maps > %d", who, n_levels * n_anchors, DH_MAX_MAPS);
    DH_FILL_CHECK(pad_h > 0 && pad_w > 0, "%s: bad padded size", who);
    DH_FILL_CHECK(num_classes >= 1 && num_classes <= 4096, "%s: num_classes %d", who, num_classes);
    p.n_levels = n_levels, p.n_anchors = n_anchors, p.num_classes = num_classes, p.thr = iou_thresh;
    p.num_pairs = num_pairs;
    const int ch = num_classes + 4;
    int m = 0;
    for (int l = 0; l < n_levels; ++l) {
        DH_FILL_CHECK(strides[l] > 0, "%s: stride of level %d", who, l);
        DH_FILL_CHECK((!out_levels || out_levels[l]) && (!pred_levels || pred_levels[l]), "%s: level %d pointer is NULL", who, l);
        p.stride[l] = strides[l];
        const int hl = static_cast<int>(static_cast<double>(pad_h) / strides[l]);
        const int wl = static_cast<int>(static_cast<double>(pad_w) / strides[l]);
        for (int an = 0; an < n_anchors; ++an, ++m) {
            p.anchor_h[l][an] = anchor_hw[(l * n_anchors + an) * 2];
            p.anchor_w[l][an] = anchor_hw[(l * n_anchors + an) * 2 + 1];
            MapDesc& md = tt.maps[m];
            md.rows = hl * wl;
            md.height = hl, md.width = wl, md.sub = 1, md.level = l, md.anchor = an;
            const long long map_off = static_cast<long long>(an) * md.rows * ch;
            md.out = out_levels ? out_levels[l] + map_off : nullptr;
            md.pred = pred_levels ? pred_levels[l] + map_off : nullptr;
            md.image_stride = static_cast<long long>(n_anchors) * md.rows * ch;
        }
    }
    tt.n_maps = m;
    return DH_OK;
}

int fill_centernet(CenterNetPolicy::Params& p, TileTable& tt, int pad0, int pad1, int stride, int n_scales,
                   const float* box_scales, float sigma, int num_classes, int mode, float* out, const float* pred,
                   int32_t* status, const char* who) {
    DH_FILL_CHECK(mode >= 0 && mode <= 4, "%s: mode %d", who, mode);
    DH_FILL_CHECK(mode != DH_CENTERNET_ONEHOT_SCALES || (box_scales && n_scales >= 1 && n_scales <= 8),
                  "%s: mode 0 needs 1..8 box_scales", who);
    DH_FILL_CHECK(pad0 > 0 && pad1 > 0 && stride > 0, "%s: bad sizes", who);
    DH_FILL_CHECK(num_classes >= 1 && num_classes <= 4096, "%s: num_classes %d", who, num_classes);
    p.mode = mode, p.num_classes = num_classes, p.stride = stride, p.stride_f = static_cast<float>(stride);
    p.sigma = sigma, p.pad0 = pad0, p.pad1 = pad1, p.status = status;
    p.n_scales = (mode == DH_CENTERNET_ONEHOT_SCALES) ? n_scales : (mode == DH_CENTERNET_HOURGLASS4 ? 4 : 1);
    for (int n = 0; n < p.n_scales && mode == DH_CENTERNET_ONEHOT_SCALES; ++n) p.scales[n] = box_scales[n];
    if (mode == DH_CENTERNET_HOURGLASS4)  // img_dims / (8, 4, 2, 1), train_hourglass_voc.py:96-97 (square padded image)
        for (int n = 0; n < 4; ++n) p.scales[n] = static_cast<float>(static_cast<double>(pad0) / (1 << (3 - n)));
    MapDesc& md = tt.maps[0];
    tt.n_maps = 1;
    int hh, ww;
    if (mode == DH_CENTERNET_POWER_FALLOFF || mode == DH_CENTERNET_GAUSSIAN || mode == DH_CENTERNET_HOURGLASS4) {
        hh = static_cast<int>(static_cast<double>(pad0) / stride);
        ww = static_cast<int>(static_cast<double>(pad1) / stride);
    } else {  // the reference swaps the indices (tf_centernet_resnet_s8.py:259-260)
        hh = static_cast<int>(static_cast<double>(pad1) / stride);
        ww = static_cast<int>(static_cast<double>(pad0) / stride);
    }
    const int ch = num_classes + ((mode == DH_CENTERNET_POWER_FALLOFF || mode == DH_CENTERNET_GAUSSIAN || mode == DH_CENTERNET_HOURGLASS4) ? 5 : 4);
    md.out = out;
    md.pred = pred;
    md.height = hh, md.width = ww, md.sub = p.n_scales, md.level = 0, md.anchor = 0;
    md.rows = hh * ww * p.n_scales;
    md.image_stride = static_cast<long long>(md.rows) * ch;
    return DH_OK;
}

}  // namespace dh

using namespace dh;

extern "C" {

int dh_fcos_encode(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                   int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, const float* b_dim,
                   int num_classes, int mode, float* const* out_levels, int32_t* num_targets, void* stream) {
    DH_CHECK_ARG(h && boxes && img_dim && strides && out_levels, "dh_fcos_encode: NULL argument");
    DH_CHECK_ARG(batch >= 0 && max_boxes >= 0, "dh_fcos_encode: bad sizes");
    if (max_boxes > DH_MAX_BOXES) return set_error(DH_ERR_CAPACITY, "dh_fcos_encode: max_boxes %d > %d", max_boxes, DH_MAX_BOXES);
    DeviceGuard guard(h);
    EncodeArgs<FcosPolicy> a;
    memset(&a, 0, sizeof(a));
    int rc = fill_fcos(a.pp, a.tt, pad_h, pad_w, n_levels, strides, b_dim, num_classes, mode, out_levels, nullptr,
                       num_targets, "dh_fcos_encode");
    if (rc) return rc;
    a.tile_buf_bytes = finish_table(a.tt, num_classes + 5, batch, auto_tile_bytes(a.tt, num_classes + 5, batch, h->tile_bytes, h->sm_count));
    a.boxes = boxes, a.nbox = nbox, a.img_dim = img_dim, a.max_boxes = max_boxes;
    a.pp.status = h->dev_status;
    return launch_encode<FcosPolicy>(h, a, static_cast<cudaStream_t>(stream), "dh_fcos_encode");
}

int dh_retina_encode(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                     int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, int n_anchors,
                     const float* anchor_hw, float iou_thresh, int num_classes, float* const* out_levels,
                     int32_t* num_pairs, void* stream) {
    DH_CHECK_ARG(h && boxes && img_dim && strides && anchor_hw && out_levels, "dh_retina_encode: NULL argument");
    DH_CHECK_ARG(batch >= 0 && max_boxes >= 0, "dh_retina_encode: bad sizes");
    if (max_boxes > DH_MAX_BOXES) return set_error(DH_ERR_CAPACITY, "dh_retina_encode: max_boxes %d > %d", max_boxes, DH_MAX_BOXES);
    DeviceGuard guard(h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EncodeArgs<RetinaPolicy> a;
    memset(&a, 0, sizeof(a));
    int rc = fill_retina(a.pp, a.tt, pad_h, pad_w, n_levels, strides, n_anchors, anchor_hw, iou_thresh, num_classes,
                         out_levels, nullptr, num_pairs, "dh_retina_encode");
    if (rc) return rc;
    a.tile_buf_bytes = finish_table(a.tt, num_classes + 4, batch, auto_tile_bytes(a.tt, num_classes + 4, batch, h->tile_bytes, h->sm_count));
    a.boxes = boxes, a.nbox = nbox, a.img_dim = img_dim, a.max_boxes = max_boxes;
    a.pp.status = h->dev_status;
    if (num_pairs && batch > 0) DH_CUDA(cudaMemsetAsync(num_pairs, 0, sizeof(int32_t) * batch, st));
    return launch_encode<RetinaPolicy>(h, a, st, "dh_retina_encode");
}

int dh_centernet_encode(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                        int max_boxes, int pad0, int pad1, int stride, int n_scales, const float* box_scales,
                        float sigma, int num_classes, int mode, float* out, int32_t* status, void* stream) {
    DH_CHECK_ARG(h && boxes && img_dim && out, "dh_centernet_encode: NULL argument");
    DH_CHECK_ARG(batch >= 0 && max_boxes >= 0, "dh_centernet_encode: bad sizes");
    if (max_boxes > DH_MAX_BOXES) return set_error(DH_ERR_CAPACITY, "dh_centernet_encode: max_boxes %d > %d", max_boxes, DH_MAX_BOXES);
    DeviceGuard guard(h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EncodeArgs<CenterNetPolicy> a;
    memset(&a, 0, sizeof(a));
    int rc = fill_centernet(a.pp, a.tt, pad0, pad1, stride, n_scales, box_scales, sigma, num_classes, mode, out, nullptr,
                            status, "dh_centernet_encode");
    if (rc) return rc;
    const int ch = num_classes + ((mode == DH_CENTERNET_POWER_FALLOFF || mode == DH_CENTERNET_GAUSSIAN || mode == DH_CENTERNET_HOURGLASS4) ? 5 : 4);
    a.tile_buf_bytes = finish_table(a.tt, ch, batch, auto_tile_bytes(a.tt, ch, batch, h->tile_bytes, h->sm_count));
    a.boxes = boxes, a.nbox = nbox, a.img_dim = img_dim, a.max_boxes = max_boxes;
    if (status) DH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    else a.pp.status = h->dev_status;
    return launch_encode<CenterNetPolicy>(h, a, st, "dh_centernet_encode");
}

}  // extern "C"
