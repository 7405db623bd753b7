// C-ABI launchers for the target encoders (see include/densehead.h).
#include <cstring>

#include "dh_encode_kernel.cuh"
#include "dh_host.h"
#include "dh_launch.h"

namespace dh {

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

template <class P>
static int launch_encode(dh_handle_s* h, EncodeArgs<P>& a, cudaStream_t st, const char* who) {
    const long long total = static_cast<long long>(a.tt.batch) * a.tt.tiles_per_image;
    if (total == 0) return DH_OK;
    const EncodeSmemLayout lay = encode_smem_layout<P>(a.tile_buf_bytes);
    if (lay.total > 227 * 1024)
        return set_error(DH_ERR_CAPACITY, "%s: tile of %d bytes needs %d bytes of shared memory", who, a.tile_buf_bytes,
                         lay.total);
    static bool attr_done = false;  // per template instantiation
    if (!attr_done) {
        DH_CUDA(cudaFuncSetAttribute(encode_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    long long grid = static_cast<long long>(h->sm_count) * h->ctas_per_sm;
    if (grid > total) grid = total;
    a.use_tma_store = h->use_tma_store;
    encode_kernel<P><<<static_cast<unsigned>(grid), DH_THREADS, lay.total, st>>>(a);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

}  // namespace dh

using namespace dh;

extern "C" {

int dh_fcos_encode(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                   int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, const float* b_dim,
                   int num_classes, int mode, float* const* out_levels, int32_t* num_targets, void* stream) {
    DH_CHECK_ARG(h && boxes && img_dim && strides && out_levels, "dh_fcos_encode: NULL argument");
    DH_CHECK_ARG(n_levels >= 1 && n_levels <= DH_MAX_LEVELS, "dh_fcos_encode: n_levels %d not in [1,%d]", n_levels,
                 DH_MAX_LEVELS);
    DH_CHECK_ARG(n_levels == 1 || b_dim, "dh_fcos_encode: b_dim is NULL");
    DH_CHECK_ARG(batch >= 0 && max_boxes >= 0 && pad_h > 0 && pad_w > 0, "dh_fcos_encode: bad sizes");
    if (max_boxes > DH_MAX_BOXES) return set_error(DH_ERR_CAPACITY, "dh_fcos_encode: max_boxes %d > %d", max_boxes, DH_MAX_BOXES);
    DH_CHECK_ARG(num_classes >= 1 && num_classes <= 4096, "dh_fcos_encode: num_classes %d", num_classes);
    DH_CHECK_ARG(mode >= 0 && mode <= 3, "dh_fcos_encode: mode %d", mode);
    DeviceGuard guard(h->device);
    EncodeArgs<FcosPolicy> a;
    memset(&a, 0, sizeof(a));
    FcosPolicy::Params& p = a.pp;
    p.n_levels = n_levels, p.num_classes = num_classes, p.mode = mode, p.num_targets = num_targets;
    a.tt.n_maps = n_levels;
    for (int l = 0; l < n_levels; ++l) {
        DH_CHECK_ARG(strides[l] > 0 && out_levels[l], "dh_fcos_encode: level %d stride/pointer", l);
        p.stride[l] = strides[l];
        p.stride_f[l] = static_cast<float>(strides[l]);
        if (l < n_levels - 1) p.b_dim[l] = b_dim[l];
        p.hl[l] = static_cast<int>(static_cast<double>(pad_h) / strides[l]);  // int(img_pad[0] / stride)
        p.wl[l] = static_cast<int>(static_cast<double>(pad_w) / strides[l]);
        MapDesc& md = a.tt.maps[l];
        md.out = out_levels[l];
        md.rows = p.hl[l] * p.wl[l];
        md.height = p.hl[l], md.width = p.wl[l], md.sub = 1, md.level = l, md.anchor = 0;
        md.image_stride = static_cast<long long>(md.rows) * (num_classes + 5);
    }
    a.tile_buf_bytes = finish_table(a.tt, num_classes + 5, batch, h->tile_bytes);
    a.boxes = boxes, a.nbox = nbox, a.img_dim = img_dim, a.max_boxes = max_boxes;
    return launch_encode<FcosPolicy>(h, a, static_cast<cudaStream_t>(stream), "dh_fcos_encode");
}

int dh_retina_encode(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                     int max_boxes, int pad_h, int pad_w, int n_levels, const int32_t* strides, int n_anchors,
                     const float* anchor_hw, float iou_thresh, int num_classes, float* const* out_levels,
                     int32_t* num_pairs, void* stream) {
    DH_CHECK_ARG(h && boxes && img_dim && strides && anchor_hw && out_levels, "dh_retina_encode: NULL argument");
    DH_CHECK_ARG(n_levels >= 1 && n_levels <= DH_MAX_LEVELS, "dh_retina_encode: n_levels %d", n_levels);
    DH_CHECK_ARG(n_anchors >= 1 && n_anchors <= 12, "dh_retina_encode: n_anchors %d not in [1,12]", n_anchors);
    if (n_levels * n_anchors > DH_MAX_MAPS)
        return set_error(DH_ERR_CAPACITY, "dh_retina_encode: %d maps > %d", n_levels * n_anchors, DH_MAX_MAPS);
    DH_CHECK_ARG(batch >= 0 && max_boxes >= 0 && pad_h > 0 && pad_w > 0, "dh_retina_encode: bad sizes");
    if (max_boxes > DH_MAX_BOXES) return set_error(DH_ERR_CAPACITY, "dh_retina_encode: max_boxes %d > %d", max_boxes, DH_MAX_BOXES);
    DH_CHECK_ARG(num_classes >= 1 && num_classes <= 4096, "dh_retina_encode: num_classes %d", num_classes);
    DeviceGuard guard(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EncodeArgs<RetinaPolicy> a;
    memset(&a, 0, sizeof(a));
    RetinaPolicy::Params& p = a.pp;
    p.n_levels = n_levels, p.n_anchors = n_anchors, p.num_classes = num_classes, p.thr = iou_thresh;
    p.num_pairs = num_pairs;
    const int ch = num_classes + 4;
    int m = 0;
    for (int l = 0; l < n_levels; ++l) {
        DH_CHECK_ARG(strides[l] > 0 && out_levels[l], "dh_retina_encode: level %d stride/pointer", l);
        p.stride[l] = strides[l];
        const int hl = static_cast<int>(static_cast<double>(pad_h) / strides[l]);
        const int wl = static_cast<int>(static_cast<double>(pad_w) / strides[l]);
        for (int an = 0; an < n_anchors; ++an, ++m) {
            p.anchor_h[l][an] = anchor_hw[(l * n_anchors + an) * 2];
            p.anchor_w[l][an] = anchor_hw[(l * n_anchors + an) * 2 + 1];
            MapDesc& md = a.tt.maps[m];
            md.rows = hl * wl;
            md.height = hl, md.width = wl, md.sub = 1, md.level = l, md.anchor = an;
            md.out = out_levels[l] + static_cast<long long>(an) * md.rows * ch;
            md.image_stride = static_cast<long long>(n_anchors) * md.rows * ch;
        }
    }
    a.tt.n_maps = m;
    a.tile_buf_bytes = finish_table(a.tt, ch, batch, h->tile_bytes);
    a.boxes = boxes, a.nbox = nbox, a.img_dim = img_dim, a.max_boxes = max_boxes;
    if (num_pairs && batch > 0) DH_CUDA(cudaMemsetAsync(num_pairs, 0, sizeof(int32_t) * batch, st));
    return launch_encode<RetinaPolicy>(h, a, st, "dh_retina_encode");
}

int dh_centernet_encode(dh_handle_t h, const float* boxes, const int32_t* nbox, const float* img_dim, int batch,
                        int max_boxes, int pad0, int pad1, int stride, int n_scales, const float* box_scales,
                        float sigma, int num_classes, int mode, float* out, int32_t* status, void* stream) {
    DH_CHECK_ARG(h && boxes && img_dim && out, "dh_centernet_encode: NULL argument");
    DH_CHECK_ARG(mode >= 0 && mode <= 2, "dh_centernet_encode: mode %d", mode);
    DH_CHECK_ARG(mode != DH_CENTERNET_ONEHOT_SCALES || (box_scales && n_scales >= 1 && n_scales <= 8),
                 "dh_centernet_encode: mode 0 needs 1..8 box_scales");
    DH_CHECK_ARG(batch >= 0 && max_boxes >= 0 && pad0 > 0 && pad1 > 0 && stride > 0, "dh_centernet_encode: bad sizes");
    if (max_boxes > DH_MAX_BOXES) return set_error(DH_ERR_CAPACITY, "dh_centernet_encode: max_boxes %d > %d", max_boxes, DH_MAX_BOXES);
    DH_CHECK_ARG(num_classes >= 1 && num_classes <= 4096, "dh_centernet_encode: num_classes %d", num_classes);
    DeviceGuard guard(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    EncodeArgs<CenterNetPolicy> a;
    memset(&a, 0, sizeof(a));
    CenterNetPolicy::Params& p = a.pp;
    p.mode = mode, p.num_classes = num_classes, p.stride = stride, p.stride_f = static_cast<float>(stride);
    p.sigma = sigma, p.pad0 = pad0, p.pad1 = pad1, p.status = status;
    p.n_scales = (mode == DH_CENTERNET_ONEHOT_SCALES) ? n_scales : 1;
    for (int n = 0; n < p.n_scales && mode == DH_CENTERNET_ONEHOT_SCALES; ++n) p.scales[n] = box_scales[n];
    MapDesc& md = a.tt.maps[0];
    a.tt.n_maps = 1;
    int hh, ww;
    if (mode == DH_CENTERNET_POWER_FALLOFF) {
        hh = static_cast<int>(static_cast<double>(pad0) / stride);
        ww = static_cast<int>(static_cast<double>(pad1) / stride);
    } else {  // reference swaps the indices (tf_centernet_resnet_s8.py:259-260)
        hh = static_cast<int>(static_cast<double>(pad1) / stride);
        ww = static_cast<int>(static_cast<double>(pad0) / stride);
    }
    const int ch = num_classes + (mode == DH_CENTERNET_POWER_FALLOFF ? 5 : 4);
    md.out = out;
    md.height = hh, md.width = ww, md.sub = p.n_scales, md.level = 0, md.anchor = 0;
    md.rows = hh * ww * p.n_scales;
    md.image_stride = static_cast<long long>(md.rows) * ch;
    a.tile_buf_bytes = finish_table(a.tt, ch, batch, h->tile_bytes);
    a.boxes = boxes, a.nbox = nbox, a.img_dim = img_dim, a.max_boxes = max_boxes;
    if (status) DH_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    return launch_encode<CenterNetPolicy>(h, a, st, "dh_centernet_encode");
}

}  // extern "C"
