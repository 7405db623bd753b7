// Tile table + per-image GT staging shared by the encode and loss kernels.
//
// Every output of the dense-head encoders is a set of row-major "maps" per image: `rows x ch`
// float32 with `ch` contiguous (FCOS level l: rows = Hl*Wl, ch = C+5; RetinaNet (level, anchor):
// rows = Hl*Wl, ch = C+4; CenterNet s8: rows = H*W*S, ch = C+4).  A map is cut into tiles of
// `rows_per_tile` rows; a tile is a contiguous byte range in HBM, which is what makes the 1-D TMA
// bulk copies possible.  Tiles are numbered image-major so a persistent CTA that walks a contiguous
// chunk of tile ids re-stages the GT boxes only when it crosses an image boundary.
#pragma once
#include "dh_common.cuh"

namespace dh {

struct MapDesc {
    float* out;              // encode target / loss target base (image 0); may be null
    const float* pred;       // prediction base (loss kernels); may be null
    long long image_stride;  // floats between the same map of consecutive images
    int rows;                // rows per image
    int height, width;       // cells
    int sub;                 // rows per cell (CenterNet s8: n_scales), else 1
    int level, anchor;
    int tile_begin, n_tiles;  // tile ids within one image
    FastDiv div_width;        // cell -> (i, j)
    FastDiv div_sub;          // row  -> cell
};

struct TileTable {
    int n_maps;
    int tiles_per_image;
    int ch;             // floats per row
    int rows_per_tile;  // multiple of 4
    int batch;
    MapDesc maps[DH_MAX_MAPS];
};

struct TileInfo {
    int b, m;       // image, map
    int r0, nrows;  // first row in the map, rows in this tile
    int level, anchor;
    int height, width, sub;
};

// A CTA walks a contiguous range of tile ids: locate the first one with a binary search, then
// advance incrementally.  `tt` must be the CTA's shared-memory copy of the table (indexed loads
// from kernel-parameter constant memory cost hundreds of cycles each when they miss).
struct TileCursor {
    int b, m, t;  // image, map, tile within the map
};
__device__ __forceinline__ void cursor_init(const TileTable& tt, long long tile, TileCursor& c) {
    c.b = static_cast<int>(tile / tt.tiles_per_image);
    const int t = static_cast<int>(tile - static_cast<long long>(c.b) * tt.tiles_per_image);
    int lo = 0, hi = tt.n_maps - 1;  // last map with tile_begin <= t
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tt.maps[mid].tile_begin <= t)
            lo = mid;
        else
            hi = mid - 1;
    }
    c.m = lo;
    c.t = t - tt.maps[lo].tile_begin;
}
__device__ __forceinline__ void cursor_next(const TileTable& tt, TileCursor& c) {
    if (++c.t == tt.maps[c.m].n_tiles) {
        c.t = 0;
        if (++c.m == tt.n_maps) c.m = 0, ++c.b;
    }
}
__device__ __forceinline__ void cursor_info(const TileTable& tt, const TileCursor& c, TileInfo& ti) {
    const MapDesc& md = tt.maps[c.m];
    ti.b = c.b;
    ti.m = c.m;
    ti.r0 = c.t * tt.rows_per_tile;
    ti.nrows = min(tt.rows_per_tile, md.rows - ti.r0);
    ti.level = md.level;
    ti.anchor = md.anchor;
    ti.height = md.height;
    ti.width = md.width;
    ti.sub = md.sub;
}
// Copy a kernel-parameter struct into shared memory (all threads; caller syncs).
template <class T>
__device__ __forceinline__ void copy_args_to_smem(const T& src, T* dst) {
    static_assert(sizeof(T) % 4 == 0, "argument structs are 4-byte granular");
    const int* s = reinterpret_cast<const int*>(&src);
    int* d = reinterpret_cast<int*>(dst);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(T) / 4); i += blockDim.x) d[i] = s[i];
}

// Stage the raw GT rows of image `b` ([max_boxes, 5] float32) into shared memory.  Uses one TMA
// bulk copy when the source is 16-byte aligned and sized, otherwise cooperative loads.
// Must be called by all threads of the CTA; ends with a __syncthreads().
__device__ __forceinline__ int stage_boxes(const float* __restrict__ boxes, const int* __restrict__ nbox, int b,
                                           int max_boxes, int box_cap, float* raw /*[box_cap*5]*/, uint64_t* bar,
                                           uint32_t& bar_parity) {
    const float* src = boxes + static_cast<long long>(b) * max_boxes * 5;
    int n = nbox ? nbox[b] : max_boxes;
    n = max(0, min(n, min(max_boxes, box_cap)));
    const uint32_t bytes = static_cast<uint32_t>(n) * 20u;
    const uint32_t bytes16 = (bytes + 15u) & ~15u;
    const bool tma_ok = n > 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0 &&
                        bytes16 <= static_cast<uint32_t>(min(max_boxes, box_cap)) * 20u;
    if (tma_ok) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, bytes16);
            bulk_g2s(raw, src, bytes16, bar);
        }
        mbar_wait(bar, bar_parity);
        bar_parity ^= 1u;
    } else {
        for (int e = threadIdx.x; e < n * 5; e += blockDim.x) raw[e] = src[e];
    }
    __syncthreads();
    return n;
}

}  // namespace dh
