// The direct-store target encoder: the same policies as the tile streamer (dh_encode_kernel.cuh), no shared-memory
// tile.  A target map is > 99 % zeros; here every CTA owns a contiguous run of rows of ONE image, writes that byte
// range as zeros with plain 128-bit stores (fire and forget: the SM never waits for them), and then the few rows that
// receive targets are written again, by the thread that owns the row (gather policies) or the box (centre-cell
// policies).  A block barrier between the two orders them (same addresses, different threads).  Per CTA the latency
// chain is one global round trip (thread k asks for GT row k, the count and the image size, then issues its zero stores
// while they travel) -> records -> candidates -> patch; there are no stage buffers to initialise or recycle, no
// mbarrier hand-off per tile, no scheduler counter, no shared-memory copy of the arguments and no drain at the end,
// which is what a 4 MB problem (FCOS, 8 images) spends its time on in the streamer.  Work is cut statically: CTA c
// takes chunk c (image-aligned; the launcher sizes chunks for one wave of resident CTAs, 8..128 KB each).
#pragma once
#include "dh_encode_kernel.cuh"

namespace dh {

struct DirectSmemLayout {
    int rec_off, raw_off, cand_off, cand2_off, dlist_off, misc_off, total;
};
template <class P>
__host__ __device__ inline DirectSmemLayout direct_smem_layout(int box_cap) {
    DirectSmemLayout l;
    l.rec_off = 0;
    l.raw_off = (static_cast<int>(sizeof(typename P::Rec)) * box_cap + 127) & ~127;
    l.cand_off = l.raw_off + ((box_cap * 20 + 127) & ~127);
    l.cand2_off = l.cand_off + DH_THREADS * 2;
    l.dlist_off = l.cand2_off + ((box_cap * 2 + 127) & ~127);
    l.misc_off = l.dlist_off + ((box_cap * 2 + 127) & ~127);
    l.total = l.misc_off + 128;
    return l;
}

// zeros over floats [0, n) at `dst` (any 4-byte alignment): scalar head up to the first 16-byte boundary, 128-bit body
__device__ __forceinline__ void zero_fill(float* __restrict__ dst, long long n) {
    const int tid = threadIdx.x;
    const int head = static_cast<int>((4u - ((reinterpret_cast<uintptr_t>(dst) >> 2) & 3u)) & 3u);
    const long long h = head < n ? head : n;
    if (tid < h) dst[tid] = 0.f;
    const long long body4 = (n - h) >> 2;
    float4* __restrict__ d4 = reinterpret_cast<float4*>(dst + h);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    long long e = tid;
#pragma unroll 4
    for (; e < body4; e += DH_THREADS) d4[e] = z;
    const long long tail0 = h + (body4 << 2);
    if (tail0 + tid < n) dst[tail0 + tid] = 0.f;
}

// Segment `seg`-th run of the chunk's tiles inside one map: rows [ti.r0, ti.r0 + ti.nrows) of map ti.m.  Advances the cursor.
__device__ __forceinline__ int next_segment(const TileTable& tt, TileCursor& cur, int tiles_left, TileInfo& ti) {
    const MapDesc& md = tt.maps[cur.m];
    const int nseg = min(md.n_tiles - cur.t, tiles_left);
    cursor_info(tt, cur, ti);
    ti.nrows = min(nseg * tt.rows_per_tile, md.rows - ti.r0);
    cur.t += nseg;
    if (cur.t == md.n_tiles) {
        cur.t = 0;
        if (++cur.m == tt.n_maps) cur.m = 0, ++cur.b;
    }
    return nseg;
}

static_assert(DH_MAX_BOXES <= DH_THREADS, "thread k builds the record of box k from its own registers");

template <class P>
__global__ void __launch_bounds__(DH_THREADS) encode_direct_kernel(const __grid_constant__ EncodeArgs<P> a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const DirectSmemLayout lay = direct_smem_layout<P>(a.box_cap);
    // (the arguments are read where they lie, in the kernel-parameter constant bank: every access below is uniform
    // across the CTA and a CTA touches one or two map descriptors, so a shared-memory copy of the 6 KB struct -- what
    // the persistent kernels work from -- would only lengthen the latency chain)
    typename P::Rec* recs = reinterpret_cast<typename P::Rec*>(smem + lay.rec_off);
    unsigned short* cand = reinterpret_cast<unsigned short*>(smem + lay.cand_off);
    unsigned short* cand_dense = reinterpret_cast<unsigned short*>(smem + lay.cand2_off);
    unsigned short* dlist = reinterpret_cast<unsigned short*>(smem + lay.dlist_off);
    int* wcnt = reinterpret_cast<int*>(smem + lay.misc_off);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ch = a.tt.ch, tpi = a.tt.tiles_per_image;
    const long long chunk = blockIdx.x;
    // chunk -> images and tile range (chunks_per_image > 1 splits an image, else a chunk is images_per_chunk whole images)
    int img0, img1, t_first, t_last;
    if (a.chunks_per_image > 1) {
        img0 = static_cast<int>(chunk / a.chunks_per_image);
        img1 = img0 + 1;
        const int sub = static_cast<int>(chunk - static_cast<long long>(img0) * a.chunks_per_image);
        t_first = sub * a.chunk_tiles;
        t_last = min(t_first + a.chunk_tiles, tpi);
    } else {
        img0 = static_cast<int>(chunk * a.images_per_chunk);
        img1 = min(img0 + a.images_per_chunk, a.tt.batch);
        t_first = 0, t_last = tpi;
    }
#pragma unroll 1
    for (int img = img0; img < img1; ++img) {
        // ---- 1. request GT row `tid`, the count and the image size (registers; nothing waits yet) -----------------
        const float* __restrict__ src = a.boxes + (static_cast<long long>(img) * a.max_boxes + tid) * 5;
        float g[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (tid < a.max_boxes) {
#pragma unroll
            for (int k = 0; k < 5; ++k) g[k] = __ldg(src + k);
        }
        int n_boxes = a.nbox ? __ldg(a.nbox + img) : a.max_boxes;
        const float hi = __ldg(a.img_dim + 2 * img), wi = __ldg(a.img_dim + 2 * img + 1);
        // ---- 2. zeros over every row this CTA owns of the image: fire and forget ---------------------------------
        {
            TileCursor cur;
            cursor_init(a.tt, static_cast<long long>(img) * tpi + t_first, cur);
            for (int tile = t_first; tile < t_last;) {
                TileInfo ti;
                tile += next_segment(a.tt, cur, t_last - tile, ti);
                const MapDesc& md = a.tt.maps[ti.m];
                zero_fill(md.out + static_cast<long long>(img) * md.image_stride + static_cast<long long>(ti.r0) * ch,
                          static_cast<long long>(ti.nrows) * ch);
            }
        }
        // ---- 3. records -------------------------------------------------------------------------------------------
        if (img != img0) __syncthreads();  // nobody still reads the records of the previous image
        n_boxes = max(0, min(n_boxes, min(a.max_boxes, a.box_cap)));
        if (tid < n_boxes) P::make_record(a.pp, g, hi, wi, tid, recs[tid]);
        __syncthreads();  // also orders the zero stores above before the patches below (same addresses, other threads)
        if (t_first == 0) P::image_prologue(a.pp, recs, n_boxes, img);
        // ---- 4. the rows that receive targets ----------------------------------------------------------------------
        TileCursor cur;
        cursor_init(a.tt, static_cast<long long>(img) * tpi + t_first, cur);
#pragma unroll 1
        for (int tile = t_first; tile < t_last;) {
            TileInfo ti;
            tile += next_segment(a.tt, cur, t_last - tile, ti);
            const MapDesc& md = a.tt.maps[ti.m];
            float* gdst = md.out + static_cast<long long>(img) * md.image_stride + static_cast<long long>(ti.r0) * ch;
            build_candidates<P>(a.pp, recs, n_boxes, ti, md, cand, wcnt, warp, lane);
            __syncthreads();
            const int ncand = total_candidates(wcnt);
            if (ncand > 0) {  // block-uniform
                int base = 0;
                for (int w = 0; w < warp; ++w) base += wcnt[w];
                if (lane < wcnt[warp]) cand_dense[base + lane] = cand[warp * 32 + lane];
                __syncthreads();
                bool scattered = false;
                if constexpr (P::kScatter) {
                    if (P::use_scatter(a.pp)) {
                        scattered = true;
                        P::emit_tile(a.pp, ti, md, gdst, ch, recs, cand_dense, ncand, dlist);
                    }
                }
                if (!scattered) {
                    int painted = 0;
                    for (int r = tid; r < ti.nrows; r += DH_THREADS)
                        painted += P::emit_row(a.pp, ti, md, ti.r0 + r, gdst + static_cast<long long>(r) * ch, recs, cand_dense, ncand);
                    P::tile_epilogue(a.pp, ti, painted);
                }
            }
            if (tile < t_last) __syncthreads();  // cand / wcnt / cand_dense are rebuilt for the next segment
        }
    }
}

}  // namespace dh
