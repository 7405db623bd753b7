// The encode "tile streamer": one persistent kernel template for every target encoder.
//
// Target maps are > 99 % zeros, so the job is a memset with sparse patches and the only thing that
// matters is how the zeros reach HBM.  Each CTA owns kStages zero-initialised tile buffers in
// shared memory.  Per tile: (1) the few rows that receive targets are written by the thread that
// owns the row (atomic-free gather, see dh_policies.cuh); (2) one elected thread hands the whole
// tile to the TMA engine as a single 1-D bulk store (cp.async.bulk.global.shared::cta), so the SM
// issues no per-element store instructions; (3) when the buffer comes round again the owner
// threads re-zero only the rows they dirtied.  Unaligned tile edges (< 16 B) are written with
// scalar stores.  A fallback path (`use_tma_store == 0`) copies the tile out with 128-bit
// st.global instead; it exists for A/B measurements and as a safety net.
#pragma once
#include "dh_policies.cuh"

namespace dh {

constexpr int kStages = 2;

template <class P>
struct EncodeArgs {
    TileTable tt;
    typename P::Params pp;
    const float* boxes;    // [B, max_boxes, 5]
    const int* nbox;       // [B] or null (= max_boxes)
    const float* img_dim;  // [B, 2] (H, W) unpadded content size
    int max_boxes;
    int tile_buf_bytes;  // shared-memory bytes per stage (>= rows_per_tile*ch*4 + 16, multiple of 128)
    int use_tma_store;
};

struct EncodeSmemLayout {
    int stage_off, rec_off, raw_off, cand_off, misc_off, args_off, total;
};
template <class P>
__host__ __device__ inline EncodeSmemLayout encode_smem_layout(int tile_buf_bytes) {
    EncodeSmemLayout l;
    l.stage_off = 0;
    l.rec_off = kStages * tile_buf_bytes;
    l.raw_off = l.rec_off + ((static_cast<int>(sizeof(typename P::Rec)) * DH_MAX_BOXES + 127) & ~127);
    l.cand_off = l.raw_off + DH_MAX_BOXES * 5 * 4;  // 5120, multiple of 128
    l.misc_off = l.cand_off + DH_MAX_BOXES * 2;     // 512
    l.args_off = l.misc_off + 128;
    l.total = l.args_off + ((static_cast<int>(sizeof(EncodeArgs<P>)) + 127) & ~127);
    return l;
}

template <class P>
__global__ void __launch_bounds__(DH_THREADS) encode_kernel(const __grid_constant__ EncodeArgs<P> ga) {
    extern __shared__ __align__(128) unsigned char smem[];
    const EncodeSmemLayout lay = encode_smem_layout<P>(ga.tile_buf_bytes);
    // work from a shared-memory copy of the arguments: the tile table and the policy tables are
    // indexed dynamically, which is slow from kernel-parameter constant memory
    const EncodeArgs<P>& a = *reinterpret_cast<const EncodeArgs<P>*>(smem + lay.args_off);
    copy_args_to_smem(ga, reinterpret_cast<EncodeArgs<P>*>(smem + lay.args_off));
    typename P::Rec* recs = reinterpret_cast<typename P::Rec*>(smem + lay.rec_off);
    float* raw = reinterpret_cast<float*>(smem + lay.raw_off);
    unsigned short* cand = reinterpret_cast<unsigned short*>(smem + lay.cand_off);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lay.misc_off);
    int* wcount = reinterpret_cast<int*>(smem + lay.misc_off + 16);  // [8] per-warp candidate counts

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ch = ga.tt.ch;
    const long long total_tiles = static_cast<long long>(ga.tt.batch) * ga.tt.tiles_per_image;
    const long long t_begin = total_tiles * blockIdx.x / gridDim.x;
    const long long t_end = total_tiles * (blockIdx.x + 1) / gridDim.x;
    if (t_begin >= t_end) return;

    // zero the stage buffers once; afterwards only dirtied rows are re-zeroed
    {
        float4* z = reinterpret_cast<float4*>(smem + lay.stage_off);
        const int n4 = kStages * ga.tile_buf_bytes / 16;
        for (int e = tid; e < n4; e += DH_THREADS) z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
    }
    __syncthreads();

    uint32_t bar_parity = 0;
    static_assert(kStages == 2, "per-stage registers below are written out for two stages");
    uint32_t dirty0 = 0u, dirty1 = 0u;  // rows this thread dirtied in each stage buffer
    int mis0 = 0, mis1 = 0;             // float offset of the tile inside its stage buffer (global 16-B phase)
    int cur_img = -1, n_boxes = 0;
    int it = 0;
    TileCursor cur;
    cursor_init(a.tt, t_begin, cur);
    for (long long tile = t_begin; tile < t_end; ++tile, ++it, cursor_next(a.tt, cur)) {
        TileInfo ti;
        cursor_info(a.tt, cur, ti);
        const MapDesc& md = a.tt.maps[ti.m];

        if (ti.b != cur_img) {  // crossed an image boundary: stage its GT rows, build records
            __syncthreads();    // nobody still reads recs/raw of the previous image
            n_boxes = stage_boxes(a.boxes, a.nbox, ti.b, a.max_boxes, raw, bar, bar_parity);
            const float hi = a.img_dim[2 * ti.b], wi = a.img_dim[2 * ti.b + 1];
            if (tid < n_boxes) P::make_record(a.pp, raw + 5 * tid, hi, wi, tid, recs[tid]);
            __syncthreads();
            // the CTA that owns the image's first tile publishes per-image side outputs
            if (tile == static_cast<long long>(ti.b) * a.tt.tiles_per_image) P::image_prologue(a.pp, recs, n_boxes, ti.b);
            cur_img = ti.b;
        }

        // candidate list of this tile (ballot compaction keeps ascending GT order)
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

        const int s = it % kStages;
        float* buf = reinterpret_cast<float*>(smem + lay.stage_off + s * a.tile_buf_bytes);
        if (it >= kStages) {  // buffer reuse: its previous bulk store must have finished reading
            if (tid == 0) bulk_wait_read<kStages - 1>();
            __syncthreads();
            uint32_t d = s ? dirty1 : dirty0;
            const int old_mis = s ? mis1 : mis0;
            for (int k = 0; d; ++k, d >>= 1)
                if (d & 1u) {
                    float* row = buf + old_mis + (tid + k * DH_THREADS) * ch;
                    for (int c = 0; c < ch; ++c) row[c] = 0.f;
                }
        }
        __syncthreads();  // cand[] complete, buffer clean

        // destination of this tile and its 16-byte phase
        float* gdst = md.out + static_cast<long long>(ti.b) * md.image_stride + static_cast<long long>(ti.r0) * ch;
        const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(gdst) >> 2) & 3u);
        if (s) mis1 = mis; else mis0 = mis;
        float* tile_smem = buf + mis;
        uint32_t dmask = 0u;
        if (ncand > 0) {  // block-uniform
            int painted = 0;
            for (int k = 0, r = tid; r < ti.nrows; r += DH_THREADS, ++k) {
                const int n = P::emit_row(a.pp, ti, md, ti.r0 + r, tile_smem + r * ch, recs, cand, ncand);
                if (n > 0) dmask |= (1u << k);
                painted += n;
            }
            P::tile_epilogue(a.pp, ti, painted);
        }
        if (s) dirty1 = dmask; else dirty0 = dmask;

        const int nfl = ti.nrows * ch;           // floats in the tile
        const int head = (4 - mis) & 3;          // floats before the first 16-B boundary
        const int body = (nfl - head) > 0 ? ((nfl - head) & ~3) : 0;
        if (a.use_tma_store) {
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                if (body > 0) bulk_s2g(gdst + head, tile_smem + head, static_cast<uint32_t>(body) * 4u);
                bulk_commit();
            }
            // ragged edges (at most 3 + 3 floats)
            if (tid >= 32 && tid < 32 + head && tid - 32 < nfl) gdst[tid - 32] = tile_smem[tid - 32];
            const int tail0 = head + body;
            if (tid >= 64 && tid - 64 < nfl - tail0) gdst[tail0 + tid - 64] = tile_smem[tail0 + tid - 64];
        } else {
            __syncthreads();
            for (int e = tid; e < head && e < nfl; e += DH_THREADS) gdst[e] = tile_smem[e];
            const float4* src4 = reinterpret_cast<const float4*>(tile_smem + head);
            float4* dst4 = reinterpret_cast<float4*>(gdst + head);
            for (int e = tid; e < body / 4; e += DH_THREADS) dst4[e] = src4[e];
            for (int e = head + body + tid; e < nfl; e += DH_THREADS) gdst[e] = tile_smem[e];
            if (tid == 0) bulk_commit();  // keep group accounting uniform (empty group)
        }
    }
    if (tid == 0) bulk_wait_read<0>();
}

}  // namespace dh
