// The encode "tile streamer": one persistent kernel template for every target encoder.
//
// Target maps are > 99 % zeros, so the job is a memset with sparse patches and the only thing that
// matters is how the zeros reach HBM.  Each CTA owns kStages zero-initialised tile buffers in
// shared memory.  Per tile: (1) warp 0 compacts the GT boxes that can touch the tile into a
// candidate list (ballot, no atomics); (2) the rows that receive targets are written by the thread
// that owns the row (gather policies) or by the thread that owns the box (scatter policies, with a
// pairwise "who paints last" check) -- atomic-free either way, see dh_policies.cuh; (3) one elected
// thread hands the whole tile to the TMA engine as a single 1-D bulk store
// (cp.async.bulk.global.shared::cta), so the SM issues no per-element store instructions; (4) when
// the buffer comes round again only the dirtied rows are re-zeroed.  Unaligned tile edges (< 16 B)
// are written with scalar stores.  A fallback path (`use_tma_store == 0`) copies the tile out with
// 128-bit st.global instead; it exists for A/B measurements and as a safety net.
//
// The per-tile SM work is a chain of short latencies (two block barriers, a few shared-memory round
// trips), not throughput, so the kernel is sized for many small CTAs per SM: their bulk stores
// overlap each other's bookkeeping.
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
    int box_cap;         // shared-memory capacity in boxes: max_boxes rounded up to 32
    int tile_buf_bytes;  // shared-memory bytes per stage (>= rows_per_tile*ch*4 + 16, multiple of 128)
    int use_tma_store;
    int chunk_tiles;          // dynamic scheduler: tiles per chunk (chunks never straddle an image when cpi > 1)
    int chunks_per_image;     // > 1: an image is cut into this many chunks; 1: a chunk is images_per_chunk whole images
    int images_per_chunk;
    long long n_chunks;
    unsigned int* sched;      // device counter, zeroed by the launcher before every launch
    long long* phase_cycles;  // profiling aid: CTA 0 adds its per-phase clock64 totals here ([8]); usually null
};

struct EncodeSmemLayout {
    int stage_off, rec_off, raw_off, cand_off, cand2_off, dlist_off, dlist_stride, misc_off, args_off, total;
};
template <class P>
__host__ __device__ inline EncodeSmemLayout encode_smem_layout(int tile_buf_bytes, int box_cap) {
    EncodeSmemLayout l;
    l.stage_off = 0;
    l.rec_off = kStages * tile_buf_bytes;
    l.raw_off = l.rec_off + ((static_cast<int>(sizeof(typename P::Rec)) * box_cap + 127) & ~127);
    l.cand_off = l.raw_off + ((box_cap * 20 + 127) & ~127);
    l.cand2_off = l.cand_off + DH_THREADS * 2;  // per-warp segments cover DH_THREADS boxes
    l.dlist_off = l.cand2_off + ((box_cap * 2 + 127) & ~127);
    l.dlist_stride = (box_cap * 2 + 127) & ~127;
    l.misc_off = l.dlist_off + (P::kScatter ? kStages * l.dlist_stride : 0);
    l.args_off = l.misc_off + 128;
    l.total = l.args_off + ((static_cast<int>(sizeof(EncodeArgs<P>)) + 127) & ~127);
    return l;
}

// Candidate list of a tile: warp w tests boxes [32w, 32w+32) and compacts its hits (ballot, ascending
// GT order) into its own 32-entry segment of `cand`; `wcnt[w]` is the segment length.  No block-level
// prefix is needed: consumers walk the segments in warp order.
template <class P>
__device__ __forceinline__ void build_candidates(const typename P::Params& pp, const typename P::Rec* recs, int n_boxes,
                                                 const TileInfo& ti, const MapDesc& md, unsigned short* cand, int* wcnt,
                                                 int warp, int lane) {
    const int k = warp * 32 + lane;
    const bool hit = k < n_boxes && P::tile_hit(pp, recs[k], ti, md);
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (hit) cand[warp * 32 + __popc(bal & ((1u << lane) - 1u))] = static_cast<unsigned short>(k);
    if (lane == 0) wcnt[warp] = __popc(bal);
}
// after the barrier: squeeze the per-warp segments into one dense list (every thread computes the same
// prefix; thread `q` moves entry q).  Returns the total.
__device__ __forceinline__ int total_candidates(const int* wcnt) {
    int n = 0;
#pragma unroll
    for (int w = 0; w < DH_THREADS / 32; ++w) n += wcnt[w];
    return n;
}

template <class P>
__global__ void __launch_bounds__(DH_THREADS) encode_kernel(const __grid_constant__ EncodeArgs<P> ga) {
    extern __shared__ __align__(128) unsigned char smem[];
    const EncodeSmemLayout lay = encode_smem_layout<P>(ga.tile_buf_bytes, ga.box_cap);
    // work from a shared-memory copy of the arguments: the tile table and the policy tables are
    // indexed dynamically, which is slow from kernel-parameter constant memory
    const EncodeArgs<P>& a = *reinterpret_cast<const EncodeArgs<P>*>(smem + lay.args_off);
    copy_args_to_smem(ga, reinterpret_cast<EncodeArgs<P>*>(smem + lay.args_off));
    typename P::Rec* recs = reinterpret_cast<typename P::Rec*>(smem + lay.rec_off);
    float* raw = reinterpret_cast<float*>(smem + lay.raw_off);
    unsigned short* cand = reinterpret_cast<unsigned short*>(smem + lay.cand_off);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lay.misc_off);
    int* dcount = reinterpret_cast<int*>(smem + lay.misc_off + 16);   // [kStages] scatter mode: dirty rows per stage
    long long* next_chunk = reinterpret_cast<long long*>(smem + lay.misc_off + 24);
    int* wcnt = reinterpret_cast<int*>(smem + lay.misc_off + 32);     // [8] per-warp candidate counts
    unsigned short* cand_dense = reinterpret_cast<unsigned short*>(smem + lay.cand2_off);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ch = ga.tt.ch;
    const long long total_tiles = static_cast<long long>(ga.tt.batch) * ga.tt.tiles_per_image;
    // dynamic scheduler: CTA b starts with chunk b and fetches further chunk ids from the device counter
    // (tiles differ 2x in cost, a static split tails).  Chunks are image-aligned so that the GT staging
    // is paid once per chunk at most.
    const long long n_chunks = ga.n_chunks;
    const int tpi = ga.tt.tiles_per_image;
    long long chunk = blockIdx.x;
    if (chunk >= n_chunks) {  // (the launchers clamp the grid to the chunk count; kept for safety)
        if (threadIdx.x == 0) sched_release(ga.sched);
        return;
    }

    // zero the stage buffers once; afterwards only dirtied rows are re-zeroed
    {
        float4* z = reinterpret_cast<float4*>(smem + lay.stage_off);
        const int n4 = kStages * ga.tile_buf_bytes / 16;
        for (int e = tid; e < n4; e += DH_THREADS) z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
        dcount[0] = dcount[1] = 0;
    }
    __syncthreads();

    uint32_t bar_parity = 0;
    static_assert(kStages == 2, "per-stage registers below are written out for two stages");
    uint32_t dirty0 = 0u, dirty1 = 0u;  // gather mode: rows this thread dirtied in each stage buffer
    int mis0 = 0, mis1 = 0;             // float offset of the tile inside its stage buffer (global 16-B phase)
    int cur_img = -1, n_boxes = 0;
    int it = 0;
    const bool prof = ga.phase_cycles != nullptr && blockIdx.x == 0 && tid == 0;
    long long pc[6] = {0, 0, 0, 0, 0, 0}, pt = prof ? clock64() : 0;
    unsigned long long gt0 = 0;
    if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#define DH_PHASE(i)                      \
    if (prof) {                          \
        const long long now = clock64(); \
        pc[i] += now - pt;               \
        pt = now;                        \
    }
    for (; chunk < n_chunks;) {
    long long t_begin, t_end;
    if (ga.chunks_per_image > 1) {
        const long long img = chunk / ga.chunks_per_image;
        const int sub = static_cast<int>(chunk - img * ga.chunks_per_image);
        t_begin = img * tpi + static_cast<long long>(sub) * ga.chunk_tiles;
        t_end = min(t_begin + ga.chunk_tiles, (img + 1) * tpi);
    } else {
        t_begin = chunk * ga.images_per_chunk * tpi;
        t_end = min(t_begin + static_cast<long long>(ga.images_per_chunk) * tpi, total_tiles);
    }
    if (tid == 0) *next_chunk = static_cast<long long>(atomicAdd(ga.sched, 1u)) + gridDim.x;  // prefetch the next chunk id
    TileCursor cur;
    cursor_init(a.tt, t_begin, cur);
    for (long long tile = t_begin; tile < t_end; ++tile, ++it, cursor_next(a.tt, cur)) {
        TileInfo ti;
        cursor_info(a.tt, cur, ti);
        const MapDesc& md = a.tt.maps[ti.m];

        if (ti.b != cur_img) {  // crossed an image boundary: stage its GT rows, build records
            __syncthreads();    // nobody still reads recs/raw of the previous image
            n_boxes = stage_boxes(a.boxes, a.nbox, ti.b, a.max_boxes, a.box_cap, raw, bar, bar_parity);
            const float hi = a.img_dim[2 * ti.b], wi = a.img_dim[2 * ti.b + 1];
            if (tid < n_boxes) P::make_record(a.pp, raw + 5 * tid, hi, wi, tid, recs[tid]);
            __syncthreads();
            // the CTA that owns the image's first tile publishes per-image side outputs
            if (tile == static_cast<long long>(ti.b) * a.tt.tiles_per_image) P::image_prologue(a.pp, recs, n_boxes, ti.b);
            cur_img = ti.b;
        }
        DH_PHASE(0)

        const int s = it & 1;
        float* buf = reinterpret_cast<float*>(smem + lay.stage_off + s * a.tile_buf_bytes);
        const int nd = P::kScatter ? dcount[s] : 0;  // written two tiles ago; read before tid 0 can overwrite it
        build_candidates<P>(a.pp, recs, n_boxes, ti, md, cand, wcnt, warp, lane);
        if (tid == 0 && it >= kStages) bulk_wait_read<kStages - 1>();  // this buffer's previous bulk store has read it
        __syncthreads();  // barrier A: candidate segments complete, stage buffer free
        const int ncand = total_candidates(wcnt);
        if (ncand > 0) {  // squeeze the segments into a dense list (uniform branch)
            int base = 0;
            for (int w = 0; w < warp; ++w) base += wcnt[w];
            if (lane < wcnt[warp]) cand_dense[base + lane] = cand[warp * 32 + lane];
        }

        // destination of this tile and its 16-byte phase
        float* gdst = md.out + static_cast<long long>(ti.b) * md.image_stride + static_cast<long long>(ti.r0) * ch;
        const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(gdst) >> 2) & 3u);
        const int old_mis = s ? mis1 : mis0;
        float* tile_smem = buf + mis;
        unsigned short* dl = reinterpret_cast<unsigned short*>(smem + lay.dlist_off + s * lay.dlist_stride);
        // re-zero what the previous use of this buffer dirtied
        bool resync = old_mis != mis;  // row slots shift with the 16-B phase (uniform, rare)
        if constexpr (P::kScatter) {
            for (int q = warp; q < nd; q += DH_THREADS / 32) {  // scatter mode: rows are not owned by fixed threads
                float* row = buf + old_mis + static_cast<int>(dl[q]) * ch;
                for (int c = lane; c < ch; c += 32) row[c] = 0.f;
            }
            resync = resync || nd > 0;
        }
        {
            uint32_t d = s ? dirty1 : dirty0;
            for (int k = 0; d; ++k, d >>= 1)
                if (d & 1u) {
                    float* row = buf + old_mis + (tid + k * DH_THREADS) * ch;
                    for (int c = 0; c < ch; ++c) row[c] = 0.f;
                }
        }
        if (resync || ncand > 0) __syncthreads();  // (uniform) dense candidate list + re-zeroed rows visible
        if (s) mis1 = mis; else mis0 = mis;
        DH_PHASE(1)

        uint32_t dmask = 0u;
        bool scattered = false;
        if constexpr (P::kScatter) {
            if (P::use_scatter(a.pp)) {
                scattered = true;
                P::emit_tile(a.pp, ti, md, tile_smem, ch, recs, cand_dense, ncand, dl);
                if (tid == 0) dcount[s] = ncand;
            } else if (tid == 0) {
                dcount[s] = 0;
            }
        }
        if (!scattered && ncand > 0) {  // block-uniform
            int painted = 0;
            for (int k = 0, r = tid; r < ti.nrows; r += DH_THREADS, ++k) {
                const int n = P::emit_row(a.pp, ti, md, ti.r0 + r, tile_smem + r * ch, recs, cand_dense, ncand);
                if (n > 0) dmask |= (1u << k);
                painted += n;
            }
            P::tile_epilogue(a.pp, ti, painted);
        }
        if (s) dirty1 = dmask; else dirty0 = dmask;
        DH_PHASE(2)

        const int nfl = ti.nrows * ch;   // floats in the tile
        const int head = (4 - mis) & 3;  // floats before the first 16-B boundary
        const int body = (nfl - head) > 0 ? ((nfl - head) & ~3) : 0;
        if (a.use_tma_store) {
            fence_async_smem();
            __syncthreads();  // barrier B: the tile is complete and visible to the async proxy
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
        DH_PHASE(3)
    }
    __syncthreads();  // every thread is past the tile loop; the prefetched chunk id is visible
    chunk = *next_chunk;
    __syncthreads();  // ... and read by all before thread 0 overwrites it
    }
    if (tid == 0) {
        bulk_wait_read<0>();
        sched_release(ga.sched);
    }
    if (prof) {
        DH_PHASE(4)
        for (int i = 0; i < 5; ++i)
            atomicAdd(reinterpret_cast<unsigned long long*>(ga.phase_cycles) + i, static_cast<unsigned long long>(pc[i]));
        atomicAdd(reinterpret_cast<unsigned long long*>(ga.phase_cycles) + 5, static_cast<unsigned long long>(it));
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        atomicAdd(reinterpret_cast<unsigned long long*>(ga.phase_cycles) + 6, gt1 - gt0);  // nanoseconds
    }
#undef DH_PHASE
}

}  // namespace dh
