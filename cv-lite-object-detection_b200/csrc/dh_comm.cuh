// The loss-scalar exchange over NVLink peer memory (the only collective of the path is one
// sum of {cls, reg, cen, n_pos} per step -- a few floats, so it is pure latency and a general-purpose collective
// library call is the slow way to do it).
//
// Every rank owns a small "mailbox" in its own HBM that all peers can address (cudaIpc handles between processes,
// cudaDeviceEnablePeerAccess inside one process).  To add its values to the sum a rank STORES them into its slot of
// every peer's mailbox over NVLink (posted writes: no round trip), each value as ONE 64-bit word that carries the step's
// sequence number in its upper half -- an aligned 8-byte store arrives whole, so a word that shows the number holds the
// value: no fence, no separate flag, and the (peer, value) stores leave from different lanes at once.  The rank then
// waits until all words of its OWN mailbox carry the number and adds them up in rank order -- every rank computes the same
// float64 sum in the same order, so the result is bit-identical everywhere and run-to-run deterministic.  Slots are double-buffered by the parity of the sequence number: a rank can only be one
// step ahead of the slowest one (it cannot finish step s before everybody has written step s, and everybody writes
// step s only after reading step s - 1), so the slot of step s + 1 never overwrites one that is still being read.
// The code below is executed by ONE warp; it is called from the last CTA of the fused loss kernel (the exchange is
// then part of the same launch as the loss) and from the stand-alone kernel behind dh_allreduce_loss.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/densehead.h"  // DH_COMM_MAX_RANKS, DH_STATUS_COMM_TIMEOUT

#define DH_COMM_MAX_VALUES 12
#define DH_COMM_SLOT_WORDS 16  // 64-bit words per (parity, rank): 12 values {sequence number : float bits} + padding = 128 bytes
#define DH_COMM_BOX_BYTES (2 * DH_COMM_MAX_RANKS * DH_COMM_SLOT_WORDS * 8)

namespace dh {

struct CommDev {
    unsigned int* box[DH_COMM_MAX_RANKS];  // box[r] = rank r's mailbox as seen from this device; box[rank] is local
    unsigned int* seq;                     // this rank's step counter (device memory, starts at 0)
    int* status;                           // the handle's status word
    int world, rank;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_relaxed_sys64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// vals[0..count) (count <= DH_COMM_MAX_VALUES; generic pointer, read by every lane, written by lanes 0 .. count-1) are
// replaced by their sum over all ranks.  All 32 lanes of the calling warp must arrive.
// `prev_seq` = the value of *c.seq, read by the caller with peer_allreduce_seq() as early as it likes (the loss kernel's last
// CTA asks for it before it reduces the partials, so the load's round trip is not part of the exchange).
__device__ __forceinline__ unsigned int peer_allreduce_seq(const CommDev& c) { return ld_relaxed_sys(c.seq); }
__device__ __forceinline__ void peer_allreduce_warp(const CommDev& c, float* vals, int count, unsigned int prev_seq) {
    const int lane = threadIdx.x & 31;
    const unsigned int s = prev_seq + 1u;
    __syncwarp();
    const unsigned int par = (s & 1u) * DH_COMM_MAX_RANKS;
    const int n_words = c.world * count;  // (peer, value) on the way out, (rank, value) on the way in
    for (int e = lane; e < n_words; e += 32) {
        const int peer = e / count, k = e - peer * count;
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(c.box[peer]) + (par + c.rank) * DH_COMM_SLOT_WORDS + k;
        st_relaxed_sys64(dst, (static_cast<unsigned long long>(s) << 32) | __float_as_uint(vals[k]));
    }
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(c.box[c.rank]) + par * DH_COMM_SLOT_WORDS;
    bool ok = true;
    unsigned long long word = 0ull;  // the last word this lane waited for: the whole exchange when n_words <= 32
    const unsigned long long t0 = global_ns();
    for (int e = lane; e < n_words && ok; e += 32) {
        const int r = e / count, k = e - r * count;
        const unsigned long long* src = mine + r * DH_COMM_SLOT_WORDS + k;
        while (static_cast<unsigned int>((word = ld_relaxed_sys64(src)) >> 32) != s) {
            if (global_ns() - t0 > c.timeout_ns) {
                ok = false;
                break;
            }
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    double acc = 0.0;
    const int k = lane < count ? lane : 0;
    if (n_words <= 32) {  // lane r * count + k holds rank r's value k
        const float v = __uint_as_float(static_cast<unsigned int>(word));
        for (int r = 0; r < c.world; ++r) acc += static_cast<double>(__shfl_sync(0xffffffffu, v, r * count + k));
    } else if (ok) {  // every word has arrived: lane k reads its column again
        for (int r = 0; r < c.world; ++r)
            acc += static_cast<double>(__uint_as_float(static_cast<unsigned int>(ld_relaxed_sys64(mine + r * DH_COMM_SLOT_WORDS + k))));
    }
    __syncwarp();  // every lane has read vals[] before lanes 0 .. count-1 overwrite it
    if (lane < count) vals[lane] = ok ? static_cast<float>(acc) : __int_as_float(0x7fc00000);
    if (lane == 0) {
        if (!ok && c.status) atomicOr(c.status, DH_STATUS_COMM_TIMEOUT);
        st_relaxed_sys(c.seq, s);
    }
    __syncwarp();
}

}  // namespace dh
