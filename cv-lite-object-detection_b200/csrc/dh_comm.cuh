// The loss-scalar exchange over NVLink peer memory (the only collective of the path is one
// sum of {cls, reg, cen, n_pos} per step -- a few floats, so it is pure latency and a general-purpose collective
// library call is the slow way to do it).
//
// Every rank owns a small "mailbox" in its own HBM that all peers can address (cudaIpc handles between processes,
// cudaDeviceEnablePeerAccess inside one process).  To add its values to the sum a rank STORES them into its slot of
// every peer's mailbox over NVLink (posted writes: no round trip), publishes them with a release store of the step's
// sequence number, then waits until all slots of its OWN mailbox carry that number and adds them up in rank order --
// every rank computes the same float64 sum in the same order, so the result is bit-identical everywhere and
// run-to-run deterministic.  Slots are double-buffered by the parity of the sequence number: a rank can only be one
// step ahead of the slowest one (it cannot finish step s before everybody has written step s, and everybody writes
// step s only after reading step s - 1), so the slot of step s + 1 never overwrites one that is still being read.
// The code below is executed by ONE warp; it is called from the last CTA of the fused loss kernel (the exchange is
// then part of the same launch as the loss) and from the stand-alone kernel behind dh_allreduce_loss.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/densehead.h"  // DH_COMM_MAX_RANKS, DH_STATUS_COMM_TIMEOUT

#define DH_COMM_MAX_VALUES 12
#define DH_COMM_SLOT_WORDS 16  // 12 values + padding + the sequence word: one 64-byte line per (parity, rank)
#define DH_COMM_BOX_BYTES (2 * DH_COMM_MAX_RANKS * DH_COMM_SLOT_WORDS * 4)

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

// vals[0..count) (count <= DH_COMM_MAX_VALUES; generic pointer, read and written by lane 0 .. world-1 / lane 0) are
// replaced by their sum over all ranks.  All 32 lanes of the calling warp must arrive.
__device__ __forceinline__ void peer_allreduce_warp(const CommDev& c, float* vals, int count) {
    const int lane = threadIdx.x & 31;
    const unsigned int s = ld_relaxed_sys(c.seq) + 1u;
    __syncwarp();
    const unsigned int par = (s & 1u) * DH_COMM_MAX_RANKS;
    if (lane < c.world) {  // lane p -> peer p
        unsigned int* dst = c.box[lane] + (par + c.rank) * DH_COMM_SLOT_WORDS;
        for (int k = 0; k < count; ++k) st_relaxed_sys(dst + k, __float_as_uint(vals[k]));
        __threadfence_system();
        st_release_sys(dst + DH_COMM_SLOT_WORDS - 1, s);
    }
    float v[DH_COMM_MAX_VALUES];
#pragma unroll
    for (int k = 0; k < DH_COMM_MAX_VALUES; ++k) v[k] = 0.f;
    bool ok = true;
    if (lane < c.world) {  // lane r <- rank r's slot of my mailbox
        const unsigned int* src = c.box[c.rank] + (par + lane) * DH_COMM_SLOT_WORDS;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(src + DH_COMM_SLOT_WORDS - 1) != s) {
            if (global_ns() - t0 > c.timeout_ns) {
                ok = false;
                break;
            }
            __nanosleep(20);
        }
#pragma unroll
        for (int k = 0; k < DH_COMM_MAX_VALUES; ++k)
            if (k < count) v[k] = __uint_as_float(ld_relaxed_sys(src + k));
    }
    ok = __all_sync(0xffffffffu, ok);
#pragma unroll
    for (int k = 0; k < DH_COMM_MAX_VALUES; ++k) {
        if (k < count) {  // warp-uniform
            double acc = 0.0;
            for (int r = 0; r < c.world; ++r) acc += static_cast<double>(__shfl_sync(0xffffffffu, v[k], r));
            if (lane == 0) vals[k] = ok ? static_cast<float>(acc) : __int_as_float(0x7fc00000);
        }
    }
    if (lane == 0) {
        if (!ok && c.status) atomicOr(c.status, DH_STATUS_COMM_TIMEOUT);
        st_relaxed_sys(c.seq, s);
    }
    __syncwarp();
}

}  // namespace dh
