// Common device helpers for the dense-head kernels (sm_100a only).
//
//   * TMA 1-D bulk copies (cp.async.bulk) global->shared with an mbarrier and shared->global as a
//     bulk group: the encoders stream mostly-zero target tiles out of shared memory with one
//     elected thread, so the SM issues almost no store instructions; the loss kernels stream
//     prediction/target tiles in the same way.
//   * exact-arithmetic wrappers (__fmul_rn & co) so that the integer decisions the reference takes
//     from float32 expressions (cell indices, IoU > thr) can never be changed by FMA contraction.
//   * warp/block reductions, magic-number division.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DH_MAX_BOXES 256   // GT boxes per image staged in shared memory
#define DH_MAX_MAPS 48     // output maps per image (RetinaNet: 5 levels x 9 anchors)
#define DH_MAX_LEVELS 8
#define DH_THREADS 256

namespace dh {

// ------------------------------------------------------------------ exact float32 ops
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
// Python int(float): truncate toward zero
__device__ __forceinline__ int trunc_i(float v) { return __float2int_rz(v); }

// ------------------------------------------------------------------ shared-memory address
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ------------------------------------------------------------------ TMA 1-D bulk copies
// global -> shared; completion is signalled on `bar` (complete_tx::bytes).  16-byte aligned
// addresses, size a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global as part of the calling thread's current bulk group.
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                     __cvta_generic_to_global(dst_gmem)),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA) that will read them
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ division by a runtime constant
struct FastDiv {
    uint32_t d, mul, shr;  // n / d == umulhi(n, mul) >> shr   for 0 <= n < 2^31 (d >= 1)
};
__host__ __device__ inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    if (d <= 1) {
        f.mul = 0;
        f.shr = 0;
        return f;
    }
    uint32_t l = 0;
    while ((1u << l) < d) ++l;  // ceil(log2 d)
    uint64_t m = ((1ull << (31 + l)) + d - 1) / d;
    f.mul = static_cast<uint32_t>(m);  // fits: m < 2^32 because 2^(31+l)/d < 2^32
    f.shr = l - 1;
    return f;
}
__device__ __forceinline__ uint32_t fdiv_u32(uint32_t n, const FastDiv& f) {
    return f.d <= 1 ? n : (__umulhi(n, f.mul) >> f.shr);
}

// ------------------------------------------------------------------ dynamic tile scheduler counter
// sched[0] hands out chunk ids, sched[1] counts finished CTAs; the last CTA to finish zeroes both, so the line is
// ready for the next launch that gets it (launches on one stream are ordered; the host rotates lines between launches).
// Called by one thread per CTA after the CTA's last use of sched[0].
__device__ __forceinline__ void sched_release(unsigned int* sched) {
    __threadfence();
    const unsigned done = atomicAdd(sched + 1, 1u);
    if (done == gridDim.x - 1) {
        sched[0] = 0u;
        sched[1] = 0u;
        __threadfence();
    }
}

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dh
