// Host-side internals shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "../../include/densehead.h"

struct dh_handle_s {
    int device;
    int sm_count;
    int use_tma_store;
    int tile_bytes;
    int ctas_per_sm;
    int fused_loss_kernel;  // DH_OPT_FUSED_LOSS_KERNEL
    int nms_kernel;         // DH_OPT_NMS_KERNEL
    int fused_chunks_per_cta;  // DH_OPT_FUSED_CHUNKS_PER_CTA
    int encode_min_chunk;      // DH_OPT_ENCODE_MIN_CHUNK
    int fcos_select_mode;  // DH_OPT_FCOS_SELECT
    int nms_sort;          // DH_OPT_NMS_SORT
    int nms_filter;        // DH_OPT_NMS_FILTER
    int nms_chain;         // DH_OPT_NMS_CHAIN
    int loss_allreduce;    // DH_OPT_LOSS_ALLREDUCE
    int allreduce_mode;    // DH_OPT_ALLREDUCE
    int fused_tail;        // DH_OPT_FUSED_TAIL
    int fused_max_chunk;   // DH_OPT_FUSED_MAX_CHUNK
    int encode_kernel;     // DH_OPT_ENCODE_KERNEL
    long long launches;
    void* scratch;        // device scratch (loss partials, NMS masks), grown on demand
    size_t scratch_bytes;
    void* scratch_b;      // second arena: intermediates of the multi-kernel pipelines (detect), which call into
    size_t scratch_b_bytes;  // routines that use `scratch` themselves
    long long* phase_cycles;  // DH_OPT_PHASE_TIMING: device [8] counters (profiling aid)
    unsigned int* sched;      // ring of tile-scheduler counters (one per launch in flight)
    int sched_next;
    // Entry points hold this for the duration of the call (host side only): two threads sharing a handle cannot race
    // on the scratch arenas or the scheduler ring.  Recursive because the pipelines call other entry points.
    std::recursive_mutex mu;
    long long* trace;  // dh_set_trace: caller-owned device buffer for the fused kernel's per-CTA time stamps
    long long trace_bytes;
    int* dev_status;  // device word: sticky input-validation bits (dh_get_status)
    void* comm;  // dh::Comm* once dh_comm_* has been called (comm.cu)
};

namespace dh {

int set_error(int code, const char* fmt, ...);
// Ensure the handle's scratch holds at least `bytes`; returns device pointer or null (error set).
void* scratch(dh_handle_s* h, size_t bytes);
void* scratch_b(dh_handle_s* h, size_t bytes);
// A zeroed (stream-ordered) device counter for one kernel's dynamic tile scheduler; null on error.
unsigned int* next_sched_counter(dh_handle_s* h, cudaStream_t st);

#define DH_CHECK_ARG(cond, ...)                                      \
    do {                                                             \
        if (!(cond)) return ::dh::set_error(DH_ERR_BAD_ARG, __VA_ARGS__); \
    } while (0)

#define DH_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return ::dh::set_error(DH_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                   __FILE__, __LINE__);                                            \
    } while (0)

// Function attributes (the opt-in to more than 48 KB of dynamic shared memory) are per device: a block that sets them runs
// once for every device a handle is created on, not once per process.
//   DH_ONCE_PER_DEVICE(h) { DH_CUDA(cudaFuncSetAttribute(...)); }
#define DH_ONCE_PER_DEVICE(h) \
    for (static unsigned long long dh_once_mask_ = 0ull; !((dh_once_mask_ >> ((h)->device & 63)) & 1ull); dh_once_mask_ |= 1ull << ((h)->device & 63))

struct DeviceGuard {
    int prev;
    bool ok;
    std::recursive_mutex* mu;
    explicit DeviceGuard(int dev) : prev(-1), ok(false), mu(nullptr) { enter(dev); }
    // the form the entry points use: also serialises the calling threads on the handle
    explicit DeviceGuard(dh_handle_s* h) : prev(-1), ok(false), mu(&h->mu) {
        mu->lock();
        enter(h->device);
    }
    void enter(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && (prev == dev || cudaSetDevice(dev) == cudaSuccess)) ok = true;
    }
    ~DeviceGuard() {
        if (ok && prev >= 0) cudaSetDevice(prev);
        if (mu) mu->unlock();
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

}  // namespace dh
