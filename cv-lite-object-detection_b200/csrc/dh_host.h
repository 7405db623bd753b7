// Host-side internals shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/densehead.h"

struct dh_handle_s {
    int device;
    int sm_count;
    int use_tma_store;
    int tile_bytes;
    int ctas_per_sm;
    long long launches;
    void* scratch;        // device scratch (loss partials, NMS masks), grown on demand
    size_t scratch_bytes;
};

namespace dh {

int set_error(int code, const char* fmt, ...);
// Ensure the handle's scratch holds at least `bytes`; returns device pointer or null (error set).
void* scratch(dh_handle_s* h, size_t bytes);

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

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(false) {
        if (cudaGetDevice(&prev) == cudaSuccess && (prev == dev || cudaSetDevice(dev) == cudaSuccess)) ok = true;
    }
    ~DeviceGuard() {
        if (ok && prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace dh
