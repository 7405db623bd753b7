// Handle, options and error reporting for libdensehead.so (see include/densehead.h).
#include <cstring>

#include "dh_host.h"

namespace dh {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void* scratch(dh_handle_s* h, size_t bytes) {
    if (bytes <= h->scratch_bytes) return h->scratch;
    if (h->scratch) {
        cudaDeviceSynchronize();  // a kernel in flight may still use the old block
        cudaFree(h->scratch);
        h->scratch = nullptr;
        h->scratch_bytes = 0;
    }
    size_t want = bytes + (bytes >> 1) + (1u << 20);
    cudaError_t e = cudaMalloc(&h->scratch, want);
    if (e != cudaSuccess) {
        set_error(DH_ERR_CUDA, "cudaMalloc(%zu) for scratch failed: %s", want, cudaGetErrorString(e));
        return nullptr;
    }
    h->scratch_bytes = want;
    return h->scratch;
}

void* scratch_b(dh_handle_s* h, size_t bytes) {
    if (bytes <= h->scratch_b_bytes) return h->scratch_b;
    if (h->scratch_b) {
        cudaDeviceSynchronize();
        cudaFree(h->scratch_b);
        h->scratch_b = nullptr;
        h->scratch_b_bytes = 0;
    }
    size_t want = bytes + (bytes >> 2) + (1u << 20);
    cudaError_t e = cudaMalloc(&h->scratch_b, want);
    if (e != cudaSuccess) {
        set_error(DH_ERR_CUDA, "cudaMalloc(%zu) for pipeline scratch failed: %s", want, cudaGetErrorString(e));
        return nullptr;
    }
    h->scratch_b_bytes = want;
    return h->scratch_b;
}

constexpr int kSchedRing = 64;

unsigned int* next_sched_counter(dh_handle_s* h, cudaStream_t st) {
    (void)st;
    if (!h->sched) {
        // one 128-byte line per launch in flight: word 0 = next chunk, word 1 = CTAs done.  Zeroed once; the last CTA of
        // every launch puts its line back to zero (sched_release), so no per-launch memset node is needed.
        cudaError_t e = cudaMalloc(&h->sched, kSchedRing * 32 * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMemset(h->sched, 0, kSchedRing * 32 * sizeof(unsigned int));
        if (e != cudaSuccess) {
            set_error(DH_ERR_CUDA, "allocating the scheduler counters failed: %s", cudaGetErrorString(e));
            return nullptr;
        }
    }
    return h->sched + (h->sched_next++ % kSchedRing) * 32;
}

}  // namespace dh

extern "C" {

const char* dh_version(void) { return "densehead-b200 0.1.0 (sm_100a)"; }

const char* dh_last_error(void) { return dh::g_err; }

int dh_create(dh_handle_t* out, int device) {
    DH_CHECK_ARG(out != nullptr, "dh_create: out is NULL");
    int n = 0;
    DH_CUDA(cudaGetDeviceCount(&n));
    DH_CHECK_ARG(device >= 0 && device < n, "dh_create: device %d out of range (%d visible)", device, n);
    cudaDeviceProp prop;
    DH_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return dh::set_error(DH_ERR_CUDA, "dh_create: device %d is sm_%d%d; this library is built for sm_100a only",
                             device, prop.major, prop.minor);
    dh_handle_s* h = new dh_handle_s();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->use_tma_store = 1;
    h->tile_bytes = 49152;
    h->ctas_per_sm = 4;
    h->fused_loss_kernel = 0;
    h->nms_kernel = 0;
    h->fused_chunks_per_cta = 12;
    h->encode_min_chunk = 2;
    h->fcos_select_mode = 0;
    h->nms_sort = 0;
    h->nms_filter = 1;
    h->nms_chain = 0;
    h->loss_allreduce = 0;
    h->allreduce_mode = 0;
    h->fused_tail = 1;
    h->fused_max_chunk = 16;
    h->encode_kernel = 0;
    h->launches = 0;
    h->scratch = nullptr;
    h->scratch_bytes = 0;
    h->scratch_b = nullptr;
    h->scratch_b_bytes = 0;
    h->phase_cycles = nullptr;
    h->sched = nullptr;
    h->sched_next = 0;
    h->comm = nullptr;
    h->dev_status = nullptr;
    h->trace = nullptr;
    h->trace_bytes = 0;
    {
        dh::DeviceGuard g(device);
        cudaError_t e = cudaMalloc(&h->dev_status, 128);
        if (e == cudaSuccess) e = cudaMemset(h->dev_status, 0, 128);
        if (e != cudaSuccess) {
            delete h;
            return dh::set_error(DH_ERR_CUDA, "dh_create: allocating the status word failed: %s", cudaGetErrorString(e));
        }
    }
    *out = h;
    return DH_OK;
}

int dh_destroy(dh_handle_t h) {
    if (!h) return DH_OK;
    dh_comm_destroy(h);
    {
        dh::DeviceGuard g(h->device);
        if (h->scratch) cudaFree(h->scratch);
        if (h->scratch_b) cudaFree(h->scratch_b);
        if (h->sched) cudaFree(h->sched);
        if (h->phase_cycles) cudaFree(h->phase_cycles);
        if (h->dev_status) cudaFree(h->dev_status);
    }
    delete h;
    return DH_OK;
}

int dh_set_option(dh_handle_t h, int option, int value) {
    DH_CHECK_ARG(h != nullptr, "dh_set_option: handle is NULL");
    switch (option) {
        case DH_OPT_TMA_STORE:
            h->use_tma_store = value ? 1 : 0;
            return DH_OK;
        case DH_OPT_TILE_BYTES:
            DH_CHECK_ARG(value >= 2048 && value <= 98304, "DH_OPT_TILE_BYTES must be in [2048, 98304]");
            h->tile_bytes = value & ~127;
            return DH_OK;
        case DH_OPT_CTAS_PER_SM:
            DH_CHECK_ARG(value >= 1 && value <= 8, "DH_OPT_CTAS_PER_SM must be in [1, 8]");
            h->ctas_per_sm = value;
            return DH_OK;
        case DH_OPT_FUSED_LOSS_KERNEL:
            DH_CHECK_ARG(value == 0 || value == 1, "DH_OPT_FUSED_LOSS_KERNEL must be 0 or 1");
            h->fused_loss_kernel = value;
            return DH_OK;
        case DH_OPT_NMS_KERNEL:
            DH_CHECK_ARG(value >= 0 && value <= 2, "DH_OPT_NMS_KERNEL must be 0, 1 or 2");
            h->nms_kernel = value;
            return DH_OK;
        case DH_OPT_FCOS_SELECT:
            DH_CHECK_ARG(value >= 0 && value <= 4, "DH_OPT_FCOS_SELECT must be in [0, 4]");
            h->fcos_select_mode = value;
            return DH_OK;
        case DH_OPT_NMS_SORT:
            DH_CHECK_ARG(value == 0 || value == 1, "DH_OPT_NMS_SORT must be 0 or 1");
            h->nms_sort = value;
            return DH_OK;
        case DH_OPT_LOSS_ALLREDUCE:
            DH_CHECK_ARG(value == 0 || value == 1, "DH_OPT_LOSS_ALLREDUCE must be 0 or 1");
            h->loss_allreduce = value;
            return DH_OK;
        case DH_OPT_ALLREDUCE:
            DH_CHECK_ARG(value >= 0 && value <= 2, "DH_OPT_ALLREDUCE must be 0, 1 or 2");
            h->allreduce_mode = value;
            return DH_OK;
        case DH_OPT_FUSED_TAIL:
            DH_CHECK_ARG(value >= 0 && (value >> 1) <= 5, "DH_OPT_FUSED_TAIL must be 0 or 1 (+ 2 * log2 of the spans of a chunk's last tile, 1..5)");
            h->fused_tail = value;
            return DH_OK;
        case DH_OPT_FUSED_MAX_CHUNK:
            DH_CHECK_ARG(value >= 4 && value <= 16, "DH_OPT_FUSED_MAX_CHUNK must be in [4, 16]");
            h->fused_max_chunk = value;
            return DH_OK;
        case DH_OPT_NMS_FILTER:
            DH_CHECK_ARG(value >= 0 && value <= 64, "DH_OPT_NMS_FILTER must be in [0, 64]");
            h->nms_filter = value;
            return DH_OK;
        case DH_OPT_NMS_CHAIN:
            DH_CHECK_ARG(value == 0 || value == 1, "DH_OPT_NMS_CHAIN must be 0 or 1");
            h->nms_chain = value;
            return DH_OK;
        case DH_OPT_ENCODE_KERNEL:
            DH_CHECK_ARG(value >= 0 && value <= 2, "DH_OPT_ENCODE_KERNEL must be 0, 1 or 2");
            h->encode_kernel = value;
            return DH_OK;
        case DH_OPT_FUSED_CHUNKS_PER_CTA:
            DH_CHECK_ARG(value >= 1 && value <= 64, "DH_OPT_FUSED_CHUNKS_PER_CTA must be in [1, 64]");
            h->fused_chunks_per_cta = value;
            return DH_OK;
        case DH_OPT_ENCODE_MIN_CHUNK:
            DH_CHECK_ARG(value >= 1 && value <= 64, "DH_OPT_ENCODE_MIN_CHUNK must be in [1, 64]");
            h->encode_min_chunk = value;
            return DH_OK;
        case DH_OPT_PHASE_TIMING: {
            dh::DeviceGuard g(h->device);
            if (value && !h->phase_cycles) {
                DH_CUDA(cudaMalloc(&h->phase_cycles, 8 * sizeof(long long)));
            }
            if (h->phase_cycles) DH_CUDA(cudaMemset(h->phase_cycles, 0, 8 * sizeof(long long)));
            if (!value && h->phase_cycles) {
                cudaFree(h->phase_cycles);
                h->phase_cycles = nullptr;
            }
            return DH_OK;
        }
        default:
            return dh::set_error(DH_ERR_BAD_ARG, "dh_set_option: unknown option %d", option);
    }
}

long long dh_launch_count(dh_handle_t h) { return h ? h->launches : 0; }

int dh_set_trace(dh_handle_t h, long long* buf, long long bytes) {
    DH_CHECK_ARG(h, "dh_set_trace: handle is NULL");
    h->trace = buf;
    h->trace_bytes = buf ? bytes : 0;
    return DH_OK;
}

int dh_get_status(dh_handle_t h, int32_t* out, int reset) {
    DH_CHECK_ARG(h && out, "dh_get_status: NULL argument");
    dh::DeviceGuard g(h);
    DH_CUDA(cudaMemcpy(out, h->dev_status, sizeof(int32_t), cudaMemcpyDeviceToHost));  // synchronises with the null stream
    if (reset && *out) DH_CUDA(cudaMemset(h->dev_status, 0, sizeof(int32_t)));
    return DH_OK;
}

int dh_read_phase_timing(dh_handle_t h, long long* out8) {
    DH_CHECK_ARG(h && out8, "dh_read_phase_timing: NULL argument");
    DH_CHECK_ARG(h->phase_cycles, "dh_read_phase_timing: DH_OPT_PHASE_TIMING is off");
    dh::DeviceGuard g(h->device);
    DH_CUDA(cudaMemcpy(out8, h->phase_cycles, 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    return DH_OK;
}

}  // extern "C"
