// Multi-GPU plumbing of the dense-head path behind the C ABI (include/densehead.h, "multi-GPU" section): the batch is
// sharded by image, one rank per GPU, and the only exchange is the sum of the loss scalars.
// Two transports:
//   * peer mailboxes over NVLink (dh_comm.cuh) -- a few posted 4-byte stores per peer and a local spin; runs as one
//     tiny kernel or inside the last CTA of the fused loss kernel (DH_OPT_LOSS_ALLREDUCE);
//   * NCCL (ncclAllReduce on the caller's stream, capturable in a CUDA graph).  libnccl.so.2 is opened with dlopen
//     at the first dh_comm_* call, so the library itself has no link-time dependency on it.
#include <dlfcn.h>

#include <cstring>

#include "dh_comm.cuh"
#include "dh_host.h"

namespace dh {

// ---- the few NCCL entry points used, resolved at run time ------------------------------------------------------
struct NcclUniqueId {
    char internal[128];
};
typedef int (*nccl_get_unique_id_t)(NcclUniqueId*);
typedef int (*nccl_comm_init_rank_t)(void**, int, NcclUniqueId, int);
typedef int (*nccl_comm_init_all_t)(void**, int, const int*);
typedef int (*nccl_all_reduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_comm_destroy_t)(void*);
typedef const char* (*nccl_get_error_string_t)(int);
struct NcclApi {
    void* lib;
    nccl_get_unique_id_t get_unique_id;
    nccl_comm_init_rank_t comm_init_rank;
    nccl_comm_init_all_t comm_init_all;
    nccl_all_reduce_t all_reduce;
    nccl_comm_destroy_t comm_destroy;
    nccl_get_error_string_t error_string;
};
constexpr int kNcclFloat = 7, kNcclSum = 0;

static NcclApi* nccl_api() {
    static NcclApi api = {};
    static bool tried = false;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!tried) {
        tried = true;
        const char* names[] = {getenv("DENSEHEAD_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (api.lib) {
            api.get_unique_id = reinterpret_cast<nccl_get_unique_id_t>(dlsym(api.lib, "ncclGetUniqueId"));
            api.comm_init_rank = reinterpret_cast<nccl_comm_init_rank_t>(dlsym(api.lib, "ncclCommInitRank"));
            api.comm_init_all = reinterpret_cast<nccl_comm_init_all_t>(dlsym(api.lib, "ncclCommInitAll"));
            api.all_reduce = reinterpret_cast<nccl_all_reduce_t>(dlsym(api.lib, "ncclAllReduce"));
            api.comm_destroy = reinterpret_cast<nccl_comm_destroy_t>(dlsym(api.lib, "ncclCommDestroy"));
            api.error_string = reinterpret_cast<nccl_get_error_string_t>(dlsym(api.lib, "ncclGetErrorString"));
            if (!api.get_unique_id || !api.comm_init_rank || !api.comm_init_all || !api.all_reduce || !api.comm_destroy) {
                dlclose(api.lib);
                api.lib = nullptr;
            }
        }
    }
    return api.lib ? &api : nullptr;
}

#define DH_NCCL(api, expr)                                                                                          \
    do {                                                                                                            \
        int r_ = (expr);                                                                                            \
        if (r_ != 0)                                                                                                \
            return set_error(DH_ERR_NCCL, "%s failed: %s", #expr, (api)->error_string ? (api)->error_string(r_) : "?"); \
    } while (0)

struct Comm {
    int world, rank;
    void* nccl;             // ncclComm_t or null
    unsigned int* box;      // this rank's mailbox (cudaMalloc: cudaIpc needs a whole allocation)
    unsigned int* seq;      // step counter, lives behind the mailbox in the same allocation
    bool exported;
    bool peers;             // dev.box[] is complete
    bool ipc_open[DH_COMM_MAX_RANKS];
    CommDev dev;
};

static Comm* comm_of(dh_handle_s* h) {
    if (!h->comm) {
        Comm* c = new Comm();
        memset(c, 0, sizeof(*c));
        c->world = 1;
        h->comm = c;
    }
    return static_cast<Comm*>(h->comm);
}

static int ensure_box(dh_handle_s* h, Comm* c) {
    if (c->box) return DH_OK;
    void* p = nullptr;
    DH_CUDA(cudaMalloc(&p, DH_COMM_BOX_BYTES + 128));
    DH_CUDA(cudaMemset(p, 0, DH_COMM_BOX_BYTES + 128));
    DH_CUDA(cudaDeviceSynchronize());
    c->box = static_cast<unsigned int*>(p);
    c->seq = c->box + DH_COMM_BOX_BYTES / 4;
    c->dev.seq = c->seq;
    c->dev.status = h->dev_status;
    c->dev.timeout_ns = 10ull * 1000ull * 1000ull * 1000ull;
    return DH_OK;
}

__global__ void peer_allreduce_kernel(const __grid_constant__ CommDev c, float* vals, int count) {
    peer_allreduce_warp(c, vals, count, peer_allreduce_seq(c));
}

// used by loss.cu: the device view of the communicator when the in-kernel exchange is on, else null
const CommDev* comm_dev_if_fused(dh_handle_s* h) {
    if (!h->comm || !h->loss_allreduce) return nullptr;
    Comm* c = static_cast<Comm*>(h->comm);
    return (c->peers && c->world > 1) ? &c->dev : nullptr;
}

}  // namespace dh

using namespace dh;

extern "C" {

int dh_comm_get_unique_id(void* out128) {
    DH_CHECK_ARG(out128, "dh_comm_get_unique_id: NULL argument");
    NcclApi* api = nccl_api();
    if (!api) return set_error(DH_ERR_NCCL, "dh_comm_get_unique_id: libnccl.so.2 could not be loaded (%s)", dlerror());
    NcclUniqueId id;
    DH_NCCL(api, api->get_unique_id(&id));
    memcpy(out128, &id, sizeof(id));
    return DH_OK;
}

int dh_comm_init_rank(dh_handle_t h, int world, int rank, const void* unique_id128) {
    DH_CHECK_ARG(h && unique_id128, "dh_comm_init_rank: NULL argument");
    DH_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "dh_comm_init_rank: rank %d of %d", rank, world);
    NcclApi* api = nccl_api();
    if (!api) return set_error(DH_ERR_NCCL, "dh_comm_init_rank: libnccl.so.2 could not be loaded");
    DeviceGuard guard(h);
    Comm* c = comm_of(h);
    DH_CHECK_ARG(!c->nccl, "dh_comm_init_rank: the handle already has an NCCL communicator");
    DH_CHECK_ARG(!c->peers || (c->world == world && c->rank == rank), "dh_comm_init_rank: rank/world differ from the peer setup");
    NcclUniqueId id;
    memcpy(&id, unique_id128, sizeof(id));
    DH_NCCL(api, api->comm_init_rank(&c->nccl, world, id, rank));
    c->world = world, c->rank = rank;
    return DH_OK;
}

int dh_comm_init_all(dh_handle_t* handles, int ndev) {
    DH_CHECK_ARG(handles && ndev >= 1 && ndev <= DH_COMM_MAX_RANKS, "dh_comm_init_all: 1..%d handles", DH_COMM_MAX_RANKS);
    for (int r = 0; r < ndev; ++r) DH_CHECK_ARG(handles[r], "dh_comm_init_all: handle %d is NULL", r);
    // peer mailboxes: one process, so the raw device pointers are valid everywhere once peer access is on
    for (int r = 0; r < ndev; ++r) {
        DeviceGuard guard(handles[r]);
        Comm* c = comm_of(handles[r]);
        DH_CHECK_ARG(!c->peers && !c->nccl, "dh_comm_init_all: handle %d already has a communicator", r);
        int rc = ensure_box(handles[r], c);
        if (rc) return rc;
        for (int q = 0; q < ndev; ++q) {
            if (q == r || handles[q]->device == handles[r]->device) continue;
            int can = 0;
            DH_CUDA(cudaDeviceCanAccessPeer(&can, handles[r]->device, handles[q]->device));
            if (!can)
                return set_error(DH_ERR_CUDA, "dh_comm_init_all: device %d cannot access device %d", handles[r]->device, handles[q]->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(handles[q]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return set_error(DH_ERR_CUDA, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
        }
    }
    for (int r = 0; r < ndev; ++r) {
        Comm* c = comm_of(handles[r]);
        c->world = ndev, c->rank = r;
        c->dev.world = ndev, c->dev.rank = r;
        for (int q = 0; q < ndev; ++q) c->dev.box[q] = comm_of(handles[q])->box;
        c->peers = true;
    }
    // NCCL communicators as well when the library is there (DH_OPT_ALLREDUCE 1 selects them)
    if (NcclApi* api = nccl_api()) {
        void* comms[DH_COMM_MAX_RANKS];
        int devs[DH_COMM_MAX_RANKS];
        for (int r = 0; r < ndev; ++r) devs[r] = handles[r]->device;
        DH_NCCL(api, api->comm_init_all(comms, ndev, devs));
        for (int r = 0; r < ndev; ++r) comm_of(handles[r])->nccl = comms[r];
    }
    return DH_OK;
}

int dh_comm_peer_export(dh_handle_t h, void* out64) {
    DH_CHECK_ARG(h && out64, "dh_comm_peer_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == DH_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    DeviceGuard guard(h);
    Comm* c = comm_of(h);
    int rc = ensure_box(h, c);
    if (rc) return rc;
    cudaIpcMemHandle_t ipc;
    DH_CUDA(cudaIpcGetMemHandle(&ipc, c->box));
    memcpy(out64, &ipc, sizeof(ipc));
    c->exported = true;
    return DH_OK;
}

int dh_comm_peer_import(dh_handle_t h, int world, int rank, const void* handles) {
    DH_CHECK_ARG(h && handles, "dh_comm_peer_import: NULL argument");
    DH_CHECK_ARG(world >= 1 && world <= DH_COMM_MAX_RANKS && rank >= 0 && rank < world, "dh_comm_peer_import: rank %d of %d (at most %d ranks)",
                 rank, world, DH_COMM_MAX_RANKS);
    DeviceGuard guard(h);
    Comm* c = comm_of(h);
    DH_CHECK_ARG(c->exported, "dh_comm_peer_import: call dh_comm_peer_export first");
    DH_CHECK_ARG(!c->peers, "dh_comm_peer_import: peers are already imported");
    DH_CHECK_ARG(!c->nccl || (c->world == world && c->rank == rank), "dh_comm_peer_import: rank/world differ from the NCCL setup");
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            c->dev.box[r] = c->box;
            continue;
        }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, static_cast<const char*>(handles) + r * DH_IPC_HANDLE_BYTES, sizeof(ipc));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q)
                if (c->ipc_open[q]) cudaIpcCloseMemHandle(c->dev.box[q]), c->ipc_open[q] = false;
            return set_error(DH_ERR_CUDA, "dh_comm_peer_import: cudaIpcOpenMemHandle of rank %d failed: %s", r, cudaGetErrorString(e));
        }
        c->dev.box[r] = static_cast<unsigned int*>(p);
        c->ipc_open[r] = true;
    }
    c->world = world, c->rank = rank;
    c->dev.world = world, c->dev.rank = rank;
    c->peers = true;
    return DH_OK;
}

int dh_comm_info(dh_handle_t h, int* world, int* rank, int* transports) {
    DH_CHECK_ARG(h, "dh_comm_info: handle is NULL");
    Comm* c = static_cast<Comm*>(h->comm);
    if (world) *world = c ? c->world : 1;
    if (rank) *rank = c ? c->rank : 0;
    if (transports) *transports = c ? ((c->peers ? 2 : 0) | (c->nccl ? 1 : 0)) : 0;
    return DH_OK;
}

int dh_allreduce_loss(dh_handle_t h, float* scalars, int count, void* stream) {
    DH_CHECK_ARG(h && scalars, "dh_allreduce_loss: NULL argument");
    DH_CHECK_ARG(count >= 1 && count <= DH_COMM_MAX_VALUES, "dh_allreduce_loss: count %d not in [1,%d]", count, DH_COMM_MAX_VALUES);
    Comm* c = static_cast<Comm*>(h->comm);
    if (!c || c->world <= 1) return DH_OK;  // a single rank: the sum is the input
    DeviceGuard guard(h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool want_nccl = h->allreduce_mode == 1 || !c->peers;
    if (want_nccl) {
        NcclApi* api = nccl_api();
        if (!c->nccl || !api) return set_error(DH_ERR_NCCL, "dh_allreduce_loss: no NCCL communicator on this handle (dh_comm_init_rank / dh_comm_init_all)");
        DH_NCCL(api, api->all_reduce(scalars, scalars, static_cast<size_t>(count), kNcclFloat, kNcclSum, c->nccl, st));
        return DH_OK;
    }
    peer_allreduce_kernel<<<1, 32, 0, st>>>(c->dev, scalars, count);
    DH_CUDA(cudaGetLastError());
    h->launches += 1;
    return DH_OK;
}

int dh_comm_destroy(dh_handle_t h) {
    if (!h || !h->comm) return DH_OK;
    DeviceGuard guard(h);
    Comm* c = static_cast<Comm*>(h->comm);
    cudaDeviceSynchronize();
    if (c->nccl) {
        if (NcclApi* api = nccl_api()) api->comm_destroy(c->nccl);
    }
    for (int r = 0; r < DH_COMM_MAX_RANKS; ++r)
        if (c->ipc_open[r]) cudaIpcCloseMemHandle(c->dev.box[r]);
    if (c->box) cudaFree(c->box);
    delete c;
    h->comm = nullptr;
    return DH_OK;
}

}  // extern "C"
