// Development probe: how fast can one CTA stream zero tiles to HBM with cp.async.bulk stores, as a
// function of tile size, stages in flight and CTAs/SM -- and how does st.global.v4 compare?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_store_probe tools/tma_store_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../cv-lite-object-detection_b200/csrc/dh_common.cuh"
using namespace dh;

template <int STAGES>
__device__ __forceinline__ void wait_read_n() { bulk_wait_read<STAGES - 1>(); }

// mode 0: TMA bulk store; mode 1: st.global.v4 from smem; mode 2: st.global.v4 zeros from registers
template <int STAGES>
__global__ void __launch_bounds__(256) probe(float* out, long long total_bytes, int tile_bytes, int mode, int interleave,
                                             long long* cyc) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    for (int e = tid; e < STAGES * tile_bytes / 16; e += 256) reinterpret_cast<float4*>(smem)[e] = make_float4(0, 0, 0, 0);
    __syncthreads();
    const long long n_tiles = total_bytes / tile_bytes;
    long long t0, t1, step;
    if (interleave) { t0 = blockIdx.x; t1 = n_tiles; step = gridDim.x; }
    else { t0 = n_tiles * blockIdx.x / gridDim.x; t1 = n_tiles * (blockIdx.x + 1) / gridDim.x; step = 1; }
    long long c0 = clock64();
    int it = 0;
    for (long long t = t0; t < t1; t += step, ++it) {
        const int s = it % STAGES;
        unsigned char* buf = smem + s * tile_bytes;
        char* g = reinterpret_cast<char*>(out) + t * tile_bytes;
        if (mode == 0) {
            if (it >= STAGES) { if (tid == 0) wait_read_n<STAGES>(); }
            __syncthreads();
            // (a real kernel writes a few rows here)
            if (tid == 3) buf[16] = 0;
            fence_async_smem();
            __syncthreads();
            if (tid == 0) { bulk_s2g(g, buf, tile_bytes); bulk_commit(); }
        } else if (mode == 1) {
            __syncthreads();
            const float4* s4 = reinterpret_cast<const float4*>(buf);
            float4* d4 = reinterpret_cast<float4*>(g);
            for (int e = tid; e < tile_bytes / 16; e += 256) d4[e] = s4[e];
        } else {
            float4* d4 = reinterpret_cast<float4*>(g);
            const float4 z = make_float4(0, 0, 0, 0);
            for (int e = tid; e < tile_bytes / 16; e += 256) d4[e] = z;
        }
    }
    if (mode == 0 && tid == 0) bulk_wait_read<0>();
    if (tid == 0 && blockIdx.x == 0 && cyc) cyc[0] = (clock64() - c0) / (it > 0 ? it : 1);
}

template <int STAGES>
void run(float* out, long long bytes, int tile, int mode, int ctas, int inter, long long* dcyc) {
    int sm = 148;
    size_t smem = (size_t)STAGES * tile + 128;
    cudaFuncSetAttribute(probe<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) probe<STAGES><<<sm * ctas, 256, smem>>>(out, bytes, tile, mode, inter, dcyc);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) probe<STAGES><<<sm * ctas, 256, smem>>>(out, bytes, tile, mode, inter, dcyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    long long cyc = 0; cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
    cudaError_t err = cudaGetLastError();
    printf("mode=%d stages=%d tile=%6d ctas=%d inter=%d : %8.1f us  %7.1f GB/s  cyc/tile(blk0)=%lld %s\n", mode, STAGES, tile, ctas,
           inter, ms * 1e3, bytes / (ms * 1e-3) / 1e9, cyc, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    const long long bytes = 1536ll << 20;
    float* out; cudaMalloc(&out, bytes);
    long long* dcyc; cudaMalloc(&dcyc, 8);
    for (int inter = 0; inter < 2; ++inter) {
        for (int tile : {8192, 16384, 32768}) {
            for (int ctas : {1, 2, 4}) {
                run<2>(out, bytes, tile, 0, ctas, inter, dcyc);
                if ((size_t)4 * tile * ctas <= 200 * 1024) run<4>(out, bytes, tile, 0, ctas, inter, dcyc);
                if ((size_t)8 * tile * ctas <= 200 * 1024) run<8>(out, bytes, tile, 0, ctas, inter, dcyc);
            }
        }
        for (int tile : {8192, 32768}) for (int ctas : {1, 2, 4}) { run<2>(out, bytes, tile, 1, ctas, inter, dcyc); run<2>(out, bytes, tile, 2, ctas, inter, dcyc); }
    }
    run<2>(out, bytes, 32768, 2, 8, 1, dcyc);
    cudaMemset(out, 0, bytes); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); for (int r = 0; r < 5; ++r) cudaMemsetAsync(out, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("cudaMemset: %.1f GB/s\n", bytes / (ms / 5 * 1e-3) / 1e9);
    return 0;
}
