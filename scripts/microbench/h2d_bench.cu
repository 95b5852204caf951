// How to get 8 separate page-locked host arrays (4.4 MB each, in 4 row chunks) onto the device fastest:
// DMA copies on one stream / several streams / cudaMemcpyBatchAsync, SM-driven reads of mapped memory, and both at once.
//   nvcc -arch=sm_100a -O3 -o h2d_bench h2d_bench.cu
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

struct List { const uint4* s[8]; uint4* d[8]; size_t n16; };
__global__ void __launch_bounds__(256) gather(const List L, int unroll) {
    const int a = blockIdx.y;
    const uint4* s = L.s[a]; uint4* d = L.d[a];
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (unroll == 4) {
        for (; i + 3 * nth < L.n16; i += 4 * nth) {
            const uint4 v0 = __ldcv(s + i), v1 = __ldcv(s + i + nth), v2 = __ldcv(s + i + 2 * nth), v3 = __ldcv(s + i + 3 * nth);
            d[i] = v0; d[i + nth] = v1; d[i + 2 * nth] = v2; d[i + 3 * nth] = v3;
        }
    }
    for (; i < L.n16; i += nth) d[i] = __ldcv(s + i);
}

int main() {
    const int NA = 8, NCH = 4;
    const size_t bytes = 1101824 * 4, chunk = bytes / NCH;
    void* h[NA]; void* d[NA];
    for (int a = 0; a < NA; ++a) { CK(cudaHostAlloc(&h[a], bytes, cudaHostAllocDefault)); CK(cudaMalloc(&d[a], bytes)); }
    cudaStream_t st[9];
    for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char* name, auto&& fn) {
        for (int w = 0; w < 3; ++w) fn();
        cudaDeviceSynchronize();
        const int reps = 20;
        auto t0 = std::chrono::high_resolution_clock::now();
        for (int r = 0; r < reps; ++r) { fn(); cudaDeviceSynchronize(); }
        auto t1 = std::chrono::high_resolution_clock::now();
        const double us = std::chrono::duration<double, std::micro>(t1 - t0).count() / reps;
        printf("%-58s %8.1f us  %6.1f GB/s\n", name, us, NA * bytes / us / 1e3);
    };
    timeit("DMA 8 arrays x 4 chunks, one stream", [&] {
        for (int c = 0; c < NCH; ++c) for (int a = 0; a < NA; ++a)
            cudaMemcpyAsync((char*)d[a] + c * chunk, (char*)h[a] + c * chunk, chunk, cudaMemcpyHostToDevice, st[0]);
    });
    timeit("DMA 8 arrays x 4 chunks, stream per chunk", [&] {
        for (int c = 0; c < NCH; ++c) for (int a = 0; a < NA; ++a)
            cudaMemcpyAsync((char*)d[a] + c * chunk, (char*)h[a] + c * chunk, chunk, cudaMemcpyHostToDevice, st[c]);
    });
    timeit("DMA 8 arrays x 4 chunks, stream per array", [&] {
        for (int c = 0; c < NCH; ++c) for (int a = 0; a < NA; ++a)
            cudaMemcpyAsync((char*)d[a] + c * chunk, (char*)h[a] + c * chunk, chunk, cudaMemcpyHostToDevice, st[a]);
    });
    timeit("DMA 8 whole arrays, one stream", [&] {
        for (int a = 0; a < NA; ++a) cudaMemcpyAsync(d[a], h[a], bytes, cudaMemcpyHostToDevice, st[0]);
    });
    timeit("DMA 8 whole arrays, stream per array", [&] {
        for (int a = 0; a < NA; ++a) cudaMemcpyAsync(d[a], h[a], bytes, cudaMemcpyHostToDevice, st[a]);
    });
#if CUDART_VERSION >= 12080
    timeit("cudaMemcpyBatchAsync, 4 batches of 8 chunks, one stream", [&] {
        for (int c = 0; c < NCH; ++c) {
            void* ds[NA]; void* ss[NA]; size_t sz[NA];
            for (int a = 0; a < NA; ++a) { ds[a] = (char*)d[a] + c * chunk; ss[a] = (char*)h[a] + c * chunk; sz[a] = chunk; }
            cudaMemcpyAttributes at = {};
            at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
            size_t idx = 0, fail = 0;
            cudaError_t e = cudaMemcpyBatchAsync(ds, ss, sz, NA, &at, &idx, 1, &fail, st[0]);
            if (e != cudaSuccess) { printf("batch: %s\n", cudaGetErrorString(e)); cudaGetLastError(); }
        }
    });
#endif
    for (int gx : {8, 32, 128}) for (int un : {1, 4}) {
        char nm[96];
        snprintf(nm, sizeof nm, "SM gather, 4 chunk kernels, grid (%d, 8), unroll %d", gx, un);
        timeit(nm, [&] {
            for (int c = 0; c < NCH; ++c) {
                List L; L.n16 = chunk / 16;
                for (int a = 0; a < NA; ++a) { L.s[a] = (const uint4*)((char*)h[a] + c * chunk); L.d[a] = (uint4*)((char*)d[a] + c * chunk); }
                gather<<<dim3(gx, NA), 256, 0, st[0]>>>(L, un);
            }
        });
    }
    timeit("half DMA (4 arrays, stream 0) + half SM gather (stream 1)", [&] {
        for (int c = 0; c < NCH; ++c) {
            for (int a = 0; a < 4; ++a)
                cudaMemcpyAsync((char*)d[a] + c * chunk, (char*)h[a] + c * chunk, chunk, cudaMemcpyHostToDevice, st[0]);
            List L; L.n16 = chunk / 16;
            for (int a = 0; a < 4; ++a) { L.s[a] = (const uint4*)((char*)h[4 + a] + c * chunk); L.d[a] = (uint4*)((char*)d[4 + a] + c * chunk); }
            gather<<<dim3(32, 4), 256, 0, st[1]>>>(L, 4);
        }
    });
    // one staging slab: what a caller that keeps its state in ONE allocation would get
    void* hs; void* dsl;
    CK(cudaHostAlloc(&hs, NA * bytes, cudaHostAllocDefault)); CK(cudaMalloc(&dsl, NA * bytes));
    timeit("DMA one slab of 8 arrays, 4 chunks (interleaved layout)", [&] {
        for (int c = 0; c < NCH; ++c) cudaMemcpyAsync((char*)dsl + c * NA * chunk, (char*)hs + c * NA * chunk, NA * chunk, cudaMemcpyHostToDevice, st[0]);
    });
    int ce = 0; cudaDeviceGetAttribute(&ce, cudaDevAttrAsyncEngineCount, 0);
    printf("asyncEngineCount %d\n", ce);
    return 0;
}
