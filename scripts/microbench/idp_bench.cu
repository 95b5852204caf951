// Micro-benchmark for the A role of sweep_fast_kernel (int8 LD): what one tile x row-group costs in isolation.
//   body: 4 LDS.128 (one per row of the group) + 48 IDP.4A (4 rows x 3 digits x 4 words) [+ 12 REDUX]
// Reports cycles per body for 1 / 2 / 4 warps per SM sub-partition (one CTA on one SM), and the raw issue cost of
// IDP.4A, REDUX, PRMT+FADD2+FFMA2 streams.   nvcc -arch=sm_100a -O3 -o idp_bench idp_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tmov.b64 rc, {%6,%7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(int iters, long long* out, int* sink, int rowstride) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t hl[3][4];
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int w = 0; w < 4; ++w) hl[l][w] = 0x01020304u * (l + 1) + w + tid;
    int dig[3] = {0, 0, 0};
    float f[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = 0.f;
    const uint32_t a0 = sbase + lane * 16 + (tid >> 5) * 512;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t ab = a0 + ((it * 2048) & 0x7fff);
        if (MODE == 0 || MODE == 1 || MODE == 4) {
            uint4 cv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) cv[r] = lds128(ab + r * rowstride);
            int acc[4][3];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    int a = 0;
                    a = dp4a_us(cv[r].x, hl[l][0], a);
                    a = dp4a_us(cv[r].y, hl[l][1], a);
                    a = dp4a_us(cv[r].z, hl[l][2], a);
                    a = dp4a_us(cv[r].w, hl[l][3], a);
                    acc[r][l] = a;
                }
            if (MODE == 1) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int l = 0; l < 3; ++l) {
                        const int tot = __reduce_add_sync(0xffffffffu, acc[r][l]);
                        if (lane == r) dig[l] = tot;
                    }
            } else if (MODE == 4) {
                // two tiles' worth before the reduction (what the kernel does with NVT = 2)
                uint4 cw[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) cw[r] = lds128(ab + 1024 + r * rowstride);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int l = 0; l < 3; ++l) {
                        int a = acc[r][l];
                        a = dp4a_us(cw[r].x, hl[l][0], a);
                        a = dp4a_us(cw[r].y, hl[l][1], a);
                        a = dp4a_us(cw[r].z, hl[l][2], a);
                        a = dp4a_us(cw[r].w, hl[l][3], a);
                        acc[r][l] = a;
                    }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int l = 0; l < 3; ++l) {
                        const int tot = __reduce_add_sync(0xffffffffu, acc[r][l]);
                        if (lane == r) dig[l] = tot;
                    }
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int l = 0; l < 3; ++l) dig[l] += acc[r][l];
            }
        } else if (MODE == 2) {          // 12 REDUX only
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    const int tot = __reduce_add_sync(0xffffffffu, dig[l] + r + it);
                    if (lane == r) dig[l] = tot;
                }
        } else if (MODE == 3) {          // C body: 4 LDS.128 + decode + axpy (64 PRMT, 32 FADD2, 32 FFMA2)
            uint4 cv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) cv[r] = lds128(ab + r * rowstride);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float al = __int_as_float(0x3f800000 + r + it);
                const uint32_t w[4] = {cv[r].x, cv[r].y, cv[r].z, cv[r].w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int b = 0; b < 4; b += 2) {
                        float2 m;
                        m.x = __uint_as_float(__byte_perm(w[q], 0x4b000000u, 0x7440 + b));
                        m.y = __uint_as_float(__byte_perm(w[q], 0x4b000000u, 0x7441 + b));
                        const float2 x = add2(m, make_float2(-8388736.f, -8388736.f));
                        float2 acc = make_float2(f[q * 4 + b], f[q * 4 + b + 1]);
                        acc = fma2(x, make_float2(al, al), acc);
                        f[q * 4 + b] = acc.x; f[q * 4 + b + 1] = acc.y;
                    }
            }
        }
    }
    const long long t1 = clock64();
    int s = dig[0] + dig[1] + dig[2];
#pragma unroll
    for (int e = 0; e < 16; ++e) s += __float_as_int(f[e]);
    if (s == 0x12345678) sink[tid] = s;
    if (tid == 0) out[0] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int threads, int rowstride) {
    long long* out; int* sink;
    cudaMalloc(&out, 8); cudaMalloc(&sink, 4096);
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + 8192);
    const int iters = 20000;
    bench<MODE><<<1, threads, 64 * 1024 + 8192>>>(iters, out, sink, rowstride);
    bench<MODE><<<1, threads, 64 * 1024 + 8192>>>(iters, out, sink, rowstride);
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-46s warps/SMSP %d  rowstride %5d: %8.1f cycles per body\n", name, threads / 128, rowstride, (double)h / iters);
    cudaFree(out); cudaFree(sink);
}

int main() {
    for (int threads : {128, 256, 512}) {
        run<0>("A body: 4 LDS.128 + 48 IDP.4A", threads, 4080);
        run<1>("A body + 12 REDUX", threads, 4080);
        run<4>("A body x 2 tiles + 12 REDUX (kernel's row group)", threads, 4080);
        run<2>("12 REDUX only", threads, 4080);
        run<3>("C body: 4 LDS.128 + 64 PRMT + 64 FADD/FFMA", threads, 4080);
    }
    return 0;
}
