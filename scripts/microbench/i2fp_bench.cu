#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(const uint32_t* in, float* out, int n, long long* cyc) {
    uint32_t w = in[threadIdx.x];
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t x = w + it * 8 + u;
            float f0, f1, f2, f3;
            if (MODE == 0) {
                int i0, i1, i2, i3;
                asm("prmt.b32 %0, %1, 0, 0x8880;" : "=r"(i0) : "r"(x));
                asm("prmt.b32 %0, %1, 0, 0x9991;" : "=r"(i1) : "r"(x));
                asm("prmt.b32 %0, %1, 0, 0xaaa2;" : "=r"(i2) : "r"(x));
                asm("prmt.b32 %0, %1, 0, 0xbbb3;" : "=r"(i3) : "r"(x));
                f0 = __int2float_rn(i0); f1 = __int2float_rn(i1); f2 = __int2float_rn(i2); f3 = __int2float_rn(i3);
            } else if (MODE == 1) {
                f0 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7540)) - 8388736.f;
                f1 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7541)) - 8388736.f;
                f2 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7542)) - 8388736.f;
                f3 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7543)) - 8388736.f;
            } else {
                f0 = __uint_as_float(x); f1 = __uint_as_float(x + 1); f2 = __uint_as_float(x + 2); f3 = __uint_as_float(x + 3);
            }
            a0 += f0; a1 += f1; a2 += f2; a3 += f3;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    uint32_t* in; float* out; long long* cyc;
    cudaMalloc(&in, 4096); cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
    cudaMemset(in, 1, 4096);
    const int n = 4096;
    for (int threads : {128, 256, 512}) {
        long long c[3];
        for (int m = 0; m < 3; ++m) {
            for (int rep = 0; rep < 2; ++rep) {
                if (m == 0) k<0><<<148, threads>>>(in, out, n, cyc);
                if (m == 1) k<1><<<148, threads>>>(in, out, n, cyc);
                if (m == 2) k<2><<<148, threads>>>(in, out, n, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&c[m], cyc, 8, cudaMemcpyDeviceToHost);
        }
        const double el = (double)n * 8 * 4 * (threads / 128);
        printf("warps/SMSP %d: cycles per 32 converted elements: prmt.sext+i2fp %.2f  prmt+fadd %.2f  base %.2f\n",
               threads / 128, c[0] / el, c[1] / el, c[2] / el);
    }
    return 0;
}
