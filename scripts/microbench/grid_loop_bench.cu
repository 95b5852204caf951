// Micro-benchmark of the grid sweep's bulk inner loop (viprs_b200/csrc/grid.cuh): one thread = 16 LD columns x 8 grid
// columns of q in registers (64 float2 accumulators); per LD row 64 FFMA2 (decoded LD value as 32-bit broadcast operand,
// the two scaled deltas of a grid-column pair as a 64-bit operand).  Reports FMA-pipe cycles per row per SM sub-partition
// for NW warps per sub-partition and several loop forms, against the 2 cycles per FFMA2 the pipe needs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o grid_loop_bench grid_loop_bench.cu && ./grid_loop_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
          "l"(reinterpret_cast<unsigned long long&>(c)));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return d;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void decode4(uint32_t w, float* o) {
    float2 p0 = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7541)));
    float2 p1 = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7542)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7543)));
    const float2 m = make_float2(-8388736.f, -8388736.f);
    p0 = add2(p0, m); p1 = add2(p1, m);
    o[0] = p0.x; o[1] = p0.y; o[2] = p1.x; o[3] = p1.y;
}

// the same through the conversion unit: I2F.S8 with a byte selector (no PRMT, no FADD2; XU pipe, 8 cycles per warp instruction);
// the codes are stored biased (code ^ 0x80), so one LOP3 per word restores two's complement first
__device__ __forceinline__ void decode4_i2f(uint32_t w, float* o) {
    w ^= 0x80808080u;
    asm("{.reg .b8 b0,b1,b2,b3; mov.b32 {b0,b1,b2,b3}, %4; cvt.rn.f32.s8 %0, b0; cvt.rn.f32.s8 %1, b1; cvt.rn.f32.s8 %2, b2; cvt.rn.f32.s8 %3, b3;}"
        : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "r"(w));
}

// two's-complement codes (a signed copy of the dense block, no bias): PRMT with sign replication (selector msb) extends a
// byte to 32 bits, I2FP.F32.S32 converts it -- two ALU-side instructions per code and nothing on the FMA pipe
__device__ __forceinline__ void decode4_i2fp(uint32_t w, float* o) {
    int i0, i1, i2, i3;
    asm("prmt.b32 %0, %1, 0, 0x8880;" : "=r"(i0) : "r"(w));
    asm("prmt.b32 %0, %1, 0, 0x9991;" : "=r"(i1) : "r"(w));
    asm("prmt.b32 %0, %1, 0, 0xaaa2;" : "=r"(i2) : "r"(w));
    asm("prmt.b32 %0, %1, 0, 0xbbb3;" : "=r"(i3) : "r"(w));
    o[0] = __int2float_rn(i0); o[1] = __int2float_rn(i1); o[2] = __int2float_rn(i2); o[3] = __int2float_rn(i3);
}

constexpr int ROWS = 16, ROWB = 4096;

// MODE 0: FFMA2 only (values fixed in registers)   1: + deltas from shared memory every row
//      2: + codes from shared memory, decode (the kernel's plain loop)   3: software-pipelined decode (two-row body)
template <int MODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k(int nrows, float* out, long long* cyc) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    for (int i = tid; i < (ROWS * ROWB + ROWS * 32) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x80808080u + (uint32_t)i * 0x01030507u % 7u;
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t rowp = sbase + (uint32_t)(tid % 256) * 16u, abase = sbase + ROWS * ROWB;
    float2 q[16][4];
#pragma unroll
    for (int e = 0; e < 16; ++e)
#pragma unroll
        for (int g = 0; g < 4; ++g) q[e][g] = make_float2(0.f, 0.f);
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = (float)(tid + e);
    float2 al[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) al[g] = make_float2(1e-3f * g, 2e-3f * g);
    __syncthreads();
    const long long t0 = clock64();
    if (MODE <= 2 || MODE >= 4) {
#pragma unroll 1
        for (int r = 0; r < nrows; ++r) {
            const uint32_t rr = (uint32_t)(r & (ROWS - 1));
            if (MODE >= 1) {
                const uint4 a0 = lds128(abase + rr * 32), a1 = lds128(abase + rr * 32 + 16);
                al[0] = make_float2(__uint_as_float(a0.x), __uint_as_float(a0.y));
                al[1] = make_float2(__uint_as_float(a0.z), __uint_as_float(a0.w));
                al[2] = make_float2(__uint_as_float(a1.x), __uint_as_float(a1.y));
                al[3] = make_float2(__uint_as_float(a1.z), __uint_as_float(a1.w));
            }
            if (MODE == 2) {
                const uint4 cv = lds128(rowp + rr * ROWB);
                decode4(cv.x, v); decode4(cv.y, v + 4); decode4(cv.z, v + 8); decode4(cv.w, v + 12);
            }
            if (MODE == 4) {
                const uint4 cv = lds128(rowp + rr * ROWB);
                decode4_i2f(cv.x, v); decode4_i2f(cv.y, v + 4); decode4_i2f(cv.z, v + 8); decode4_i2f(cv.w, v + 12);
            }
            if (MODE == 5) {
                const uint4 cv = lds128(rowp + rr * ROWB);
                decode4_i2f(cv.x, v); decode4(cv.y, v + 4); decode4_i2f(cv.z, v + 8); decode4(cv.w, v + 12);
            }
            if (MODE == 7) {
                const uint4 cv = lds128(rowp + rr * ROWB);
                decode4_i2fp(cv.x, v); decode4_i2fp(cv.y, v + 4); decode4_i2fp(cv.z, v + 8); decode4_i2fp(cv.w, v + 12);
            }
            if (MODE == 8) {
                const uint4 cv = lds128(rowp + rr * ROWB);
                decode4_i2fp(cv.x, v); decode4(cv.y, v + 4); decode4_i2fp(cv.z, v + 8); decode4(cv.w, v + 12);
            }
            if (MODE == 6) {
                const uint4 cv = lds128(rowp + rr * ROWB);
                decode4_i2f(cv.x, v); decode4(cv.y, v + 4); decode4(cv.z, v + 8); decode4(cv.w, v + 12);
            }
#pragma unroll
            for (int e = 0; e < 16; ++e)
#pragma unroll
                for (int g = 0; g < 4; ++g) q[e][g] = fma2(make_float2(v[e], v[e]), al[g], q[e][g]);
        }
    } else {
        const uint32_t last = (uint32_t)(nrows - 1);
        { const uint4 cv = lds128(rowp); decode4(cv.x, v); decode4(cv.y, v + 4); decode4(cv.z, v + 8); decode4(cv.w, v + 12); }
        uint4 cvA = lds128(rowp + ROWB), cvB;
        uint4 aA0 = lds128(abase), aA1 = lds128(abase + 16), aB0, aB1;
        auto row = [&](const uint4& a0, const uint4& a1, uint4& b0, uint4& b1, const uint4& cdec, uint4& cld, uint32_t r) {
            const uint32_t r1 = min(r + 1, last) & (ROWS - 1), r2 = min(r + 2, last) & (ROWS - 1);
            cld = lds128(rowp + r2 * ROWB);
            b0 = lds128(abase + r1 * 32); b1 = lds128(abase + r1 * 32 + 16);
            float2 a[4];
            a[0] = make_float2(__uint_as_float(a0.x), __uint_as_float(a0.y));
            a[1] = make_float2(__uint_as_float(a0.z), __uint_as_float(a0.w));
            a[2] = make_float2(__uint_as_float(a1.x), __uint_as_float(a1.y));
            a[3] = make_float2(__uint_as_float(a1.z), __uint_as_float(a1.w));
            const uint32_t wn[4] = {cdec.x, cdec.y, cdec.z, cdec.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int e = 4 * i; e < 4 * i + 4; ++e)
#pragma unroll
                    for (int g = 0; g < 4; ++g) q[e][g] = fma2(make_float2(v[e], v[e]), a[g], q[e][g]);
                decode4(wn[i], v + 4 * i);
            }
        };
#pragma unroll 1
        for (uint32_t r = 0; r <= last; r += 2) {
            row(aA0, aA1, aB0, aB1, cvA, cvB, r);
            if (r + 1 > last) break;
            row(aB0, aB1, aA0, aA1, cvB, cvA, r + 1);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e)
#pragma unroll
        for (int g = 0; g < 4; ++g) s += q[e][g].x + q[e][g].y;
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
static void run(const char* name, float* out, long long* cyc) {
    const int smem = ROWS * ROWB + ROWS * 32;
    cudaFuncSetAttribute(k<MODE, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<MODE, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int nrows = 8192;
    printf("%-44s", name);
    for (int threads : {128, 256, 384}) {
        long long c = 0;
        for (int rep = 0; rep < 3; ++rep) {
            if (threads <= 256) k<MODE, 256><<<148, threads, smem>>>(nrows, out, cyc);
            else k<MODE, 384><<<148, threads, smem>>>(nrows, out, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const int nw = threads / 128;
        // per sub-partition and row: nw warps x 64 FFMA2 need nw * 128 pipe cycles
        printf("  %d warp/SMSP: %6.1f cyc/row (FFMA2 floor %d, %4.1f%%)", nw, (double)c / nrows, nw * 128, 100.0 * nw * 128 * nrows / c);
    }
    printf("\n");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 384 * 4); cudaMalloc(&cyc, 8);
    run<0>("0 FFMA2 only", out, cyc);
    run<1>("1 + deltas from shared memory", out, cyc);
    run<2>("2 + codes from shared memory, decode (plain)", out, cyc);
    run<3>("3 software-pipelined two-row body", out, cyc);
    run<4>("4 plain, decode = 16 I2F.S8", out, cyc);
    run<5>("5 plain, decode = 8 I2F.S8 + 8 PRMT/4 FADD2", out, cyc);
    run<6>("6 plain, decode = 4 I2F.S8 + 12 PRMT/6 FADD2", out, cyc);
    run<7>("7 plain, decode = 16 PRMT.sext + 16 I2FP", out, cyc);
    run<8>("8 plain, 8 PRMT.sext/I2FP + 8 PRMT/4 FADD2", out, cyc);
    return 0;
}
