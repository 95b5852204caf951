// viprs_b200 -- shared device helpers (sm_100a): mbarrier / 1-D TMA bulk copy PTX wrappers,
// LD element decoding (int8/int16 dequantised in registers), small warp utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/viprs_b200.h"

namespace vb {

constexpr int WARP = 32;
constexpr int PMAX = 16;        // rows per panel (chain-warp window is 2*PMAX = 32 lanes)
constexpr int NSTAGE = 4;       // TMA ring depth (panels u-2 (axpy), u-1 (chain), u (dot), u+1 (in flight))
constexpr int NSLOT = 4;        // ring depth of the small per-panel mailboxes (dot partials, eta_new)

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk-copy wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D TMA: global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// LD element traits.  One "vector" is 16 bytes of LD data = EPV elements.
// decode() turns the 16 bytes into EPV values of the state type T (exact for integer codes:
// the dequantisation scale is folded into scalars, as in e_step.hpp:421 `dq_scale*eta_diff_j`).
// ---------------------------------------------------------------------------------------------
template <typename U> struct LdTraits;
template <> struct LdTraits<int8_t>  { static constexpr int EPV = 16; static constexpr int DT = VIPRS_B200_I8; };
template <> struct LdTraits<int16_t> { static constexpr int EPV = 8;  static constexpr int DT = VIPRS_B200_I16; };
template <> struct LdTraits<float>   { static constexpr int EPV = 4;  static constexpr int DT = VIPRS_B200_F32; };
template <> struct LdTraits<double>  { static constexpr int EPV = 2;  static constexpr int DT = VIPRS_B200_F64; };

// int8 -> fp32 without I2F: place the (sign-flipped) byte in the low mantissa bits of 2^23 and
// subtract 2^23+128.  One PRMT + one FADD per element, both exact.
__device__ __forceinline__ void decode4_i8(uint32_t w, float* o) {
    const uint32_t x = w ^ 0x80808080u;
    o[0] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7540)) - 8388736.0f;
    o[1] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7541)) - 8388736.0f;
    o[2] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7542)) - 8388736.0f;
    o[3] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7543)) - 8388736.0f;
}
__device__ __forceinline__ void decode2_i16(uint32_t w, float* o) {
    const uint32_t x = w ^ 0x80008000u;
    o[0] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7410)) - 8421376.0f;
    o[1] = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7432)) - 8421376.0f;
}

template <typename T, typename U> struct Decode;

template <> struct Decode<float, int8_t> {
    static __device__ __forceinline__ void vec(const uint4& v, float* o) {
        decode4_i8(v.x, o); decode4_i8(v.y, o + 4); decode4_i8(v.z, o + 8); decode4_i8(v.w, o + 12);
    }
};
template <> struct Decode<float, int16_t> {
    static __device__ __forceinline__ void vec(const uint4& v, float* o) {
        decode2_i16(v.x, o); decode2_i16(v.y, o + 2); decode2_i16(v.z, o + 4); decode2_i16(v.w, o + 6);
    }
};
template <> struct Decode<float, float> {
    static __device__ __forceinline__ void vec(const uint4& v, float* o) {
        o[0] = __uint_as_float(v.x); o[1] = __uint_as_float(v.y); o[2] = __uint_as_float(v.z); o[3] = __uint_as_float(v.w);
    }
};
template <> struct Decode<double, int8_t> {
    static __device__ __forceinline__ void vec(const uint4& v, double* o) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b) o[4 * i + b] = (double)(int)(int8_t)(w[i] >> (8 * b));
    }
};
template <> struct Decode<double, int16_t> {
    static __device__ __forceinline__ void vec(const uint4& v, double* o) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = (double)(int)(int16_t)(w[i] & 0xffffu);
            o[2 * i + 1] = (double)(int)(int16_t)(w[i] >> 16);
        }
    }
};
template <> struct Decode<double, float> {
    static __device__ __forceinline__ void vec(const uint4& v, double* o) {
        o[0] = (double)__uint_as_float(v.x); o[1] = (double)__uint_as_float(v.y);
        o[2] = (double)__uint_as_float(v.z); o[3] = (double)__uint_as_float(v.w);
    }
};
template <> struct Decode<double, double> {
    static __device__ __forceinline__ void vec(const uint4& v, double* o) {
        o[0] = __hiloint2double((int)v.y, (int)v.x);
        o[1] = __hiloint2double((int)v.w, (int)v.z);
    }
};

// scalar element fetch from shared memory (chain warp's window coefficients)
template <typename T, typename U>
__device__ __forceinline__ T ld_elem(const unsigned char* base, int byte_off) {
    return static_cast<T>(*reinterpret_cast<const U*>(base + byte_off));
}

// fused multiply-add in the state type: std::fma in the reference (e_step.hpp:101,173)
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float exp_t(float x) { return expf(x); }
__device__ __forceinline__ double exp_t(double x) { return exp(x); }
__device__ __forceinline__ float abs_t(float x) { return fabsf(x); }
__device__ __forceinline__ double abs_t(double x) { return fabs(x); }

// two-branch stable sigmoid, e_step.hpp:245-261
template <typename T>
__device__ __forceinline__ T sigmoid_t(T x) {
    if (x < T(0)) {
        const T ex = exp_t(x);
        return ex / (T(1) + ex);
    }
    return T(1) / (T(1) + exp_t(-x));
}

template <typename T>
__device__ __forceinline__ T shfl_t(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <typename T>
__device__ __forceinline__ T shfl_down_t(T v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
template <typename T>
__device__ __forceinline__ T shfl_xor_t(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_t(v, m);
    return v;
}

}  // namespace vb
