// viprs_b200 -- shared device helpers (sm_100a): mbarrier / 1-D TMA bulk copy PTX wrappers, shared-memory
// release/acquire counters, LD element decoding (int8/int16 dequantised in registers, packed f32x2 math),
// small warp utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/viprs_b200.h"

namespace vb {

constexpr int WARP = 32;
constexpr int PMAX = 16;        // max rows per panel (one TMA bulk copy)
constexpr int NST_MAX = 4;      // max TMA ring depth
constexpr int RR = 64;          // per-row ring slots (row metadata, dot partials, eta_new, window coefficients);
                                // two rows RR apart are never resident together: RR > PMAX * NST_MAX - 1
constexpr int WW = 48;          // window coefficients kept per row: columns j+1 .. j+WW (cut_j - j - 1 <= 47)
constexpr int NBW = 8;          // bulk warps per CTA
constexpr int NBT = NBW * WARP;
constexpr int WIN = 33;         // row j's forward axpy is done by the chain warp for columns < cut_j,
                                // cut_j = ceil((j + WIN) / EPV) * EPV  (block-local), by the bulk warps from cut_j on

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
#ifndef VB_TRYWAIT_NS
#define VB_TRYWAIT_NS 20000     // suspend-time hint of mbarrier.try_wait (ns)
#endif
// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) expires, so the
// waiting warps do not burn issue slots next to the working ones.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)VB_TRYWAIT_NS)
        : "memory");
    return ok != 0;
}
// non-blocking probe
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Every spin-wait in the sweep is bounded: a protocol bug traps (cudaErrorLaunchFailure) instead of hanging the GPU.
constexpr uint32_t kSpinLimit = 1u << 27;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > kSpinLimit) __trap();
    }
}
// the same on 32-bit shared-window addresses (no generic -> shared conversion in front of every barrier operation)
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"((uint32_t)VB_TRYWAIT_NS)
            : "memory");
        if (ok) break;
        if (++spins > kSpinLimit) __trap();
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
// 1-D TMA prefetch into L2 only (fire and forget): deep prefetch without spending shared memory.
__device__ __forceinline__ void tma_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// explicit shared-space accesses with 32-bit addresses (a generic `smem + offset` dereference makes ptxas
// rebuild the shared window base -- S2UR SR_CgaCtaId + ULEA -- in front of every load)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float lds_t(uint32_t addr, float) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_t(uint32_t addr, double) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_t(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_t(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// monotonic shared-memory counters with release / acquire semantics (CTA scope)
__device__ __forceinline__ void red_release_add(uint32_t* p, uint32_t v) {
    asm volatile("red.release.cta.shared.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_a(uint32_t addr, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_a(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void wait_ge(const uint32_t* p, uint32_t target) {
    uint32_t spins = 0;
    while (ld_acquire(p) < target) {
        if (++spins > kSpinLimit) __trap();
    }
}

// All per-warp progress counters reached their targets: prog[0..NA) >= ta, prog[NA..NA+NC) >= tc.
// Called by a full warp; lane l < NA+NC polls counter l.
template <int NA, int NC>
__device__ __forceinline__ void wait_progress(const uint32_t* prog, uint32_t ta, uint32_t tc, int lane) {
    uint32_t spins = 0;
    const uint32_t target = lane < NA ? ta : tc;
    for (;;) {
        const bool ok = (lane >= NA + NC) || (ld_acquire(prog + lane) >= target);
        if (__all_sync(0xffffffffu, ok)) break;
        if (++spins > kSpinLimit) __trap();
    }
}

// ---------------------------------------------------------------------------------------------
// packed fp32x2 math (sm_100: FFMA2 / FADD2 -- two IEEE fp32 operations per issue slot)
// ---------------------------------------------------------------------------------------------
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

// ---------------------------------------------------------------------------------------------
// LD element traits.  One "vector" is 16 bytes of LD data = EPV elements.
// Integer codes are stored BIASED on the device (int8: code ^ 0x80, int16: code ^ 0x8000; zero padding is
// 0x80 / 0x8000), so that a byte permute straight into the mantissa of 2^23 followed by one exact
// subtraction yields the code as fp32 -- no I2F, no sign fix-up.  The dequantisation scale is folded into
// scalars as in the reference (e_step.hpp:421 `dq_scale * eta_diff_j`).
// ---------------------------------------------------------------------------------------------
template <typename U> struct LdTraits;
template <> struct LdTraits<int8_t>  { static constexpr int EPV = 16; static constexpr int DT = VIPRS_B200_I8; };
template <> struct LdTraits<int16_t> { static constexpr int EPV = 8;  static constexpr int DT = VIPRS_B200_I16; };
template <> struct LdTraits<float>   { static constexpr int EPV = 4;  static constexpr int DT = VIPRS_B200_F32; };
template <> struct LdTraits<double>  { static constexpr int EPV = 2;  static constexpr int DT = VIPRS_B200_F64; };

constexpr float kMagicI8 = 8388736.0f;    // 2^23 + 128
constexpr float kMagicI16 = 8421376.0f;   // 2^23 + 32768

// accumulator type of a 16-byte-vector dot product
template <typename T> struct Pk;
template <> struct Pk<float> {
    using acc_t = float2;
    static __device__ __forceinline__ acc_t zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ float sum(acc_t a) { return a.x + a.y; }
};
template <> struct Pk<double> {
    using acc_t = double;
    static __device__ __forceinline__ acc_t zero() { return 0.0; }
    static __device__ __forceinline__ double sum(acc_t a) { return a; }
};

// VecOps<T,U>: dot(c, x, acc): acc += sum_e code_e * x[e];  axpy(c, a, f): f[e] += code_e * a
template <typename T, typename U> struct VecOps;

template <> struct VecOps<float, int8_t> {
    static __device__ __forceinline__ void pairs(uint32_t w, float2& p0, float2& p1) {
        p0 = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540)),
                         __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7541)));
        p1 = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7542)),
                         __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7543)));
        const float2 m = make_float2(-kMagicI8, -kMagicI8);
        p0 = add2(p0, m);
        p1 = add2(p1, m);
    }
    static __device__ __forceinline__ void dot(const uint4& c, const float* x, float2& acc) {
        const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 p0, p1;
            pairs(w[i], p0, p1);
            acc = fma2(p0, make_float2(x[4 * i], x[4 * i + 1]), acc);
            acc = fma2(p1, make_float2(x[4 * i + 2], x[4 * i + 3]), acc);
        }
    }
    static __device__ __forceinline__ void axpy(const uint4& c, float a, float* f) {
        const uint32_t w[4] = {c.x, c.y, c.z, c.w};
        const float2 aa = make_float2(a, a);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 p0, p1;
            pairs(w[i], p0, p1);
            const float2 r0 = fma2(p0, aa, make_float2(f[4 * i], f[4 * i + 1]));
            const float2 r1 = fma2(p1, aa, make_float2(f[4 * i + 2], f[4 * i + 3]));
            f[4 * i] = r0.x; f[4 * i + 1] = r0.y; f[4 * i + 2] = r1.x; f[4 * i + 3] = r1.y;
        }
    }
};
template <> struct VecOps<float, int16_t> {
    static __device__ __forceinline__ float2 pair(uint32_t w) {
        float2 p = make_float2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7410)),
                               __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7432)));
        return add2(p, make_float2(-kMagicI16, -kMagicI16));
    }
    static __device__ __forceinline__ void dot(const uint4& c, const float* x, float2& acc) {
        const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) acc = fma2(pair(w[i]), make_float2(x[2 * i], x[2 * i + 1]), acc);
    }
    static __device__ __forceinline__ void axpy(const uint4& c, float a, float* f) {
        const uint32_t w[4] = {c.x, c.y, c.z, c.w};
        const float2 aa = make_float2(a, a);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 r = fma2(pair(w[i]), aa, make_float2(f[2 * i], f[2 * i + 1]));
            f[2 * i] = r.x; f[2 * i + 1] = r.y;
        }
    }
};
template <> struct VecOps<float, float> {
    static __device__ __forceinline__ void dot(const uint4& c, const float* x, float2& acc) {
        acc = fma2(make_float2(__uint_as_float(c.x), __uint_as_float(c.y)), make_float2(x[0], x[1]), acc);
        acc = fma2(make_float2(__uint_as_float(c.z), __uint_as_float(c.w)), make_float2(x[2], x[3]), acc);
    }
    static __device__ __forceinline__ void axpy(const uint4& c, float a, float* f) {
        const float2 aa = make_float2(a, a);
        const float2 r0 = fma2(make_float2(__uint_as_float(c.x), __uint_as_float(c.y)), aa, make_float2(f[0], f[1]));
        const float2 r1 = fma2(make_float2(__uint_as_float(c.z), __uint_as_float(c.w)), aa, make_float2(f[2], f[3]));
        f[0] = r0.x; f[1] = r0.y; f[2] = r1.x; f[3] = r1.y;
    }
};
// double state: plain scalar decode + DFMA (these configurations are HBM-bound by a wide margin)
template <typename U> struct DecodeD;
template <> struct DecodeD<int8_t> {
    static __device__ __forceinline__ void vec(const uint4& v, double* o) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b) o[4 * i + b] = (double)((int)((w[i] >> (8 * b)) & 0xffu) - 128);
    }
};
template <> struct DecodeD<int16_t> {
    static __device__ __forceinline__ void vec(const uint4& v, double* o) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = (double)((int)(w[i] & 0xffffu) - 32768);
            o[2 * i + 1] = (double)((int)(w[i] >> 16) - 32768);
        }
    }
};
template <> struct DecodeD<float> {
    static __device__ __forceinline__ void vec(const uint4& v, double* o) {
        o[0] = (double)__uint_as_float(v.x); o[1] = (double)__uint_as_float(v.y);
        o[2] = (double)__uint_as_float(v.z); o[3] = (double)__uint_as_float(v.w);
    }
};
template <> struct DecodeD<double> {
    static __device__ __forceinline__ void vec(const uint4& v, double* o) {
        o[0] = __hiloint2double((int)v.y, (int)v.x);
        o[1] = __hiloint2double((int)v.w, (int)v.z);
    }
};
template <typename U> struct VecOps<double, U> {
    static constexpr int EPV = LdTraits<U>::EPV;
    static __device__ __forceinline__ void dot(const uint4& c, const double* x, double& acc) {
        double v[EPV];
        DecodeD<U>::vec(c, v);
#pragma unroll
        for (int e = 0; e < EPV; ++e) acc = fma(v[e], x[e], acc);
    }
    static __device__ __forceinline__ void axpy(const uint4& c, double a, double* f) {
        double v[EPV];
        DecodeD<U>::vec(c, v);
#pragma unroll
        for (int e = 0; e < EPV; ++e) f[e] = fma(v[e], a, f[e]);
    }
};

// scalar element fetch from the device layout (biased integer codes) -> value in the state type
template <typename T, typename U> __device__ __forceinline__ T ld_elem(const unsigned char* p);
template <> __device__ __forceinline__ float ld_elem<float, int8_t>(const unsigned char* p) { return (float)((int)*p - 128); }
template <> __device__ __forceinline__ float ld_elem<float, int16_t>(const unsigned char* p) {
    return (float)((int)*reinterpret_cast<const uint16_t*>(p) - 32768);
}
template <> __device__ __forceinline__ float ld_elem<float, float>(const unsigned char* p) { return *reinterpret_cast<const float*>(p); }
template <> __device__ __forceinline__ double ld_elem<double, int8_t>(const unsigned char* p) { return (double)((int)*p - 128); }
template <> __device__ __forceinline__ double ld_elem<double, int16_t>(const unsigned char* p) {
    return (double)((int)*reinterpret_cast<const uint16_t*>(p) - 32768);
}
template <> __device__ __forceinline__ double ld_elem<double, float>(const unsigned char* p) { return (double)*reinterpret_cast<const float*>(p); }
template <> __device__ __forceinline__ double ld_elem<double, double>(const unsigned char* p) { return *reinterpret_cast<const double*>(p); }

// fused multiply-add in the state type: std::fma in the reference (e_step.hpp:101,173)
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float exp_t(float x) { return expf(x); }
__device__ __forceinline__ double exp_t(double x) { return exp(x); }
__device__ __forceinline__ float abs_t(float x) { return fabsf(x); }
__device__ __forceinline__ double abs_t(double x) { return fabs(x); }
__device__ __forceinline__ float rcp_t(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double rcp_t(double x) { return 1.0 / x; }

// rounding-explicit scalar ops (no FMA contraction: the chain warp evaluates the same update twice -- once on
// the critical path, once for the outputs -- and both must give the same bits)
__device__ __forceinline__ float mul_t(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_t(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_t(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_t(double a, double b) { return __dadd_rn(a, b); }
// exp(-|x|) and 1/x on the per-SNP critical path.  float: MUFU.EX2 / MUFU.RCP (<= 2 ulp; the argument scaling
// adds |x| * 6e-8 relative error to e, i.e. < 3e-7 absolute in gamma) ; double: full precision.
__device__ __forceinline__ float expneg_t(float ax) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(ax, -1.4426950408889634f)));
    return r;
}
__device__ __forceinline__ double expneg_t(double ax) { return exp(-ax); }
__device__ __forceinline__ float fast_rcp_t(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double fast_rcp_t(double x) { return 1.0 / x; }

// stable sigmoid, e_step.hpp:245-261 (two-branch form evaluated branch-free: e = exp(-|x|),
// x >= 0: 1/(1+e);  x < 0: e/(1+e))
template <typename T>
__device__ __forceinline__ T sigmoid_t(T x) {
    const T e = expneg_t(abs_t(x));
    const T num = x < T(0) ? e : T(1);
    return mul_t(num, fast_rcp_t(add_t(T(1), e)));
}

// the same sigmoid for a logit given in base-2 units (x * log2 e): e = 2^-|x2| straight from MUFU.EX2
__device__ __forceinline__ float sigmoid2_t(float x2) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-fabsf(x2)));
    const float num = x2 < 0.f ? e : 1.f;
    return __fmul_rn(num, fast_rcp_t(__fadd_rn(1.f, e)));
}
__device__ __forceinline__ double sigmoid2_t(double x2) { return sigmoid_t(x2 * 0.6931471805599453); }

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

// Reduce 8 per-lane accumulators across the warp with 9 shuffles: on return acc[0] of lane l holds the
// warp total of accumulator  r = 4*bit4(l) + 2*bit3(l) + bit2(l).
template <typename T>
__device__ __forceinline__ int warp_reduce8(T* acc, int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const T send = b4 ? acc[i] : acc[i + 4];
        const T keep = b4 ? acc[i + 4] : acc[i];
        acc[i] = keep + shfl_xor_t(send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const T send = b3 ? acc[i] : acc[i + 2];
        const T keep = b3 ? acc[i + 2] : acc[i];
        acc[i] = keep + shfl_xor_t(send, 8);
    }
    {
        const T send = b2 ? acc[0] : acc[1];
        const T keep = b2 ? acc[1] : acc[0];
        acc[0] = keep + shfl_xor_t(send, 4);
    }
    acc[0] += shfl_xor_t(acc[0], 2);
    acc[0] += shfl_xor_t(acc[0], 1);
    return (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
}

// Reduce 4 per-lane accumulators across the warp with 6 shuffles: on return acc[0] of lane l holds the warp
// total of accumulator  r = 2*bit4(l) + bit3(l).
template <typename T>
__device__ __forceinline__ int warp_reduce4(T* acc, int lane) {
    const bool b4 = lane & 16, b3 = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const T send = b4 ? acc[i] : acc[i + 2];
        const T keep = b4 ? acc[i + 2] : acc[i];
        acc[i] = keep + shfl_xor_t(send, 16);
    }
    {
        const T send = b3 ? acc[0] : acc[1];
        const T keep = b3 ? acc[1] : acc[0];
        acc[0] = keep + shfl_xor_t(send, 8);
    }
    acc[0] += shfl_xor_t(acc[0], 4);
    acc[0] += shfl_xor_t(acc[0], 2);
    acc[0] += shfl_xor_t(acc[0], 1);
    return (b4 ? 2 : 0) + (b3 ? 1 : 0);
}

}  // namespace vb
