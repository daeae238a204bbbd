// common.cuh -- precision traits, softened inverse distances, sm_100a async-copy helpers.
//
// Part of tupan_b200: B200-native replacements for the pairwise kernels of ggf84/tupan
// (reference: tupan/lib/src).  Nothing here is derived from the reference sources; the
// reference lines cited in comments say which behaviour a routine has to reproduce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tupan {

#define TUPAN_DEV __device__ __forceinline__

// ---------------------------------------------------------------------------------------
// Precision traits.  One shared object per precision, as in the reference
// (cffi_backend.py:35-36,84-87): fp64 lib -> REAL=double, UINT=unsigned long, INT=long;
// fp32 lib -> float, unsigned int, int.
// ---------------------------------------------------------------------------------------
#ifdef TUPAN_FP64
typedef double real_t;
typedef unsigned long abi_uint;
typedef long abi_int;
#else
typedef float real_t;
typedef unsigned int abi_uint;
typedef int abi_int;
#endif

template <typename T> struct Vec16;            // how many T fit a 16-byte shared-memory load
template <> struct Vec16<double> { enum { N = 2 }; typedef double2 type; };
template <> struct Vec16<float>  { enum { N = 4 }; typedef float4 type; };

constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------
// Zero test of the separation vector.  The reference masks a pair when r2 > 0 is false
// (e.g. acc_jerk_kernel_common.h:36).  We test the separation itself -- for doubles with
// integer ops, so the test stays off the FP64 pipe -- and additionally treat a pair whose
// softened r2+e2 is not a positive normal number (rsqrt seed = inf) as masked.  The two
// definitions differ only when r2 underflows to zero for a non-zero separation.
// ---------------------------------------------------------------------------------------
TUPAN_DEV bool nonzero3(double x, double y, double z)
{
    unsigned lo = (unsigned)__double2loint(x) | (unsigned)__double2loint(y) | (unsigned)__double2loint(z);
    unsigned hi = (unsigned)__double2hiint(x) | (unsigned)__double2hiint(y) | (unsigned)__double2hiint(z);
    return ((hi & 0x7fffffffu) | lo) != 0u;
}
TUPAN_DEV bool nonzero3(float x, float y, float z)
{
    return (x != 0.0f) | (y != 0.0f) | (z != 0.0f);
}

// ---------------------------------------------------------------------------------------
// Masked reciprocal square root.
//
// fp64: MUFU.RSQ64H seed (PTX rsqrt.approx.ftz.f64, ~20 bits, looks at the high word only,
// so it costs no FP64-pipe slot and no conversion) followed by ONE third-order step
//     h = 1 - x*y0^2 ,  y = y0 + y0*h*(1/2 + 3/8 h)            (error ~ 5/16 h^3 < 2^-58)
// = 5 FP64 instructions (DMUL, DFMA, DFMA, DMUL, DFMA).  The mask is applied to the SEED
// with an integer select: a zero seed stays exactly zero through the refinement, so a
// masked pair contributes exact zeros and no 0*inf is ever formed (the reference selects
// after the divide instead, smoothing.h:140-143).
//
// fp32: MUFU.RSQ (rsqrt.approx.ftz.f32, max rel. error 2^-22.4) and a select.
// ---------------------------------------------------------------------------------------
TUPAN_DEV double rsqrt_masked(double x, bool ok)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    int hi = __double2hiint(y0);
    ok = ok && (hi != 0x7ff00000);
    y0 = __hiloint2double(ok ? hi : 0, 0);
    double t = x * y0;
    double h = fma(-t, y0, 1.0);
    double p = fma(h, 0.375, 0.5);
    double q = y0 * h;
    return fma(q, p, y0);
}
TUPAN_DEV float rsqrt_masked(float x, bool ok)
{
    float y0;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(x));
    ok = ok && (x > 0.0f);
    // one Newton step: keeps fp32 results as close to the correctly rounded reference
    // chain (1/x, sqrt, *) as the accumulation order allows
    y0 = ok ? y0 : 0.0f;
    float h = fmaf(-x * y0, y0, 1.0f);
    return fmaf(0.5f * y0, h, y0);
}

template <typename T> struct InvR { T r1, r2, r3; };

// x = r2 + e2 (softened), ok = pair not masked.  Returns 1/r, 1/r^2, 1/r^3 (all exactly 0
// when masked).  Unused members are dead-code-eliminated.
template <typename T>
TUPAN_DEV InvR<T> soft_inv(T x, bool ok)
{
    InvR<T> o;
    o.r1 = rsqrt_masked(x, ok);
    o.r2 = o.r1 * o.r1;
    o.r3 = o.r2 * o.r1;
    return o;
}

TUPAN_DEV double rmax(double a, double b) { return fmax(a, b); }
TUPAN_DEV float rmax(float a, float b) { return fmaxf(a, b); }
TUPAN_DEV double rsqrt_full(double a) { return 1.0 / sqrt(a); }
TUPAN_DEV float rsqrt_full(float a) { return 1.0f / sqrtf(a); }

// ---------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -- sm_90+/sm_100a PTX.
// ---------------------------------------------------------------------------------------
TUPAN_DEV uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
TUPAN_DEV void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
TUPAN_DEV void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
TUPAN_DEV void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
TUPAN_DEV bool mbar_try_wait(uint64_t* bar, unsigned parity)
{
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
TUPAN_DEV void mbar_wait(uint64_t* bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
TUPAN_DEV void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace tupan
