// common.cuh -- precision traits, softened inverse distances, sm_100a async-copy helpers.
//
// Part of tupan_b200: B200-native replacements for the pairwise kernels of ggf84/tupan
// (reference: tupan/lib/src).  The routines of THIS file are not derived from the reference
// sources; the reference lines cited in comments say which behaviour a routine has to reproduce.
// (Two files of the package do restate reference formulas closely and say so in their headers:
// pn_ops.cuh -- the post-Newtonian polynomials -- and kepler.cuh -- the universal-variable solver.)
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
// Masked reciprocal square root.
//
// The reference masks a pair when `r2 > 0` is false (e.g. acc_jerk_kernel_common.h:36) and
// forces inv_r2 = 0 after the divide (smoothing.h:140-143).  Here r2 = rx^2+ry^2+rz^2 is kept
// separate from the softened x = r2 + e2 and the mask is read off r2 itself.
//
// What shapes this code (measured on B200, tools/microbench.cu + tools/kernel_lab.cu):
//   * the FP64 pipe takes one warp instruction per 2 cycles per SM sub-partition, and every
//     OTHER instruction issued costs about one more cycle -- integer mask arithmetic is not
//     free in the shadow of the FP64 pipe.  Testing the six words of (rx, ry, rz) cost 7
//     integer instructions per pair; the exponent test of r2 costs 2 (ISETP + SEL);
//   * a DFMA with three distinct register operands takes 3 cycles (64-bit operands, two
//     register banks), with two it takes 2: the refinement below has none with three.
//
// fp64: MUFU.RSQ64H seed (PTX rsqrt.approx.ftz.f64; measured max rel. error 2^-20.06; reads
// the high word only) and ONE third-order step
//     h = 1 - x*y0^2 ,  y = y0 * (k0 + h*(k1 + k2*h)),   (k0,k1,k2) = k*(1, 1/2, 3/8)
// = 5 FP64 instructions (DMUL, DFMA, DFMA, DFMA, DMUL) returning k/sqrt(x) -- a constant
// factor k rides along for free (acc_jerk uses k = sqrt 3).  Remaining error ~ 5/16 h^3 < 2^-58.
// The mask zeroes the seed's HIGH word with an integer select: r2 < 2^-1022 (zero or
// denormal, where the reference's r2 > 0 divide would overflow anyway) -> seed ~ 0 -> every
// product downstream is exactly 0, no 0*inf is ever formed.  x <= 0 implies r2 fails the test
// as long as e2 >= 0, so the seed is never inf/NaN for an unmasked pair.
//   CLEAN = true : the seed's low word is zeroed too (one more MOV): masked r1 is exactly 0.
//   CLEAN = false: the low word keeps the raw MUFU bits (perturbs an unmasked seed by < 2^-20,
//                  absorbed by the cubic step); a masked r1 is a denormal ~1e-314 whose square
//                  underflows to exactly 0 -- for kernels that use r1 only through r1^2, r1^3.
//
// fp32: MUFU.RSQ (rsqrt.approx.ftz.f32, max rel. error 2^-22.4), select on r2 > 0, scaled by k0;
// no refinement (see below; -DTUPAN_FP32_NEWTON restores one Newton step y0 * (k0 + k1*h)).
// ---------------------------------------------------------------------------------------
template <bool CLEAN>
TUPAN_DEV double rsqrt_seed_masked(double x, double r2)
{
    double y0;
    if (CLEAN) {
        asm("{\n"
            ".reg .pred p;\n"
            ".reg .f64 y;\n"
            ".reg .b32 lo, hi, yl, yh;\n"
            "mov.b64 {lo, hi}, %2;\n"
            "setp.ge.u32 p, hi, 0x00100000;\n"
            "rsqrt.approx.ftz.f64 y, %1;\n"
            "mov.b64 {yl, yh}, y;\n"
            "selp.b32 yh, yh, 0, p;\n"
            "mov.b64 %0, {0, yh};\n"
            "}\n"
            : "=d"(y0)
            : "d"(x), "d"(r2));
    } else {
        asm("{\n"
            ".reg .pred p;\n"
            ".reg .f64 y;\n"
            ".reg .b32 lo, hi, yl, yh, ys;\n"
            "mov.b64 {lo, hi}, %2;\n"
            "setp.ge.u32 p, hi, 0x00100000;\n"
            "rsqrt.approx.ftz.f64 y, %1;\n"
            "mov.b64 {yl, yh}, y;\n"
            "selp.b32 ys, yh, 0, p;\n"
            "mov.b64 %0, {yh, ys};\n"
            "}\n"
            : "=d"(y0)
            : "d"(x), "d"(r2));
    }
    return y0;
}

// k/sqrt(x), masked by r2 (see above).  k0 = k, k1 = k/2, k2 = 3k/8.
template <bool CLEAN>
TUPAN_DEV double rsqrt_scaled(double x, double r2, double k0, double k1, double k2)
{
    const double y0 = rsqrt_seed_masked<CLEAN>(x, r2);
    const double t = x * y0;
    const double h = fma(-t, y0, 1.0);
    const double p = fma(h, k2, k1);
    const double c = fma(h, p, k0);
    return y0 * c;
}
template <bool CLEAN>
TUPAN_DEV float rsqrt_scaled(float x, float r2, float k0, float k1, float)
{
    float y0;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(x));
    y0 = (r2 > 0.0f) ? y0 : 0.0f;
#ifdef TUPAN_FP32_NEWTON
    const float h = fmaf(-x * y0, y0, 1.0f);
    return y0 * fmaf(h, k1, k0);
#else
    // MUFU.RSQ is good to 2^-22.4 (1.5 ulp); the fp32 kernels are issue-bound (31 FP32 + 4 other
    // instructions per pair, tools/microbench_fp32.cu) and a Newton step costs 3-4 of them for
    // half an ulp per pair, far below the rounding of the j sum itself.  north_star asks for the
    // refinement on the fp64 paths only.
    (void)k1;
    return y0 * k0;
#endif
}

// ---------------------------------------------------------------------------------------
// The same seed and step for the G pairs of a group (pair_kernel_grouped), written operation by
// operation, with the softening of the i-particle FOLDED into the r2 chain:
//     c = fma(rz, rz, fma(ry, ry, fma(rx, rx, e2_i))),   x = c + e2_j
// is one FP64 instruction per pair less than r2 + (e2_i + e2_j).  The chain c no longer holds the
// bare r2 the reference's mask reads (`r2 > 0`, e.g. acc_jerk_kernel_common.h:36 -- it applies to
// coincident particles even when they are softened), so the group tests a NECESSARY condition for
// a masked pair instead -- the high word of c still equals the high word of e2_i, i.e.
// r2 < 2^-20 e2_i or both zero -- with one predicate-chained ISETP per pair, and only a group that
// has such a pair forms r2 again (bare_r2(p)) and zeroes the seeds the mask really applies to, by
// rsqrt_seed_masked's criterion (r2 zero or denormal).  A particle meets itself once per sweep;
// pairs inside 1e-3 softening lengths are as rare.  The seed's low word is zero (CLEAN).
//   c[p]      the chain;  ei_hi(p)  high word of the pair's e2_i;  ej(p)  the pair's e2_j
//   x[p], y0[p]  out: softened r2 and the (masked) rsqrt seed
// tests/test_headline_gpu.py::test_grouped_kernel_mask_cases drives every way through the branch.
// ---------------------------------------------------------------------------------------
template <int G, class EI, class EJ, class R2>
TUPAN_DEV void group_fold_seeds(const double (&c)[G], EI ei_hi, EJ ej, R2 bare_r2, double (&x)[G], double (&y0)[G])
{
    bool cand = false;
#pragma unroll
    for (int p = 0; p < G; ++p) cand = cand || (__double2hiint(c[p]) == ei_hi(p));
#pragma unroll
    for (int p = 0; p < G; ++p) x[p] = c[p] + ej(p);
#pragma unroll
    for (int p = 0; p < G; ++p) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[p]) : "d"(x[p]));
    if (cand) {
#pragma unroll
        for (int p = 0; p < G; ++p) {
            const double q = bare_r2(p);
            if ((unsigned)__double2hiint(q) < 0x00100000u) y0[p] = 0.0;
        }
    }
}
template <int G, class EI, class EJ, class R2>
TUPAN_DEV void group_fold_seeds(const float (&)[G], EI, EJ, R2, float (&)[G], float (&)[G]) {}   // fp64 only

// y[p] = k / sqrt(x[p]) from the seeds: the cubic step of rsqrt_scaled, operation by operation.
template <int G, typename T>
TUPAN_DEV void group_rsqrt_step(const T (&x)[G], const T (&y0)[G], T k0, T k1, T k2, T (&y)[G])
{
    T t[G], h[G];
#pragma unroll
    for (int p = 0; p < G; ++p) t[p] = x[p] * y0[p];
#pragma unroll
    for (int p = 0; p < G; ++p) h[p] = fma(-t[p], y0[p], T(1));
#pragma unroll
    for (int p = 0; p < G; ++p) t[p] = fma(h[p], k2, k1);
#pragma unroll
    for (int p = 0; p < G; ++p) t[p] = fma(h[p], t[p], k0);
#pragma unroll
    for (int p = 0; p < G; ++p) y[p] = y0[p] * t[p];
}

// k/sqrt(x) for an x >= 0 whose zero must give a FINITE result rather than a masked one (tstep's second
// seed: x = w2 is exactly 0 for a masked pair, and so is the factor the result scales).  One integer
// minimum clamps the seed's high word at 2^511 = rsqrt(2^-1022), the seed of the smallest normal x, instead
// of the compare + select of rsqrt_seed_masked: every normal x keeps its seed, x = 0 (seed +inf) gets
// 2^511, and 0 * (2^511 k) = 0 downstream.  The low word keeps the raw high bits, as with CLEAN = false.
TUPAN_DEV double rsqrt_scaled_clamped(double x, double k0, double k1, double k2)
{
    double y0;
    asm("{\n"
        ".reg .f64 y;\n"
        ".reg .b32 yl, yh, yc;\n"
        "rsqrt.approx.ftz.f64 y, %1;\n"
        "mov.b64 {yl, yh}, y;\n"
        "min.u32 yc, yh, 0x5fe00000;\n"
        "mov.b64 %0, {yh, yc};\n"
        "}\n"
        : "=d"(y0)
        : "d"(x));
    const double t = x * y0;
    const double h = fma(-t, y0, 1.0);
    const double p = fma(h, k2, k1);
    const double c = fma(h, p, k0);
    return y0 * c;
}
TUPAN_DEV float rsqrt_scaled_clamped(float x, float k0, float k1, float k2) { return rsqrt_scaled<false>(x, x, k0, k1, k2); }

template <typename T> struct InvR { T r1, r2, r3; };

// x = r2 + e2 (softened).  Returns 1/r, 1/r^2, 1/r^3 of the softened distance, all 0 when the
// pair is masked (CLEAN = false: r1 is a denormal instead, see above).  Unused members are
// dead-code-eliminated.
template <bool CLEAN, typename T>
TUPAN_DEV InvR<T> soft_inv(T x, T r2)
{
    InvR<T> o;
    o.r1 = rsqrt_scaled<CLEAN>(x, r2, T(1), T(0.5), T(0.375));
    o.r2 = o.r1 * o.r1;
    o.r3 = o.r2 * o.r1;
    return o;
}

// 1/x without the IEEE-division slow path (a call inside a hot loop): MUFU.RCP64H seed (20
// bits) and one third-order step, y (1 + e + e^2) with e = 1 - x y; error ~ e^3 < 2^-58.
TUPAN_DEV double rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}
TUPAN_DEV float rcp_fast(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

TUPAN_DEV double rmax(double a, double b) { return fmax(a, b); }
// max(acc, w) for a running maximum that starts at +0 and therefore never is negative: the
// bit patterns of non-negative doubles order like the numbers, and any negative w (sign bit set) is
// a negative integer, so a SIGNED 64-bit integer maximum is exactly fmax here -- two ISETP and two
// SEL instead of DSETP (an FP64-pipe instruction) + FSEL + SEL + LOP3 + moves (tstep, ops.cuh).
TUPAN_DEV double rmax_nonneg(double acc, double w)
{
    const long long a = __double_as_longlong(acc), b = __double_as_longlong(w);
    return __longlong_as_double(a > b ? a : b);
}
TUPAN_DEV float rmax_nonneg(float acc, float w)
{
    const int a = __float_as_int(acc), b = __float_as_int(w);
    return __int_as_float(a > b ? a : b);
}
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
