// kernel_lab.cu -- variants of the acc_jerk fp64 inner loop on the production pair engine,
// timed at exactly whole waves (steady state), to pick launch shape / mask strategy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DTUPAN_FP64 \
//        -o tools/bin/kernel_lab tools/kernel_lab.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../tupan_b200/csrc/ops.cuh"

using namespace tupan;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// MASK 0: production (masked seed).  MASK 1: predicated accumulation (inline PTX).
// MASK 2: no mask at all (upper bound; wrong for r = 0).
template <int W, int U, int MASK> struct AJ : AccJerkOp<double> {
    typedef double T;
    enum { WPT = W, UNROLL = U };
    static __device__ __forceinline__ void pair_v2(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA]);
    static __device__ __forceinline__ void pair_v3(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA]);
    static TUPAN_DEV void pack_j(const T* const* j, long long r, T (&row)[NJP])
    {
        if (MASK == 0) { AccJerkOp<double>::pack_j(j, r, row); return; }
        pack_row8(j, r, row);
        if (MASK >= 8) row[JM] = row[JM] * 0.19245008972987526;   // m / (3 sqrt 3)
    }
    static TUPAN_DEV void pair(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA], const Params& prm)
    {
        if (MASK == 0) { AccJerkOp<double>::pair(s, row, a, prm); return; }
        if (MASK >= 8) { pair_v3(s, row, a); return; }
        if (MASK >= 4) { pair_v2(s, row, a); return; }
        T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
        T vx = s[IVX] - row[J8_VX], vy = s[IVY] - row[J8_VY], vz = s[IVZ] - row[J8_VZ];
        T x = s[IE] + row[J8_E2];
        x = fma(rx, rx, x); x = fma(ry, ry, x); x = fma(rz, rz, x);
        T rv = rx * vx; rv = fma(ry, vy, rv); rv = fma(rz, vz, rv);
        T y0;
        if (MASK == 3) {
            // one predicate (separation non-zero AND seed finite), one select on the seed's high word
            asm("{\n"
                ".reg .pred p;\n"
                ".reg .f64 y;\n"
                ".reg .b32 lo, hi, yl, yh;\n"
                ".reg .b32 l0, l1, l2, h0, h1, h2;\n"
                "rsqrt.approx.ftz.f64 y, %1;\n"
                "mov.b64 {l0, h0}, %2;\n"
                "mov.b64 {l1, h1}, %3;\n"
                "mov.b64 {l2, h2}, %4;\n"
                "mov.b64 {yl, yh}, y;\n"
                "or.b32 lo, l0, l1;\n"
                "or.b32 lo, lo, l2;\n"
                "or.b32 hi, h0, h1;\n"
                "or.b32 hi, hi, h2;\n"
                "and.b32 hi, hi, 0x7fffffff;\n"
                "or.b32 lo, lo, hi;\n"
                "setp.ne.u32 p, lo, 0;\n"
                "setp.ne.and.u32 p, yh, 0x7ff00000, p;\n"
                "selp.b32 yh, yh, 0, p;\n"
                "mov.b64 %0, {0, yh};\n"
                "}\n"
                : "=d"(y0)
                : "d"(x), "d"(rx), "d"(ry), "d"(rz));
        } else {
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
        }
        T t = x * y0;
        T h = fma(-t, y0, 1.0);
        T p = fma(h, 0.375, 0.5);
        T q = y0 * h;
        T r1 = fma(q, p, y0);
        T r2 = r1 * r1;
        T r3 = r2 * r1;
        T alpha = (3.0 * r2) * rv;
        vx = fma(-alpha, rx, vx); vy = fma(-alpha, ry, vy); vz = fma(-alpha, rz, vz);
        T g = -(row[JM] * r3);
        if (MASK == 2 || MASK == 3) {
            a[0] = fma(g, rx, a[0]); a[1] = fma(g, ry, a[1]); a[2] = fma(g, rz, a[2]);
            a[3] = fma(g, vx, a[3]); a[4] = fma(g, vy, a[4]); a[5] = fma(g, vz, a[5]);
        } else {
            asm("{\n"
                ".reg .pred p;\n"
                ".reg .b32 lo, hi, xl, xh;\n"
                ".reg .b32 l0, l1, l2, h0, h1, h2;\n"
                "mov.b64 {l0, h0}, %7;\n"
                "mov.b64 {l1, h1}, %8;\n"
                "mov.b64 {l2, h2}, %9;\n"
                "mov.b64 {xl, xh}, %13;\n"
                "or.b32 lo, l0, l1;\n"
                "or.b32 lo, lo, l2;\n"
                "or.b32 hi, h0, h1;\n"
                "or.b32 hi, hi, h2;\n"
                "and.b32 hi, hi, 0x7fffffff;\n"
                "or.b32 lo, lo, hi;\n"
                "setp.ne.u32 p, lo, 0;\n"
                "setp.ge.and.u32 p, xh, 0x00100000, p;\n"
                "@p fma.rn.f64 %0, %6, %7, %0;\n"
                "@p fma.rn.f64 %1, %6, %8, %1;\n"
                "@p fma.rn.f64 %2, %6, %9, %2;\n"
                "@p fma.rn.f64 %3, %6, %10, %3;\n"
                "@p fma.rn.f64 %4, %6, %11, %4;\n"
                "@p fma.rn.f64 %5, %6, %12, %5;\n"
                "}\n"
                : "+d"(a[0]), "+d"(a[1]), "+d"(a[2]), "+d"(a[3]), "+d"(a[4]), "+d"(a[5])
                : "d"(g), "d"(rx), "d"(ry), "d"(rz), "d"(vx), "d"(vy), "d"(vz), "d"(x));
        }
    }
};

// MASK 4: r2 kept separate (x = r2 + e, one more FP64 op), mask = exponent field of r2 != 0
//         (one ISETP + one SEL on the seed); rsqrt step without a 3-register DFMA.
// MASK 5: guard only against x == 0 (no r == 0 mask when softened) -- not the reference semantics.
// MASK 6: as 4 but e2 taken as a launch constant (uniform softening), 32 FP64 ops.
// MASK 7: no mask, new rsqrt step.
template <int W, int U, int MASK>
__device__ __forceinline__ void AJ<W, U, MASK>::pair_v2(const double (&s)[NI], const double (&row)[NJP], double (&a)[NA])
{
    typedef double T;
    T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
    T vx = s[IVX] - row[J8_VX], vy = s[IVY] - row[J8_VY], vz = s[IVZ] - row[J8_VZ];
    T x, r2 = 0;
    if (MASK == 4 || MASK == 6) {
        r2 = rx * rx; r2 = fma(ry, ry, r2); r2 = fma(rz, rz, r2);
        x = r2 + (MASK == 6 ? 2.0 * s[IE] : s[IE] + row[J8_E2]);
    } else {
        x = s[IE] + row[J8_E2];
        x = fma(rx, rx, x); x = fma(ry, ry, x); x = fma(rz, rz, x);
    }
    T rv = rx * vx; rv = fma(ry, vy, rv); rv = fma(rz, vz, rv);
    T y0;
    if (MASK == 7) {
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    } else {
        const T probe = (MASK == 5) ? x : r2;
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
            : "d"(x), "d"(probe));
    }
    T t = x * y0;
    T h = fma(-t, y0, 1.0);
    T p = fma(h, 0.375, 0.5);
    T c = fma(h, p, 1.0);
    T r1 = y0 * c;
    T q2 = r1 * r1;
    T q3 = q2 * r1;
    T alpha = (3.0 * q2) * rv;
    T g = -(row[JM] * q3);
    vx = fma(-alpha, rx, vx); vy = fma(-alpha, ry, vy); vz = fma(-alpha, rz, vz);
    a[0] = fma(g, rx, a[0]); a[1] = fma(g, ry, a[1]); a[2] = fma(g, rz, a[2]);
    a[3] = fma(g, vx, a[3]); a[4] = fma(g, vy, a[4]); a[5] = fma(g, vz, a[5]);
}

// MASK 8: r2 separate, mask on the exponent of r2, seed low word left as the raw MUFU word
//         (no MOV), sqrt(3) folded into the rsqrt polynomial and 1/(3 sqrt 3) into the packed
//         mass (no x3 multiply): 32 FP64 ops, exact reference mask.
// MASK 9: as 8 with uniform softening as a launch constant: 31 FP64 ops.
// MASK 10: as 8 but accumulates written j-major (g-sharing DFMAs adjacent in the source).
template <int W, int U, int MASK>
__device__ __forceinline__ void AJ<W, U, MASK>::pair_v3(const double (&s)[NI], const double (&row)[NJP], double (&a)[NA])
{
    typedef double T;
    T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
    T vx = s[IVX] - row[J8_VX], vy = s[IVY] - row[J8_VY], vz = s[IVZ] - row[J8_VZ];
    T r2 = rx * rx; r2 = fma(ry, ry, r2); r2 = fma(rz, rz, r2);
    T x = r2 + (MASK == 9 ? 2.0 * s[IE] : s[IE] + row[J8_E2]);
    T rv = rx * vx; rv = fma(ry, vy, rv); rv = fma(rz, vz, rv);
    T y0;
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
    T t = x * y0;
    T h = fma(-t, y0, 1.0);
    T p = fma(h, 0.64951905283832900, 0.86602540378443865);     // sqrt3 * (3/8, 1/2)
    T c = fma(h, p, 1.7320508075688772);                         // sqrt3
    T r1 = y0 * c;            // sqrt(3/x)
    T q2 = r1 * r1;           // 3/x
    T q3 = q2 * r1;           // 3 sqrt3 x^-3/2
    T alpha = q2 * rv;
    T g = -(row[JM] * q3);
    vx = fma(-alpha, rx, vx); vy = fma(-alpha, ry, vy); vz = fma(-alpha, rz, vz);
    a[0] = fma(g, rx, a[0]); a[1] = fma(g, ry, a[1]); a[2] = fma(g, rz, a[2]);
    a[3] = fma(g, vx, a[3]); a[4] = fma(g, vy, a[4]); a[5] = fma(g, vz, a[5]);
}

namespace tupan {
template <int W, int U, int M> struct LabTune { enum { NT = 256 }; };
}

template <class Op, int NT>
static double run_variant(const char* name, const InRefs<double>& in, long long n_alloc, const double* jpack,
                          long long nj, double* out[6], int sms, double* checksum)
{
    constexpr int TJ = 128, STAGES = 4;
    auto k = pair_kernel<Op, NT, Op::WPT, TJ, STAGES, false>;
    const size_t smem = PairSmem<Op, TJ, STAGES>::BYTES;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NT, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k));
    const long long IB = (long long)NT * Op::WPT;
    long long waves = 2;
    long long ni = waves * occ * sms * IB;
    while (ni > n_alloc && waves > 1) { waves--; ni = waves * occ * sms * IB; }
    if (ni > n_alloc) { printf("%-34s skipped (needs %lld particles)\n", name, ni); return 0; }
    pack_j_kernel<Op><<<296, 256>>>(in, nj, const_cast<double*>(jpack));
    CK(cudaDeviceSynchronize());
    PairArgs<Op> a;
    a.i = in; a.ni = ni; a.jpack = jpack; a.j0 = 0; a.j1 = nj; a.jchunk = nj; a.js_log2 = 0; a.slot0 = 0;
    a.partial = nullptr;
    for (int q = 0; q < MAX_OUT; ++q) a.out.p[q] = q < 6 ? out[q] : nullptr;
    dim3 grid((unsigned)(ni / IB), 1);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k<<<grid, NT, smem>>>(a);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k<<<grid, NT, smem>>>(a);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    std::vector<double> h(1024);
    double cs = 0;
    for (int q = 0; q < 6; ++q) {
        CK(cudaMemcpy(h.data(), out[q], 1024 * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 1024; ++i) cs += fabs(h[i]);
    }
    *checksum = cs;
    const double gp = (double)ni * nj / (best * 1e-3) * 1e-9;
    printf("%-34s regs=%3d occ=%d ni=%7lld  %8.3f ms  %7.1f Gpair/s  %5.2f TF  checksum %.12e\n", name, fa.numRegs, occ,
           ni, best, gp, gp * 42e-3, cs);
    return gp;
}

int main(int argc, char** argv)
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const long long n = 4LL * 2 * sms * 1024;     // enough i-particles for every variant
    const long long nj = argc > 1 ? atoll(argv[1]) : 32768;
    std::vector<double> h(8 * n);
    srand(1);
    for (long long i = 0; i < n; ++i) {
        h[0 * n + i] = 1.0 / n;
        for (int k = 1; k <= 3; ++k) h[k * n + i] = (double)rand() / RAND_MAX - 0.5;
        h[4 * n + i] = 1e-6;
        for (int k = 5; k <= 7; ++k) h[k * n + i] = (double)rand() / RAND_MAX - 0.5;
    }
    double* d;
    CK(cudaMalloc(&d, 8 * n * sizeof(double)));
    CK(cudaMemcpy(d, h.data(), 8 * n * sizeof(double), cudaMemcpyHostToDevice));
    InRefs<double> in;
    for (int k = 0; k < MAX_IN; ++k) in.p[k] = k < 8 ? d + k * n : nullptr;
    double* jpack;
    CK(cudaMalloc(&jpack, nj * 8 * sizeof(double)));
    pack_j_kernel<AccJerkOp<double>><<<296, 256>>>(in, nj, jpack);
    CK(cudaDeviceSynchronize());
    double* out[6];
    for (int q = 0; q < 6; ++q) CK(cudaMalloc(&out[q], n * sizeof(double)));
    double cs;
    printf("%s, %d SMs, nj = %lld\n", p.name, sms, nj);
#define RUN(W, U, M, NT) run_variant<AJ<W, U, M>, NT>("W" #W " U" #U " mask" #M " NT" #NT, in, n, jpack, nj, out, sms, &cs)
    RUN(2, 4, 0, 256);
    RUN(2, 4, 2, 256);
    RUN(2, 4, 3, 256);
    RUN(2, 4, 8, 256);
    RUN(2, 4, 9, 256);
    RUN(2, 2, 8, 256);
    RUN(2, 2, 9, 256);
    RUN(2, 1, 8, 256);
    RUN(2, 8, 8, 256);
    RUN(1, 8, 8, 256);
    RUN(3, 2, 8, 256);
    RUN(4, 2, 8, 128);
    RUN(4, 1, 8, 128);
    RUN(2, 4, 8, 128);
    RUN(2, 4, 4, 256);
    RUN(2, 4, 5, 256);
    RUN(2, 4, 6, 256);
    RUN(2, 4, 7, 256);
    RUN(2, 2, 4, 256);
    RUN(2, 2, 6, 256);
    RUN(1, 8, 4, 256);
    RUN(1, 8, 6, 256);
    RUN(4, 2, 4, 128);
    RUN(4, 2, 6, 128);
    RUN(3, 2, 4, 256);
    RUN(3, 2, 6, 256);
    RUN(4, 2, 3, 256);
    RUN(4, 2, 2, 256);
    RUN(1, 8, 3, 256);
    RUN(1, 8, 2, 256);
    RUN(1, 8, 0, 256);
    RUN(1, 4, 3, 256);
    RUN(2, 2, 3, 256);
    RUN(2, 8, 3, 256);
    RUN(4, 2, 0, 256);
    RUN(4, 1, 3, 256);
    RUN(4, 2, 3, 128);
    RUN(2, 4, 3, 128);
    RUN(3, 2, 3, 256);
    RUN(2, 4, 3, 512);
    RUN(1, 8, 3, 512);
    return 0;
}
