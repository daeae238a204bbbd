// kernel_lab3.cu -- round 2: group shapes (W particles per thread x U rows, threads per CTA, GMODE bits) of a
// PRODUCTION grouped kernel (pair_kernel_grouped, pair_engine.cuh) next to its plain kernel (W = 0), timed
// at whole waves through the production kernel templates.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DTUPAN_FP64 \
//        [-DLAB_OP=1..4] [-DLAB_LIST_FILE='"list.h"'] [-Xptxas -regUsageLevel -Xptxas 10] \
//        -o tools/bin/kernel_lab3 tools/kernel_lab3.cu
// LAB_OP: 0 acc_jerk (default; profiles/r02_kernel_lab3_*.txt), 1 acc (r02_kernel_lab4_acc.txt), 2 phi
// (r02_kernel_lab5_phi.txt), 3 tstep (r02_kernel_lab6_tstep.txt), 4 nreg_X (r02_kernel_lab7_nregx.txt).
// LAB_LIST_FILE defines LAB_LIST as a sequence of X(W, U, NT, MODE).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../tupan_b200/csrc/ops.cuh"

using namespace tupan;

#ifndef LAB_OP
#define LAB_OP 0
#endif
#if LAB_OP == 0
typedef AccJerkOp<double> LabOp;
enum { LAB_NIN = 8, LAB_NOUT = 6, LAB_FLOPS = 42 };
static void lab_params(NoParams&) {}
#define LAB_DEFAULT X(0, 8, 256, 0) X(3, 2, 256, 27) X(3, 2, 256, 11) X(2, 4, 256, 27) X(2, 4, 256, 11) X(2, 4, 256, 3)
#elif LAB_OP == 1
typedef AccOp<double> LabOp;
enum { LAB_NIN = 5, LAB_NOUT = 3, LAB_FLOPS = 20 };
static void lab_params(NoParams&) {}
#define LAB_DEFAULT X(0, 4, 256, 0) X(4, 2, 256, 8) X(3, 2, 256, 8) X(2, 4, 256, 8) X(3, 4, 256, 8) X(4, 4, 256, 8) \
    X(6, 2, 256, 8) X(8, 1, 256, 8) X(6, 1, 256, 8) X(4, 2, 256, 0)
#elif LAB_OP == 2
typedef PhiOp<double> LabOp;
enum { LAB_NIN = 5, LAB_NOUT = 1, LAB_FLOPS = 14 };
static void lab_params(NoParams&) {}
#define LAB_DEFAULT X(0, 4, 256, 0) X(4, 2, 256, 8) X(3, 2, 256, 8) X(6, 2, 256, 8) X(8, 2, 256, 8) X(4, 4, 256, 8) \
    X(8, 1, 256, 8) X(6, 2, 256, 0) X(3, 4, 256, 8)
#elif LAB_OP == 3
typedef TstepOp<double> LabOp;
enum { LAB_NIN = 8, LAB_NOUT = 2, LAB_FLOPS = 42 };
static void lab_params(TstepParams<double>& p) { p.eta = 1.0 / 64; p.eta_k1 = p.eta * 0.5; p.eta_k2 = p.eta * 0.375; }
#define LAB_DEFAULT X(0, 4, 256, 0) X(3, 2, 256, 8) X(2, 2, 256, 8) X(2, 4, 256, 8) X(4, 2, 256, 8) X(4, 1, 256, 8) X(3, 1, 256, 8)
#else
typedef NregXOp<double> LabOp;
enum { LAB_NIN = 8, LAB_NOUT = 7, LAB_FLOPS = 37 };
static void lab_params(DtParams<double>& p) { p.dt = 1.0 / 64; }
#define LAB_DEFAULT X(0, 4, 256, 0) X(3, 2, 256, 8) X(2, 2, 256, 8) X(2, 4, 256, 8) X(4, 2, 256, 8) X(4, 1, 256, 8) X(3, 1, 256, 8)
#endif

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// the production Op with another group shape; W = 0: the round-1 kernel (not grouped)
template <int W, int U, int NT_, int MODE_> struct AJ : LabOp {
    enum { GROUPED = W > 0, GW = W > 0 ? W : 1, GU = U, GNT = NT_, GMODE = MODE_ };
};

template <class Op>
static double run_variant(const char* name, const InRefs<double>& in, long long n_alloc, const double* jpack,
                          long long nj, double* out[LAB_NOUT], int sms, bool self)
{
    typedef Tune<Op> U;
    auto k = KernelOf<Op, false>::template get<U::NT, U::TJ, U::STAGES, false>();
    const size_t smem = U::SMEM;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, U::NT, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k));
    if (occ < 1) { printf("%-34s does not fit (regs=%d smem=%zu)\n", name, fa.numRegs, smem); return 0; }
    const long long IB = (long long)U::NT * U::WPT;
    long long waves = 2;
    long long ni = waves * occ * sms * IB;
    while (ni > n_alloc && waves > 1) { waves--; ni = waves * occ * sms * IB; }
    if (ni > n_alloc) { printf("%-34s skipped\n", name); return 0; }
    PairArgs<Op> a;
    // self = false: the i particles are taken from beyond the j range, no pair has r2 == 0
    const long long ioff = self ? 0 : nj;
    for (int q = 0; q < MAX_IN; ++q) a.i.p[q] = in.p[q] ? in.p[q] + ioff : nullptr;
    a.ni = ni; a.jpack = jpack; a.seg.nseg = 0; a.j0 = 0; a.j1 = nj; a.jchunk = nj; a.js_log2 = 0; a.slot0 = 0;
    a.partial = nullptr; a.one = 1; lab_params(a.prm);
    for (int q = 0; q < MAX_OUT; ++q) a.out.p[q] = q < LAB_NOUT ? out[q] : nullptr;
    dim3 grid((unsigned)(ni / IB), 1);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k<<<grid, U::NT, smem>>>(a);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k<<<grid, U::NT, smem>>>(a);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    std::vector<double> h(1024);
    double cs = 0;
    for (int q = 0; q < LAB_NOUT; ++q) {
        CK(cudaMemcpy(h.data(), out[q], 1024 * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 1024; ++i) cs += fabs(h[i]);
    }
    const double gp = (double)ni * nj / (best * 1e-3) * 1e-9;
    printf("%-30s %s regs=%3d occ=%d smem=%6zu ni=%7lld %8.3f ms %7.1f Gpair/s %5.2f clk/pair  frac %.4f  cs %.12e\n",
           name, self ? "self" : "rect", fa.numRegs, occ, smem, ni, best, gp, 148.0 * 4 * 32 * 1.965e9 / (gp * 1e9),
           gp * (LAB_FLOPS * 1e9) / (148.0 * 128 * 1.965e9), cs);
    return gp;
}

#ifdef LAB_LIST_FILE
#include LAB_LIST_FILE
#endif
#ifndef LAB_LIST
#define LAB_LIST LAB_DEFAULT
#endif

int main(int argc, char** argv)
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const long long nj = argc > 1 ? atoll(argv[1]) : 32768;
    const long long n = nj + 2LL * 2 * sms * 1024 * 2;
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
    for (int k = 0; k < MAX_IN; ++k) in.p[k] = k < LAB_NIN ? d + k * n : nullptr;
    double* jpack;
    CK(cudaMalloc(&jpack, nj * 8 * sizeof(double)));
    pack_j_kernel<LabOp><<<296, 256>>>(in, nj, jpack);
    CK(cudaDeviceSynchronize());
    double* out[LAB_NOUT];
    for (int q = 0; q < LAB_NOUT; ++q) CK(cudaMalloc(&out[q], n * sizeof(double)));
    printf("%s, %d SMs, nj = %lld\n", p.name, sms, nj);
#define X(W, U_, NT_, D) run_variant<AJ<W, U_, NT_, D>>("W" #W " U" #U_ " NT" #NT_ " mode" #D, in, n - nj, jpack, nj, out, sms, true);
    LAB_LIST
#undef X
#define X(W, U_, NT_, D) run_variant<AJ<W, U_, NT_, D>>("W" #W " U" #U_ " NT" #NT_ " mode" #D, in, n - nj, jpack, nj, out, sms, false);
    LAB_LIST
#undef X
    return 0;
}
