// kernel_lab7.cu -- round 2: shapes of the grouped nreg_X kernel (NregXOp::group_phase1/2 on pair_kernel_grouped)
// next to the ungrouped one (W = 0), timed at whole waves; kernel_lab7.cu adapted to NregXOp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DTUPAN_FP64 \
//        -o tools/bin/kernel_lab7 tools/kernel_lab7.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../tupan_b200/csrc/ops.cuh"

using namespace tupan;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// the production Op with another group shape; W = 0: the round-1 kernel (not grouped)
template <int W, int U, int NT_, int MODE_> struct AJ : NregXOp<double> {
    enum { GROUPED = W > 0, GW = W > 0 ? W : 1, GU = U, GNT = NT_, GMODE = MODE_ };
};

template <class Op>
static double run_variant(const char* name, const InRefs<double>& in, long long n_alloc, const double* jpack,
                          long long nj, double* out[7], int sms, bool self)
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
    a.partial = nullptr; a.one = 1; a.prm.dt = 1.0 / 64;
    for (int q = 0; q < MAX_OUT; ++q) a.out.p[q] = q < 7 ? out[q] : nullptr;
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
    for (int q = 0; q < 7; ++q) {
        CK(cudaMemcpy(h.data(), out[q], 1024 * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 1024; ++i) cs += fabs(h[i]);
    }
    const double gp = (double)ni * nj / (best * 1e-3) * 1e-9;
    printf("%-30s %s regs=%3d occ=%d smem=%6zu ni=%7lld %8.3f ms %7.1f Gpair/s %5.2f clk/pair  frac %.4f  cs %.12e\n",
           name, self ? "self" : "rect", fa.numRegs, occ, smem, ni, best, gp, 148.0 * 4 * 32 * 1.965e9 / (gp * 1e9),
           gp * 42e9 / (148.0 * 128 * 1.965e9), cs);
    return gp;
}

#ifdef LAB_LIST_FILE
#include LAB_LIST_FILE
#endif
#ifndef LAB_LIST
#define LAB_LIST \
    X(0, 4, 256, 0) X(3, 2, 256, 8) X(2, 2, 256, 8) X(2, 4, 256, 8) X(4, 2, 256, 8) X(4, 1, 256, 8) X(3, 1, 256, 8)
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
    for (int k = 0; k < MAX_IN; ++k) in.p[k] = k < 8 ? d + k * n : nullptr;
    double* jpack;
    CK(cudaMalloc(&jpack, nj * 8 * sizeof(double)));
    pack_j_kernel<NregXOp<double>><<<296, 256>>>(in, nj, jpack);
    CK(cudaDeviceSynchronize());
    double* out[7];
    for (int q = 0; q < 7; ++q) CK(cudaMalloc(&out[q], n * sizeof(double)));
    printf("%s, %d SMs, nj = %lld\n", p.name, sms, nj);
#define X(W, U_, NT_, D) run_variant<AJ<W, U_, NT_, D>>("W" #W " U" #U_ " NT" #NT_ " mode" #D, in, n - nj, jpack, nj, out, sms, true);
    LAB_LIST
#undef X
#define X(W, U_, NT_, D) run_variant<AJ<W, U_, NT_, D>>("W" #W " U" #U_ " NT" #NT_ " mode" #D, in, n - nj, jpack, nj, out, sms, false);
    LAB_LIST
#undef X
    return 0;
}
