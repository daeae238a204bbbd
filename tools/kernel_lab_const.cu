// kernel_lab_const.cu -- experiment: j rows in the constant bank instead of shared memory.
// LDCU loads a row into UNIFORM registers once per warp and the FP64 instructions take the
// uniform register as an operand (DADD R, R, -UR), so the per-thread LDS.128 x4 per row and
// their 16 vector-register writes disappear.  Throughput probe only: the 32 KB bank (512 rows)
// is swept `passes` times with the accumulators in registers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DTUPAN_FP64 \
//        -o tools/bin/kernel_lab_const tools/kernel_lab_const.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../tupan_b200/csrc/ops.cuh"

using namespace tupan;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

enum { BANK_ROWS = 512 };
__constant__ double cbank[BANK_ROWS * 8];

template <int WPT, int UNROLL, int NT>
__global__ void __launch_bounds__(NT) const_kernel(InRefs<double> in, long long ni, int passes, OutRefs<double> out)
{
    typedef AccJerkOp<double> Op;
    double is[WPT][Op::NI], acc[WPT][Op::NA];
    const long long ibase = (long long)blockIdx.x * (NT * WPT);
#pragma unroll
    for (int w = 0; w < WPT; ++w) {
        long long i = ibase + (long long)w * NT + threadIdx.x;
        if (i > ni - 1) i = ni - 1;
        Op::load_i(in.p, i, is[w]);
        Op::zero(acc[w]);
    }
    NoParams prm;
    for (int p = 0; p < passes; ++p) {
#pragma unroll UNROLL
        for (int r = 0; r < BANK_ROWS; ++r) {
            double row[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) row[k] = cbank[r * 8 + k];
#pragma unroll
            for (int w = 0; w < WPT; ++w) Op::pair(is[w], row, acc[w], prm);
        }
    }
#pragma unroll
    for (int w = 0; w < WPT; ++w) {
        const long long i = ibase + (long long)w * NT + threadIdx.x;
        if (i < ni) Op::finish(in.p, i, acc[w], prm, out.p);
    }
}

template <int WPT, int UNROLL, int NT>
static void run(const char* name, const InRefs<double>& in, long long n_alloc, double* out[6], int sms, int passes)
{
    auto k = const_kernel<WPT, UNROLL, NT>;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NT, 0));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k));
    const long long IB = (long long)NT * WPT;
    long long waves = 2;
    long long ni = waves * occ * sms * IB;
    while (ni > n_alloc && waves > 1) { waves--; ni = waves * occ * sms * IB; }
    if (ni > n_alloc) { printf("%-30s skipped\n", name); return; }
    OutRefs<double> o;
    for (int q = 0; q < MAX_OUT; ++q) o.p[q] = q < 6 ? out[q] : nullptr;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k<<<(unsigned)(ni / IB), NT>>>(in, ni, passes, o);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k<<<(unsigned)(ni / IB), NT>>>(in, ni, passes, o);
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
    const double gp = (double)ni * BANK_ROWS * passes / (best * 1e-3) * 1e-9;
    printf("%-30s regs=%3d occ=%d ni=%7lld  %8.3f ms  %7.1f Gpair/s  %5.2f TF  checksum %.9e\n", name, fa.numRegs, occ,
           ni, best, gp, gp * 42e-3, cs);
}

int main(int argc, char** argv)
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const long long n = 4LL * 2 * sms * 1024;
    const int passes = argc > 1 ? atoi(argv[1]) : 64;
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
    CK(cudaMalloc(&jpack, BANK_ROWS * 8 * sizeof(double)));
    pack_j_kernel<AccJerkOp<double>><<<2, 256>>>(in, BANK_ROWS, jpack);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyToSymbol(cbank, jpack, BANK_ROWS * 8 * sizeof(double), 0, cudaMemcpyDeviceToDevice));
    double* out[6];
    for (int q = 0; q < 6; ++q) CK(cudaMalloc(&out[q], n * sizeof(double)));
    printf("%s, %d SMs, constant-bank j rows, %d passes over %d rows\n", p.name, sms, passes, (int)BANK_ROWS);
#define RUN(W, U, NT) run<W, U, NT>("const W" #W " U" #U " NT" #NT, in, n, out, sms, passes)
    RUN(2, 2, 256);
    RUN(2, 4, 256);
    RUN(2, 1, 256);
    RUN(1, 4, 256);
    RUN(1, 2, 256);
    RUN(3, 2, 256);
    RUN(4, 1, 256);
    RUN(4, 2, 128);
    RUN(2, 2, 128);
    RUN(1, 4, 128);
    RUN(2, 4, 512);
    return 0;
}
