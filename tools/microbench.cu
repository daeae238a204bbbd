// microbench.cu -- B200 FP64-pipe facts the pair kernels are designed around.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench tools/microbench.cu
// Prints: (1) accuracy of the MUFU.RSQ64H seed, (2) FP64 op throughput per SM per clock for
// DFMA / DADD / DMUL and mixes with integer, FP32, MUFU and F2F work in the shadow,
// (3) dependent-chain latency of DFMA.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double rsq64h(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

__global__ void seed_accuracy(double* maxrel, double lo, double hi, int n)
{
    double worst = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x = lo * pow(hi / lo, (double)i / n);
        double y0 = rsq64h(x);
        double exact = 1.0 / sqrt(x);
        double rel = fabs(y0 - exact) / exact;
        worst = fmax(worst, rel);
    }
    // block max
    for (int off = 16; off > 0; off >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, off));
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)maxrel, (unsigned long long)__double_as_longlong(worst));
}

// MODE: 0 DFMA, 1 DADD, 2 DMUL, 3 DFMA + 1 IMAD-ish int op per DFMA, 4 DFMA + 1 FFMA per DFMA,
//       5 DFMA + MUFU.RSQ64H every 8, 6 DFMA + F2F(f64->f32) every 8, 7 DFMA + 2 int per DFMA,
//       8 DFMA + LDS.128 every 4
__device__ long long g_clk[2];
__device__ unsigned long long g_ns[2];
__device__ __forceinline__ unsigned long long globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

//       9 DFMA with three distinct, changing register operands (no operand reuse)
//      10 DFMA / DADD / DMUL mix 19:7:6 as in the acc_jerk pair body
template <int MODE>
__global__ void __launch_bounds__(256) pipe_kernel(double* out, int iters, double a, double b)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { g_clk[0] = clock64(); g_ns[0] = globaltimer(); }
    __shared__ double2 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_double2(a, b);
    __syncthreads();
    double x[8];
    unsigned u[8];
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x + k; u[k] = threadIdx.x * 7 + k; f[k] = threadIdx.x + 0.5f * k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 1) x[k] = x[k] + a;
                else if (MODE == 2) x[k] = x[k] * a;
                else if (MODE == 9) x[k] = fma(x[k], x[(k + 3) & 7], x[(k + 5) & 7]);
                else if (MODE == 10) {
                    const int q = (r * 8 + k) & 31;
                    if (q < 19) x[k] = fma(x[k], a, b);
                    else if (q < 26) x[k] = x[k] + a;
                    else x[k] = x[k] * a;
                }
                else x[k] = fma(x[k], a, b);
                if (MODE == 3 || MODE == 7) u[k] = u[k] * 3u + (unsigned)i;
                if (MODE == 7) u[k] = (u[k] >> 3) ^ (unsigned)r;
                if (MODE == 4) f[k] = fmaf(f[k], 0.999f, 0.001f);
            }
            if (MODE == 5) x[r] = rsq64h(x[r]) + x[(r + 1) & 7];
            if (MODE == 6) f[r] += (float)x[r];
            if (MODE == 8 && (r & 3) == 0) { double2 v = sm[(i + r) & 63]; x[r] += v.x; x[(r + 1) & 7] += v.y; }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k] + (double)u[k] + (double)f[k];
    if (s == 123.456) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) { g_clk[1] = clock64(); g_ns[1] = globaltimer(); }
}

__global__ void latency_kernel(double* out, long long* cycles, int iters, double a, double b)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 64; ++k) x = fma(x, a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cycles = t1 - t0; out[0] = x; }
}

template <int MODE> static void run_pipe(const char* name, int sms, double* d)
{
    const int grid = sms * 4, block = 256, iters = 40000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    pipe_kernel<MODE><<<grid, block>>>(d, iters, 0.999999, 1e-6);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    pipe_kernel<MODE><<<grid, block>>>(d, iters, 0.999999, 1e-6);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double ops = (double)grid * block * iters * 64.0;   // FP64 ops of the main chain
    long long clk[2];
    unsigned long long ns[2];
    CK(cudaMemcpyFromSymbol(clk, g_clk, sizeof(clk)));
    CK(cudaMemcpyFromSymbol(ns, g_ns, sizeof(ns)));
    const double mhz = (double)(clk[1] - clk[0]) / (double)(ns[1] - ns[0]) * 1e3;
    const double per_clk = ops / (ms * 1e-3) / (mhz * 1e6) / 148.0;
    printf("%-44s %8.3f ms  %7.2f T FP64-op/s  (x2 = %6.2f TFLOP/s)  SM clock %.0f MHz  %.1f op/clk/SM\n", name, ms,
           ops / ms * 1e-9, 2 * ops / ms * 1e-9, mhz, per_clk);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("%s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    double* d;
    CK(cudaMalloc(&d, 1024));
    CK(cudaMemset(d, 0, 1024));
    struct { double lo, hi; } ranges[] = {{1.0, 4.0}, {1e-30, 1e30}, {1e-300, 1e300}};
    for (auto r : ranges) {
        CK(cudaMemset(d, 0, 8));
        seed_accuracy<<<296, 256>>>(d, r.lo, r.hi, 1 << 24);
        double w;
        CK(cudaMemcpy(&w, d, 8, cudaMemcpyDeviceToHost));
        printf("MUFU.RSQ64H max rel. error on [%g, %g]: %.3e = 2^%.2f\n", r.lo, r.hi, w, log2(w));
    }
    run_pipe<0>("DFMA x8 chains", p.multiProcessorCount, d);
    run_pipe<1>("DADD x8 chains", p.multiProcessorCount, d);
    run_pipe<2>("DMUL x8 chains", p.multiProcessorCount, d);
    run_pipe<3>("DFMA + 1 int op per DFMA", p.multiProcessorCount, d);
    run_pipe<7>("DFMA + 3 int ops per DFMA", p.multiProcessorCount, d);
    run_pipe<4>("DFMA + 1 FFMA per DFMA", p.multiProcessorCount, d);
    run_pipe<5>("DFMA + MUFU.RSQ64H (+DADD) per 8 DFMA", p.multiProcessorCount, d);
    run_pipe<6>("DFMA + F2F.F32.F64 (+FADD) per 8 DFMA", p.multiProcessorCount, d);
    run_pipe<8>("DFMA + LDS.128 (+2 DADD) per 32 DFMA", p.multiProcessorCount, d);
    run_pipe<9>("DFMA, 3 distinct changing operands", p.multiProcessorCount, d);
    run_pipe<10>("DFMA:DADD:DMUL = 19:7:6", p.multiProcessorCount, d);
    long long* cyc;
    CK(cudaMalloc(&cyc, 8));
    latency_kernel<<<1, 32>>>(d, cyc, 100, 0.999999, 1e-6);
    CK(cudaDeviceSynchronize());
    latency_kernel<<<1, 32>>>(d, cyc, 100, 0.999999, 1e-6);
    long long c;
    CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    printf("dependent DFMA chain, 1 warp: %.2f cycles per DFMA\n", (double)c / 6400.0);
    return 0;
}
