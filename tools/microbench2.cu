// microbench2.cu -- issue rules of the B200 FP64 pipe that decide how an acc_jerk pair body
// should be written (round 2).  Every probe is a loop of independent chains per thread whose
// SASS was checked with cuobjdump (operand kinds, .reuse flags); results are FP64 warp
// instructions per clock per SM sub-partition expressed as clocks per FP64 instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/microbench2 tools/microbench2.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ long long g_clk[2];
__device__ unsigned long long g_ns[2];
__device__ __forceinline__ unsigned long long globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double rsq64h(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

enum {
    M_DADD_RC,      // x = x + const                         (1 register read)
    M_DADD_RR,      // x_k = x_k + y_k                       (2 distinct registers, y fixed)
    M_DADD_XX,      // x_k = x_k + x_{k+3}                   (2 distinct changing registers)
    M_DMUL_RR,      // x_k = x_k * y_k
    M_DFMA_RCC,     // x = fma(x, a, b), a b launch constants (the round-1 probe)
    M_DFMA_RRI,     // x_k = fma(x_k, y_k, 1.0)              (2 registers + immediate)
    M_DFMA_RRR_FIX, // x_k = fma(x_k, a, b), a b in registers (reusable)
    M_DFMA_ACC,     // acc_k = fma(g, r_k, acc_k), k = 0..5, g changes every 6 (the accumulate pattern)
    M_DFMA_ACC_R,   // acc_k = fma(r_k, g, acc_k): g in the second slot
    M_DFMA_XXX,     // 3 distinct changing registers
    M_DFMA_SQ,      // x_k = fma(y_k, y_k, x_k)              (r2 pattern: one register twice)
    M_MIX,          // one acc_jerk-shaped pair: 17 DFMA : 8 DADD : 7 DMUL, operand shapes as in the kernel
    M_MIX_LDS,      // M_MIX + 2 LDS.128 per 32
    M_MIX_LDS1,     // M_MIX + 1 LDS.128 per 32
    M_MIX_MUFU,     // M_MIX + 1 MUFU.RSQ64H per 32
    M_MIX_INT2,     // M_MIX + ISETP + SEL per 32
    M_MIX_ALL,      // M_MIX + 2 LDS.128 + MUFU + ISETP + SEL per 32  (today's pair body)
    M_MIX_ALL1,     // M_MIX + 1 LDS.128 + MUFU + VIMNMX per 32      (candidate pair body)
    M_COUNT
};

template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, const double* in, int iters, double a, double b)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { g_clk[0] = clock64(); g_ns[0] = globaltimer(); }
    __shared__ double2 sm[256];
    sm[threadIdx.x] = make_double2(a + threadIdx.x, b);
    __syncthreads();
    double x[8], y[8];
    unsigned u = threadIdx.x, umin = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = in[threadIdx.x + 32 * k]; y[k] = in[threadIdx.x + 32 * k + 256]; }
    double ra = in[threadIdx.x + 600], rb = in[threadIdx.x + 700];
    for (int i = 0; i < iters; ++i) {
        if (MODE == M_MIX || MODE >= M_MIX_LDS) {
            // one "pair": 32 FP64 instructions with the operand shapes of the acc_jerk body
            // g = y0 (changes per pair), r_k = y[1..6]
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
                double g = x[6] * y[7];                                      // DMUL
                double t0 = y[0] - ra, t1 = y[1] - ra, t2 = y[2] - rb;       // 3 DADD
                double t3 = y[3] - rb, t4 = y[4] - ra, t5 = y[5] - rb;       // 3 DADD
                double e = y[6] + ra;                                        // DADD (e2 sum)
                double q = t0 * t0; q = fma(t1, t1, q); q = fma(t2, t2, q);  // DMUL + 2 DFMA (one register twice)
                double rv = t0 * t3; rv = fma(t1, t4, rv); rv = fma(t2, t5, rv);  // DMUL + 2 DFMA (3 registers)
                double xx = q + e;                                           // DADD (x = r2 + e2)
                double s = xx * g;                                           // DMUL
                double h = fma(-s, g, 1.0);                                  // DFMA imm
                double p = fma(h, 0.649, 0.866);                             // DFMA
                double c = fma(h, p, 1.732);                                 // DFMA
                double r1 = g * c;                                           // DMUL
                double q2 = r1 * r1;                                         // DMUL
                double al = q2 * rv;                                         // DMUL
                double gg = q2 * r1;                                         // DMUL (7th; stands for q3, mj*q3 folded)
                t3 = fma(-al, t0, t3); t4 = fma(-al, t1, t4); t5 = fma(-al, t2, t5);   // 3 DFMA
                x[0] = fma(gg, t0, x[0]); x[1] = fma(gg, t1, x[1]); x[2] = fma(gg, t2, x[2]);
                x[3] = fma(gg, t3, x[3]); x[4] = fma(gg, t4, x[4]); x[5] = fma(gg, t5, x[5]);   // 6 DFMA acc
                x[6] = fma(h, 1e-30, x[6]);                                  // keeps the g chain changing
                // the next pair's "row": renames only
                y[0] = t3; y[1] = t4; y[2] = t5; y[3] = t0; y[4] = t1; y[5] = t2; y[6] = al;
                if (MODE == M_MIX_LDS || MODE == M_MIX_ALL) {
                    double2 v0 = sm[(i + rep) & 255], v1 = sm[(i + rep + 7) & 255];
                    y[0] = v0.x; y[1] = v0.y; y[2] = v1.x; y[3] = v1.y;
                }
                if (MODE == M_MIX_LDS1 || MODE == M_MIX_ALL1) {
                    double2 v0 = sm[(i + rep) & 255];
                    y[0] = v0.x; y[1] = v0.y;
                }
                if (MODE == M_MIX_MUFU || MODE == M_MIX_ALL || MODE == M_MIX_ALL1) y[7] = rsq64h(xx);
                if (MODE == M_MIX_INT2 || MODE == M_MIX_ALL) {
                    unsigned hi = (unsigned)__double2hiint(q);
                    unsigned yh = (unsigned)__double2hiint(y[7]);
                    yh = hi >= 0x00100000u ? yh : 0u;
                    y[7] = __hiloint2double((int)yh, __double2loint(y[7]));
                }
                if (MODE == M_MIX_ALL1) {
                    unsigned hi = (unsigned)__double2hiint(q);
                    umin = min(umin, hi);
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (MODE == M_DFMA_ACC || MODE == M_DFMA_ACC_R) {
                    // 6 accumulators, g = y[6], y[7] alternating (so g changes every 6 DFMAs)
                    const double g = (r & 1) ? y[7] : y[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k)
                        x[k] = MODE == M_DFMA_ACC ? fma(g, y[k], x[k]) : fma(y[k], g, x[k]);
                    // keep 8 per group so every mode counts 8 FP64 instructions per r
                    x[6] = fma(g, y[0], x[6]);
                    x[7] = fma(g, y[1], x[7]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (MODE == M_DADD_RC) x[k] = x[k] + a;
                        if (MODE == M_DADD_RR) x[k] = x[k] + y[k];
                        if (MODE == M_DADD_XX) x[k] = x[k] + x[(k + 3) & 7];
                        if (MODE == M_DMUL_RR) x[k] = x[k] * y[k];
                        if (MODE == M_DFMA_RCC) x[k] = fma(x[k], a, b);
                        if (MODE == M_DFMA_RRI) x[k] = fma(x[k], y[k], 1.0);
                        if (MODE == M_DFMA_RRR_FIX) x[k] = fma(x[k], ra, rb);
                        if (MODE == M_DFMA_XXX) x[k] = fma(x[k], x[(k + 3) & 7], x[(k + 5) & 7]);
                        if (MODE == M_DFMA_SQ) x[k] = fma(y[k], y[k], x[k]);
                    }
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k] + y[k];
    if (s == 123.456 || umin == 12345u || u == 0xdeadbeefu) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) { g_clk[1] = clock64(); g_ns[1] = globaltimer(); }
}

static const char* NAMES[M_COUNT] = {
    "DADD  x+const (1 reg)", "DADD  x_k+y_k (2 reg)", "DADD  x_k+x_k3 (2 changing)", "DMUL  x_k*y_k (2 reg)",
    "DFMA  fma(x,ca,cb) launch consts", "DFMA  fma(x_k,y_k,1.0)", "DFMA  fma(x_k,ra,rb) fixed regs",
    "DFMA  acc_k=fma(g,r_k,acc_k)", "DFMA  acc_k=fma(r_k,g,acc_k)", "DFMA  3 changing regs",
    "DFMA  fma(y_k,y_k,x_k)", "pair-shaped mix 19:7:6", "mix + 2 LDS.128", "mix + 1 LDS.128", "mix + MUFU.RSQ64H",
    "mix + ISETP + SEL", "mix + 2 LDS + MUFU + ISETP + SEL", "mix + 1 LDS + MUFU + VIMNMX"};

template <int MODE> static void run(int sms, double* d, const double* in, int ctas_per_sm)
{
    const int grid = sms * ctas_per_sm, block = 256;
    const bool mix = MODE == M_MIX || MODE >= M_MIX_LDS;
    const int iters = mix ? 6000 : 12000;
    const double per_iter = mix ? 66.0 : 32.0;       // mix: 2 pairs x (17 DFMA + 8 DADD + 8 DMUL)       // FP64 instructions per thread per iteration
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    probe<MODE><<<grid, block>>>(d, in, iters, 0.999999, 1e-6);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    probe<MODE><<<grid, block>>>(d, in, iters, 0.999999, 1e-6);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long clk[2];
    unsigned long long ns[2];
    CK(cudaMemcpyFromSymbol(clk, g_clk, sizeof(clk)));
    CK(cudaMemcpyFromSymbol(ns, g_ns, sizeof(ns)));
    const double mhz = (double)(clk[1] - clk[0]) / (double)(ns[1] - ns[0]) * 1e3;
    // warp instructions per sub-partition: warps per SM / 4
    const double winstr = (double)ctas_per_sm * (block / 32) / 4.0 * iters * per_iter;
    const double clocks = ms * 1e-3 * mhz * 1e6;
    printf("%-40s ctas/SM %d  %8.3f ms  %5.3f clk per FP64 instr  (%5.1f op/clk/SM)  SM %4.0f MHz\n", NAMES[MODE],
           ctas_per_sm, ms, clocks / winstr, 128.0 / (clocks / winstr), mhz);
}

template <int MODE> static void run_all(int sms, double* d, const double* in)
{
    run<MODE>(sms, d, in, 4);
    run<MODE>(sms, d, in, 2);
    run<MODE>(sms, d, in, 1);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("%s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    double *d, *in;
    CK(cudaMalloc(&d, 1024));
    CK(cudaMalloc(&in, 2048 * sizeof(double)));
    double h[2048];
    for (int i = 0; i < 2048; ++i) h[i] = 0.5 + 1e-3 * (i % 97);
    CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    const int sms = p.multiProcessorCount;
    run_all<M_DADD_RC>(sms, d, in);
    run_all<M_DADD_RR>(sms, d, in);
    run_all<M_DADD_XX>(sms, d, in);
    run_all<M_DMUL_RR>(sms, d, in);
    run_all<M_DFMA_RCC>(sms, d, in);
    run_all<M_DFMA_RRI>(sms, d, in);
    run_all<M_DFMA_RRR_FIX>(sms, d, in);
    run_all<M_DFMA_ACC>(sms, d, in);
    run_all<M_DFMA_ACC_R>(sms, d, in);
    run_all<M_DFMA_XXX>(sms, d, in);
    run_all<M_DFMA_SQ>(sms, d, in);
    run_all<M_MIX>(sms, d, in);
    run_all<M_MIX_LDS>(sms, d, in);
    run_all<M_MIX_LDS1>(sms, d, in);
    run_all<M_MIX_MUFU>(sms, d, in);
    run_all<M_MIX_INT2>(sms, d, in);
    run_all<M_MIX_ALL>(sms, d, in);
    run_all<M_MIX_ALL1>(sms, d, in);
    return 0;
}
