// microbench_fp32.cu -- FP32 pipe facts for the fp32 pair kernels on B200 (sm_100a): throughput of
// FFMA vs the packed FFMA2 (PTX fma.rn.f32x2, new with sm_100) alone and with the other
// instructions of a pair body (integer/select, MUFU, LDS) issued beside them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench_fp32 tools/microbench_fp32.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float lo_of(u64 a) { float l, h; asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a)); return l; }
__device__ __forceinline__ float hi_of(u64 a) { float l, h; asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a)); return h; }

__device__ long long g_clk[2];
__device__ unsigned long long g_ns[2];
__device__ __forceinline__ unsigned long long globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// MODE 0: FFMA x8 chains.  1: FFMA2 x8 chains.  2: FFMA + 1 int op each.  3: FFMA2 + 1 int op each.
// 4: FFMA2 + 2 int ops each.  5: FFMA2 + MUFU.RSQ per 8.  6: FFMA2 with a broadcast scalar operand.
// 7: FFMA2 + LDS.128 per 16.  8: FFMA2 + 1 FFMA each.
template <int MODE>
__global__ void __launch_bounds__(256) pipe_kernel(float* out, int iters, float a, float b)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { g_clk[0] = clock64(); g_ns[0] = globaltimer(); }
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(a, b, a, b);
    __syncthreads();
    float x[8];
    u64 p[8];
    unsigned u[8];
    const u64 a2 = pk(a, a * 0.9999f), b2 = pk(b, b * 1.0001f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x + k; p[k] = pk(threadIdx.x + k, threadIdx.x - k); u[k] = threadIdx.x * 7 + k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0 || MODE == 2) x[k] = fmaf(x[k], a, b);
                else if (MODE == 6) p[k] = fma2(p[k], pk(a, a), b2);
                else p[k] = fma2(p[k], a2, b2);
                if (MODE == 2 || MODE == 3 || MODE == 4) u[k] = u[k] * 3u + (unsigned)i;
                if (MODE == 4) u[k] = (u[k] >> 3) ^ (unsigned)r;
                if (MODE == 8) x[k] = fmaf(x[k], a, b);
            }
            if (MODE == 5) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(lo_of(p[r]))); p[r] = pk(y, hi_of(p[r])); }
            if (MODE == 7 && (r & 1) == 0) { float4 v = sm[(i + r) & 63]; p[r] = fma2(p[r], pk(v.x, v.y), pk(v.z, v.w)); }
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k] + lo_of(p[k]) + hi_of(p[k]) + (float)u[k];
    if (s == 123.456f) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) { g_clk[1] = clock64(); g_ns[1] = globaltimer(); }
}

template <int MODE> static void run_pipe(const char* name, int sms, float* d, int lanes_per_op)
{
    const int grid = sms * 4, block = 256, iters = 40000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    pipe_kernel<MODE><<<grid, block>>>(d, iters, 0.999999f, 1e-6f);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    pipe_kernel<MODE><<<grid, block>>>(d, iters, 0.999999f, 1e-6f);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double instr = (double)grid * block * iters * 64.0;      // main-chain instructions (per thread)
    long long clk[2];
    unsigned long long ns[2];
    CK(cudaMemcpyFromSymbol(clk, g_clk, sizeof(clk)));
    CK(cudaMemcpyFromSymbol(ns, g_ns, sizeof(ns)));
    const double mhz = (double)(clk[1] - clk[0]) / (double)(ns[1] - ns[0]) * 1e3;
    const double per_clk = instr / (ms * 1e-3) / (mhz * 1e6) / sms;
    printf("%-40s %8.3f ms  %6.1f thread-instr/clk/SM  = %6.1f fp32 FMA/clk/SM  %6.2f TFLOP/s  SM clock %.0f MHz\n", name,
           ms, per_clk, per_clk * lanes_per_op, 2.0 * instr * lanes_per_op / ms * 1e-9, mhz);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("%s  SMs=%d\n", p.name, p.multiProcessorCount);
    float* d;
    CK(cudaMalloc(&d, 1024));
    const int sms = p.multiProcessorCount;
    run_pipe<0>("FFMA x8 chains", sms, d, 1);
    run_pipe<1>("FFMA2 x8 chains", sms, d, 2);
    run_pipe<6>("FFMA2, one broadcast scalar operand", sms, d, 2);
    run_pipe<2>("FFMA + 1 int op each", sms, d, 1);
    run_pipe<3>("FFMA2 + 1 int op each", sms, d, 2);
    run_pipe<4>("FFMA2 + 2 int ops each", sms, d, 2);
    run_pipe<5>("FFMA2 + MUFU.RSQ per 8", sms, d, 2);
    run_pipe<7>("FFMA2 + LDS.128 per 16", sms, d, 2);
    run_pipe<8>("FFMA2 + 1 FFMA each (counts FFMA2 only)", sms, d, 2);
    return 0;
}
