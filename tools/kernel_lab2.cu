// kernel_lab2.cu -- round 2: instruction ORDER variants of the acc_jerk fp64 pair body.
//
// Finding that drives this file (tools/microbench2.cu, tools/sass_rf.py): an FP64 instruction
// holds the pipe for 2 clocks, but a DFMA whose three sources are three distinct registers that
// the operand-reuse cache does not serve needs a third clock.  The round-1 kernel has 8.5 such
// DFMAs per pair (of 16): 64 + 8.5 = 72.5 clocks modelled, 74.2 measured.  The variants below
// compute the same 32 FP64 operations per pair in orders that let consecutive DFMAs share an
// operand in the same slot.  `--sass` builds need no GPU: compile with -cubin and feed each
// kernel to tools/sass_rf.py.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DTUPAN_FP64 \
//        -o tools/bin/kernel_lab2 tools/kernel_lab2.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../tupan_b200/csrc/ops.cuh"

using namespace tupan;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

#ifndef LAB_VARIANTS
#define LAB_VARIANTS 1
#endif

typedef double T;
enum { IX, IY, IZ, IE, IVX, IVY, IVZ, NI = 7, NA = 6, NJP = 8 };

__device__ __forceinline__ T seed_masked(T x, T r2)
{
    return rsqrt_seed_masked<false>(x, r2);
}
// seed without a mask; the low word is whatever the register held (perturbs the 20-bit seed by
// < 2^-20, absorbed by the cubic step)
__device__ __forceinline__ T seed_raw(T x)
{
    T y0;
    asm("{\n"
        ".reg .b32 xl, xh, yh, junk;\n"
        ".reg .f64 y;\n"
        "rsqrt.approx.ftz.f64 y, %1;\n"
        "mov.b64 {xl, yh}, y;\n"
        "mov.b64 %0, {junk, yh};\n"
        "}\n"
        : "=d"(y0)
        : "d"(x));
    return y0;
}

// BODY 0: round-1 order (AccJerkOp::pair per particle)
// BODY 1: operand-sharing order, particles one after the other
// BODY 2: as 1, r2/rv chains interleaved pairwise (ry, ry | ry, vy)
// BODY 3: as 1 with the unmasked seed + running minimum of r2's high word (mask deferred to a
//         slow path that the caller takes when the minimum says a pair needs it)
template <int BODY>
__device__ __forceinline__ void pair1(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA], unsigned& hmin)
{
    if (BODY == 0) {
        NoParams prm;
        AccJerkOp<T>::pair(s, row, a, prm);
        return;
    }
    T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
    T vx = s[IVX] - row[J8_VX], vy = s[IVY] - row[J8_VY], vz = s[IVZ] - row[J8_VZ];
    T e = s[IE] + row[J8_E2];
    T r2, rv;
    if (BODY == 2) {
        r2 = rx * rx; rv = rx * vx;
        r2 = fma(ry, ry, r2); rv = fma(ry, vy, rv);
        r2 = fma(rz, rz, r2); rv = fma(rz, vz, rv);
    } else {
        r2 = rx * rx; r2 = fma(ry, ry, r2); r2 = fma(rz, rz, r2);
        rv = rx * vx; rv = fma(ry, vy, rv); rv = fma(rz, vz, rv);
    }
    T x = r2 + e;
    T y0;
    if (BODY == 3) {
        y0 = seed_raw(x);
        hmin = min(hmin, (unsigned)__double2hiint(r2));
    } else {
        y0 = seed_masked(x, r2);
    }
    T t = x * y0;
    T h = fma(-t, y0, 1.0);
    T p = fma(h, 0.64951905283832900, 0.86602540378443865);
    T c = fma(h, p, 1.7320508075688772);
    T r1 = y0 * c;
    T q2 = r1 * r1;
    T q3 = q2 * r1;
    T nalpha = -(q2 * rv);
    T g = -(row[JM] * q3);
    // alpha group, then g group entered through the operand the two groups share (rz)
    vx = fma(nalpha, rx, vx); vy = fma(nalpha, ry, vy); vz = fma(nalpha, rz, vz);
    a[2] = fma(g, rz, a[2]); a[1] = fma(g, ry, a[1]); a[0] = fma(g, rx, a[0]);
    a[3] = fma(g, vx, a[3]); a[4] = fma(g, vy, a[4]); a[5] = fma(g, vz, a[5]);
}

// BODY >= 10: two-phase body over groups of R rows: phase 1 computes (r, v, -alpha, g) of the R*W pairs of
// the group, phase 2 does the 9 accumulate DFMAs of every pair in an operand-sharing order.  The phases are
// kept apart by SEP (10: nothing, 11: bar.warp.sync, 12: an opaque one-trip loop = separate basic blocks).
struct PairVals { T rx, ry, rz, vx, vy, vz, na, g, mj; };
template <int BODY>
__device__ __forceinline__ void phase1(const T (&s)[NI], const T (&row)[NJP], PairVals& o)
{
    T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
    T vx = s[IVX] - row[J8_VX], vy = s[IVY] - row[J8_VY], vz = s[IVZ] - row[J8_VZ];
    T e = s[IE] + row[J8_E2];
    T r2 = rx * rx; r2 = fma(ry, ry, r2); r2 = fma(rz, rz, r2);
    T rv = rx * vx; rv = fma(ry, vy, rv); rv = fma(rz, vz, rv);
    T x = r2 + e;
    T y0 = seed_masked(x, r2);
    T t = x * y0;
    T h = fma(-t, y0, 1.0);
    T p = fma(h, 0.64951905283832900, 0.86602540378443865);
    T c = fma(h, p, 1.7320508075688772);
    T r1 = y0 * c;
    T q2 = r1 * r1;
    T q3 = q2 * r1;
    o.na = -(q2 * rv);
    o.g = -(row[JM] * q3);
    o.rx = rx; o.ry = ry; o.rz = rz; o.vx = vx; o.vy = vy; o.vz = vz;
}
__device__ __forceinline__ void phase2(PairVals& o, T (&a)[NA])
{
    o.vx = fma(o.na, o.rx, o.vx); o.vy = fma(o.na, o.ry, o.vy); o.vz = fma(o.na, o.rz, o.vz);
    a[2] = fma(o.g, o.rz, a[2]); a[1] = fma(o.g, o.ry, a[1]); a[0] = fma(o.g, o.rx, a[0]);
    a[3] = fma(o.g, o.vx, a[3]); a[4] = fma(o.g, o.vy, a[4]); a[5] = fma(o.g, o.vz, a[5]);
}

// BODY 13/14/15: the G = U*W pairs of a row group written operation by operation (SoA over the
// pairs), so that the source order IS the intended issue order: dependent operations are G
// instructions apart, and consecutive DFMAs share an operand in the same slot.
// BODY 17: phase 1 cut once more, after the r2 / r.v chains (three basic blocks per group)
template <int G>
__device__ __forceinline__ void group_phase1a(const T (*s)[NI], const T (*row)[NJP], PairVals (&o)[G], T (&r2)[G],
                                              T (&rv)[G], T (&e)[G])
{
#pragma unroll
    for (int p = 0; p < G; ++p) {
        const T(&si)[NI] = s[p % 2 == 0 ? 0 : 1];
        (void)si;
    }
}
template <int G, int BODY>
__device__ __forceinline__ void group_phase1(const T (*s)[NI], const T (*row)[NJP], PairVals (&o)[G], const int (&pi)[G],
                                             const int (&pr)[G], unsigned& hmin, int one = 1)
{
    T r2[G], rv[G], e[G], x[G], y0[G];
#pragma unroll
    for (int p = 0; p < G; ++p) {
        const T(&si)[NI] = s[pi[p]];
        const T(&rw)[NJP] = row[pr[p]];
        o[p].rx = si[IX] - rw[JX]; o[p].vx = si[IVX] - rw[J8_VX];
        o[p].ry = si[IY] - rw[JY]; o[p].vy = si[IVY] - rw[J8_VY];
        o[p].rz = si[IZ] - rw[JZ]; o[p].vz = si[IVZ] - rw[J8_VZ];
        e[p] = si[IE] + rw[J8_E2];
    }
    if (BODY == 17 || BODY == 18 || BODY == 19) {
#pragma unroll 1
        for (int z = 0; z < one; ++z) {
#pragma unroll
            for (int p = 0; p < G; ++p) { r2[p] = o[p].rx * o[p].rx; rv[p] = o[p].rx * o[p].vx; }
#pragma unroll
            for (int p = 0; p < G; ++p) { r2[p] = fma(o[p].ry, o[p].ry, r2[p]); rv[p] = fma(o[p].ry, o[p].vy, rv[p]); }
#pragma unroll
            for (int p = 0; p < G; ++p) { r2[p] = fma(o[p].rz, o[p].rz, r2[p]); rv[p] = fma(o[p].rz, o[p].vz, rv[p]); }
        }
    } else {
#pragma unroll
    for (int p = 0; p < G; ++p) { r2[p] = o[p].rx * o[p].rx; rv[p] = o[p].rx * o[p].vx; }
#pragma unroll
    for (int p = 0; p < G; ++p) { r2[p] = fma(o[p].ry, o[p].ry, r2[p]); rv[p] = fma(o[p].ry, o[p].vy, rv[p]); }
#pragma unroll
    for (int p = 0; p < G; ++p) { r2[p] = fma(o[p].rz, o[p].rz, r2[p]); rv[p] = fma(o[p].rz, o[p].vz, rv[p]); }
    }
#pragma unroll
    for (int p = 0; p < G; ++p) x[p] = r2[p] + e[p];
#pragma unroll
    for (int p = 0; p < G; ++p) {
        if (BODY == 15) {
            y0[p] = seed_raw(x[p]);
            hmin = min(hmin, (unsigned)__double2hiint(r2[p]));
        } else {
            y0[p] = seed_masked(x[p], r2[p]);
        }
    }
    T t[G], h[G];
#pragma unroll
    for (int p = 0; p < G; ++p) t[p] = x[p] * y0[p];
#pragma unroll
    for (int p = 0; p < G; ++p) h[p] = fma(-t[p], y0[p], 1.0);
#pragma unroll
    for (int p = 0; p < G; ++p) t[p] = fma(h[p], 0.64951905283832900, 0.86602540378443865);
#pragma unroll
    for (int p = 0; p < G; ++p) t[p] = fma(h[p], t[p], 1.7320508075688772);
#pragma unroll
    for (int p = 0; p < G; ++p) t[p] = y0[p] * t[p];            // r1
#pragma unroll
    for (int p = 0; p < G; ++p) h[p] = t[p] * t[p];             // q2
#pragma unroll
    for (int p = 0; p < G; ++p) { o[p].na = -(h[p] * rv[p]); t[p] = h[p] * t[p]; }   // -alpha, q3
#pragma unroll
    for (int p = 0; p < G; ++p) {
        if (BODY == 16 || BODY == 18 || BODY == 19) { o[p].g = t[p]; o[p].mj = row[pr[p]][JM]; }
        else o[p].g = -(row[pr[p]][JM] * t[p]);
    }
}
__device__ __forceinline__ void chain_phase2_late_g(PairVals& o, T (&a)[NA])
{
    o.vx = fma(o.na, o.rx, o.vx);
    o.vy = fma(o.na, o.ry, o.vy);
    o.vz = fma(o.na, o.rz, o.vz);
    T g;
    asm volatile("mul.f64 %0, %1, %2;" : "=d"(g) : "d"(-o.mj), "d"(o.g));
    a[2] = fma(o.rz, g, a[2]);
    a[1] = fma(o.ry, g, a[1]);
    a[0] = fma(o.rx, g, a[0]);
    a[3] = fma(o.vx, g, a[3]);
    a[4] = fma(o.vy, g, a[4]);
    a[5] = fma(o.vz, g, a[5]);
}
__device__ __forceinline__ void chain_phase2(PairVals& o, T (&a)[NA])
{
    o.vx = fma(o.na, o.rx, o.vx);
    a[0] = fma(o.g, o.rx, a[0]);
    a[1] = fma(o.g, o.ry, a[1]);
    o.vy = fma(o.na, o.ry, o.vy);
    o.vz = fma(o.na, o.rz, o.vz);
    a[2] = fma(o.g, o.rz, a[2]);
    a[3] = fma(o.g, o.vx, a[3]);
    a[4] = fma(o.g, o.vy, a[4]);
    a[5] = fma(o.g, o.vz, a[5]);
}

template <int W, int U, int BODY, int NT_>
__global__ void __launch_bounds__(NT_) lab_kernel(InRefs<T> in, long long ni, const T* __restrict__ jpack, long long nj,
                                                  OutRefs<T> out, int one = 1)
{
    constexpr int TJ = 128, STAGES = 4, TILE_ELEMS = TJ * NJP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tiles = reinterpret_cast<T*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * TILE_ELEMS * sizeof(T));
    const int tid = threadIdx.x;
    const long long ibase = (long long)blockIdx.x * (NT_ * W);
    T is[W][NI], acc[W][NA];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        long long i = ibase + (long long)w * NT_ + tid;
        if (i > ni - 1) i = ni - 1;
        AccJerkOp<T>::load_i(in.p, i, is[w]);
#pragma unroll
        for (int k = 0; k < NA; ++k) acc[w][k] = 0;
    }
    const int ntiles = (int)(nj / TJ);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int s = t % STAGES;
        mbar_expect_tx(&full[s], TILE_ELEMS * 8);
        bulk_g2s(tiles + s * TILE_ELEMS, jpack + (long long)t * TILE_ELEMS, TILE_ELEMS * 8, &full[s]);
    };
    if (tid == 0)
        for (int t = 0; t < STAGES && t < ntiles; ++t) issue(t);
    unsigned hmin = 0xffffffffu;
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(&full[s], (unsigned)(t / STAGES) & 1u);
        const T* sj = tiles + s * TILE_ELEMS;
        if (BODY >= 13) {
            constexpr int G = U * W;
#pragma unroll 1
            for (int j = 0; j < TJ; j += U) {
                T rows[U][NJP];
                PairVals pv[G];
                int pi[G], pr[G];
#pragma unroll
                for (int u = 0; u < U; ++u) load_row<AccJerkOp<T>>(sj + (j + u) * NJP, rows[u]);
#pragma unroll
                for (int p = 0; p < G; ++p) { pi[p] = p % W; pr[p] = p / W; }
                group_phase1<G, BODY>(is, rows, pv, pi, pr, hmin, one);
                if (BODY == 13) {
#pragma unroll
                    for (int p = 0; p < G; ++p) chain_phase2(pv[p], acc[p % W]);
                } else {
#pragma unroll 1
                    for (int z = 0; z < one; ++z) {
#pragma unroll
                        if (BODY == 19) {
#pragma unroll
                            for (int p = 0; p < G; ++p) {
                                pv[p].vx = fma(pv[p].na, pv[p].rx, pv[p].vx);
                                pv[p].vy = fma(pv[p].na, pv[p].ry, pv[p].vy);
                                pv[p].vz = fma(pv[p].na, pv[p].rz, pv[p].vz);
                            }
#pragma unroll
                            for (int p = 0; p < G; ++p) {
                                T g;
                                asm volatile("mul.f64 %0, %1, %2;" : "=d"(g) : "d"(-pv[p].mj), "d"(pv[p].g));
                                T(&a)[NA] = acc[p % W];
                                a[0] = fma(pv[p].rx, g, a[0]); a[1] = fma(pv[p].ry, g, a[1]); a[2] = fma(pv[p].rz, g, a[2]);
                                a[3] = fma(pv[p].vx, g, a[3]); a[4] = fma(pv[p].vy, g, a[4]); a[5] = fma(pv[p].vz, g, a[5]);
                            }
                        } else
                        for (int p = 0; p < G; ++p) {
                            if (BODY == 16 || BODY == 18) chain_phase2_late_g(pv[p], acc[p % W]);
                            else chain_phase2(pv[p], acc[p % W]);
                        }
                    }
                }
            }
        } else if (BODY >= 10) {
            // U = rows per group here
#pragma unroll 1
            for (int j = 0; j < TJ; j += U) {
                PairVals pv[U][W];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    T row[NJP];
                    load_row<AccJerkOp<T>>(sj + (j + u) * NJP, row);
#pragma unroll
                    for (int w = 0; w < W; ++w) phase1<BODY>(is[w], row, pv[u][w]);
                }
                if (BODY == 11) __syncwarp();
                if (BODY == 12) {
#pragma unroll 1
                    for (int z = 0; z < one; ++z) {
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int w = 0; w < W; ++w) phase2(pv[u][w], acc[w]);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int w = 0; w < W; ++w) phase2(pv[u][w], acc[w]);
                }
            }
        } else {
#pragma unroll U
        for (int j = 0; j < TJ; ++j) {
            T row[NJP];
            load_row<AccJerkOp<T>>(sj + j * NJP, row);
#pragma unroll
            for (int w = 0; w < W; ++w) pair1<BODY>(is[w], row, acc[w], hmin);
        }
        }
        if (BODY == 3 && hmin < 0x00100000u) {
            // a pair of this tile needs the mask: not timed here (never taken with the lab's inputs);
            // the production kernel redoes the tile with the masked body
            acc[0][0] = nan("");
        }
        __syncthreads();
        if (tid == 0 && t + STAGES < ntiles) issue(t + STAGES);
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const long long i = ibase + (long long)w * NT_ + tid;
        if (i < ni) {
#pragma unroll
            for (int k = 0; k < NA; ++k) out.p[k][i] = acc[w][k];
        }
    }
}

template <int W, int U, int BODY, int NT_>
static double run_variant(const char* name, const InRefs<T>& in, long long n_alloc, const T* jpack, long long nj,
                          T* out[6], int sms)
{
    auto k = lab_kernel<W, U, BODY, NT_>;
    const size_t smem = 4 * 128 * NJP * sizeof(T) + 4 * 8;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NT_, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k));
    const long long IB = (long long)NT_ * W;
    long long waves = 2;
    long long ni = waves * occ * sms * IB;
    while (ni > n_alloc && waves > 1) { waves--; ni = waves * occ * sms * IB; }
    if (ni > n_alloc) { printf("%-34s skipped\n", name); return 0; }
    OutRefs<T> o;
    for (int q = 0; q < MAX_OUT; ++q) o.p[q] = q < 6 ? out[q] : nullptr;
    dim3 grid((unsigned)(ni / IB));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k<<<grid, NT_, smem>>>(in, ni, jpack, nj, o, 1);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k<<<grid, NT_, smem>>>(in, ni, jpack, nj, o, 1);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    std::vector<T> h(1024);
    double cs = 0;
    for (int q = 0; q < 6; ++q) {
        CK(cudaMemcpy(h.data(), out[q], 1024 * sizeof(T), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 1024; ++i) cs += fabs(h[i]);
    }
    const double gp = (double)ni * nj / (best * 1e-3) * 1e-9;
    const double clk = 148.0 * 4 * 32 * 1.965e9 / (gp * 1e9);
    printf("%-34s regs=%3d occ=%d ni=%7lld  %8.3f ms  %7.1f Gpair/s  %5.2f TF  %5.2f clk/pair  checksum %.12e\n", name,
           fa.numRegs, occ, ni, best, gp, gp * 42e-3, clk, cs);
    return gp;
}

#ifdef LAB_LIST_FILE
#include LAB_LIST_FILE
#endif
#ifndef LAB_LIST
#define LAB_LIST \
    X(2, 8, 0, 256) X(2, 4, 0, 256) X(2, 8, 1, 256) X(2, 8, 3, 256) X(2, 4, 3, 256) X(1, 8, 1, 256) X(1, 8, 3, 256) \
    X(3, 2, 1, 256) X(3, 4, 1, 128) X(2, 4, 3, 128) X(4, 2, 3, 128) \
    X(2, 2, 10, 256) X(2, 2, 12, 256) X(2, 2, 13, 256) X(2, 2, 14, 256) X(2, 2, 15, 256) X(2, 2, 16, 256) \
    X(1, 4, 14, 256) X(1, 4, 15, 256) X(1, 4, 16, 256) X(2, 1, 14, 256) X(1, 2, 14, 256) X(3, 1, 14, 256) \
    X(3, 2, 14, 256) X(2, 3, 14, 256) X(1, 3, 14, 256) X(1, 6, 14, 256) X(2, 4, 14, 256) X(1, 4, 14, 128) \
    X(2, 2, 14, 128) X(2, 2, 15, 128) X(1, 8, 14, 256)
#endif

int main(int argc, char** argv)
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const long long n = 4LL * 2 * sms * 1024;
    const long long nj = argc > 1 ? atoll(argv[1]) : 32768;
    std::vector<T> h(8 * n);
    srand(1);
    for (long long i = 0; i < n; ++i) {
        h[0 * n + i] = 1.0 / n;
        for (int k = 1; k <= 3; ++k) h[k * n + i] = (double)rand() / RAND_MAX - 0.5;
        h[4 * n + i] = 1e-6;
        for (int k = 5; k <= 7; ++k) h[k * n + i] = (double)rand() / RAND_MAX - 0.5;
    }
    T* d;
    CK(cudaMalloc(&d, 8 * n * sizeof(T)));
    CK(cudaMemcpy(d, h.data(), 8 * n * sizeof(T), cudaMemcpyHostToDevice));
    InRefs<T> in;
    for (int k = 0; k < MAX_IN; ++k) in.p[k] = k < 8 ? d + k * n : nullptr;
    T* jpack;
    CK(cudaMalloc(&jpack, nj * 8 * sizeof(T)));
    pack_j_kernel<AccJerkOp<T>><<<296, 256>>>(in, nj, jpack);
    CK(cudaDeviceSynchronize());
    T* out[6];
    for (int q = 0; q < 6; ++q) CK(cudaMalloc(&out[q], n * sizeof(T)));
    printf("%s, %d SMs, nj = %lld\n", p.name, sms, nj);
#define X(W, U, B, NT_) run_variant<W, U, B, NT_>("W" #W " U" #U " body" #B " NT" #NT_, in, n, jpack, nj, out, sms);
    LAB_LIST
#undef X
    return 0;
}
