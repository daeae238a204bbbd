// peer.cu -- packed j rows in peer-visible memory (NVLink / NVSwitch), one process per GPU.
//
// The i-sharded evaluation needs every rank's packed rows on every GPU.  Instead of copying
// them around (all-gather) the pair kernel can read them where they are: its tiles are moved
// by TMA bulk copies from a global address, and a peer GPU's memory mapped through CUDA IPC
// is such an address.  At N = 2^20 on 8 GPUs a rank streams 7/8 x 64 MB x 256 i-blocks = 14 GB
// per evaluation in 0.27 s -- ~50 GB/s of the 900 GB/s an NVLink-5 port gives -- and the copy
// is overlapped with the arithmetic tile by tile by the same 4-stage ring that hides L2 latency.
//
// What this file provides (Part 2b of include/libtupan_cuda.h):
//   * allocation of a buffer other processes may map, export / import of its IPC handle;
//   * a device-side barrier between the ranks on the caller's stream: every rank stores its
//     epoch into a flag word in every peer's memory (st.release.sys) and spins on its own flag
//     words (ld.acquire.sys).  The epoch counter lives in device memory, so a captured CUDA
//     graph replays correctly.  The spin is bounded: a peer that does not arrive within the
//     timeout is counted in tupan_cuda_peer_timeouts() instead of hanging the GPU.
#include "runtime.cuh"
#include "../../include/libtupan_cuda.h"

namespace tupan {

enum { PEER_MAX = 16 };
struct PeerFlags { unsigned* p[PEER_MAX]; };   // p[r]: rank r's flag block: arrive[world], epoch

__device__ unsigned int peer_timeouts = 0;

__global__ void peer_barrier_kernel(PeerFlags f, int rank, int world, unsigned long long timeout_ns)
{
    unsigned* mine = f.p[rank];
    __shared__ unsigned epoch;
    if (threadIdx.x == 0) {
        epoch = mine[world] + 1u;
        mine[world] = epoch;
    }
    __syncthreads();
    const unsigned e = epoch;
    const int r = threadIdx.x;
    if (r >= world || r == rank) return;
    __threadfence_system();   // everything this GPU wrote before the barrier is visible to peers
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[r] + rank), "r"(e) : "memory");
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine + r) : "memory");
        if ((int)(v - e) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
            atomicAdd(&peer_timeouts, 1u);
            break;
        }
        __nanosleep(100);
    }
}

}  // namespace tupan

using namespace tupan;

extern "C" {

int tupan_cuda_peer_alloc(long long bytes, void** dptr, void* handle64)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    if (!dptr || !handle64 || bytes <= 0) return c.fail(cudaErrorInvalidValue, "peer_alloc arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    TUPAN_CHECK(cudaMalloc(&p, (size_t)bytes), "peer_alloc cudaMalloc");
    TUPAN_CHECK(cudaMemset(p, 0, (size_t)bytes), "peer_alloc cudaMemset");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return c.fail(e, "cudaIpcGetMemHandle");
    }
    memcpy(handle64, &h, 64);
    *dptr = p;
    return 0;
}

int tupan_cuda_peer_open(const void* handle64, void** dptr)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    if (!dptr || !handle64) return c.fail(cudaErrorInvalidValue, "peer_open arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    TUPAN_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    *dptr = p;
    return 0;
}

int tupan_cuda_peer_close(void* dptr)
{
    if (!dptr) return 0;
    TUPAN_CHECK(cudaIpcCloseMemHandle(dptr), "cudaIpcCloseMemHandle");
    return 0;
}

int tupan_cuda_peer_free(void* dptr)
{
    if (!dptr) return 0;
    TUPAN_CHECK(cudaFree(dptr), "peer_free");
    return 0;
}

int tupan_cuda_peer_barrier_dev(void* const* flags, int rank, int world, double timeout_s, void* stream)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    if (!flags || world < 1 || world > PEER_MAX || rank < 0 || rank >= world)
        return c.fail(cudaErrorInvalidValue, "peer_barrier arguments");
    if (world == 1) return 0;
    PeerFlags f;
    for (int r = 0; r < PEER_MAX; ++r) f.p[r] = r < world ? static_cast<unsigned*>(flags[r]) : nullptr;
    const unsigned long long ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 10.0) * 1e9);
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, world, ns);
    TUPAN_CHECK(cudaGetLastError(), "peer_barrier_kernel");
    c.launches++;
    return 0;
}

/* Barriers that gave up waiting since the last call (synchronises the device); < 0: CUDA error. */
long long tupan_cuda_peer_timeouts(void)
{
    unsigned int hits = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&hits, peer_timeouts, sizeof(hits)) != cudaSuccess) return -1;
    if (hits != 0 && cudaMemcpyToSymbol(peer_timeouts, &zero, sizeof(zero)) != cudaSuccess) return -1;
    return (long long)hits;
}

}  // extern "C"
