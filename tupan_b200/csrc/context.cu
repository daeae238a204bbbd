// context.cu -- library context, Part-2 entry points of include/libtupan_cuda.h, FMA peak probe.
#include "runtime.cuh"
#include "../../include/libtupan_cuda.h"

namespace tupan {

void* DevBuf::ensure(size_t bytes)
{
    if (bytes <= cap && p) return p;
    // The old block is RETIRED, not freed: a CUDA graph captured by a caller (Integrator steps) has
    // its address baked in and must keep replaying against valid memory; blocks grow geometrically,
    // so what is retired stays below the size of the live block.
    if (p) retired.push_back(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 2 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        p = nullptr;
        return nullptr;
    }
    cap = want;
    return p;
}
void DevBuf::release()
{
    if (p) cudaFree(p);
    for (void* q : retired) cudaFree(q);
    retired.clear();
    p = nullptr;
    cap = 0;
}

void* HostBuf::ensure(size_t bytes)
{
    if (bytes <= cap && p) return p;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    if (cudaMallocHost(&p, want) != cudaSuccess) {
        p = nullptr;
        return nullptr;
    }
    cap = want;
    return p;
}
void HostBuf::release()
{
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

Context& ctx()
{
    static Context c;
    return c;
}

int Context::fail(cudaError_t e, const char* where)
{
    last_error = (int)e;
    snprintf(last_msg, sizeof(last_msg), "tupan_cuda: %s: %s", where, cudaGetErrorString(e));
    fprintf(stderr, "%s\n", last_msg);
    cudaGetLastError();  // clear the sticky flag of non-fatal errors
    return last_error ? last_error : -1;
}

int Context::init()
{
    if (ready) {
        // One context (stream, staging buffers) per process, bound to the device that was current
        // at the first call: one process per GPU, as torchrun launches them.  A caller that
        // switches devices afterwards must hear about it instead of getting wrong-device pointers.
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != device) {
            snprintf(last_msg, sizeof(last_msg),
                     "tupan_cuda: the library context lives on device %d but device %d is current "
                     "(one process per GPU)", device, cur);
            fprintf(stderr, "%s\n", last_msg);
            last_error = (int)cudaErrorInvalidDevice;
            return last_error;
        }
        return 0;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(e != cudaSuccess ? e : cudaErrorNoDevice, "no CUDA device");
    if ((e = cudaGetDevice(&device)) != cudaSuccess) return fail(e, "cudaGetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
    if (prop.major < 10) {
        snprintf(last_msg, sizeof(last_msg), "tupan_cuda: built for sm_100a, found sm_%d%d", prop.major, prop.minor);
        fprintf(stderr, "%s\n", last_msg);
        last_error = (int)cudaErrorNoKernelImageForDevice;
        return last_error;
    }
    info.sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    ready = true;
    return 0;
}

extern const KernelVTable vt_phi, vt_acc, vt_acc_jerk, vt_snap_crackle, vt_tstep, vt_pnacc, vt_nreg_x, vt_nreg_v,
    vt_sakura;

const KernelVTable* vtable(int kernel)
{
    switch (kernel) {
        case K_PHI: return &vt_phi;
        case K_ACC: return &vt_acc;
        case K_ACC_JERK: return &vt_acc_jerk;
        case K_SNAP_CRACKLE: return &vt_snap_crackle;
        case K_TSTEP: return &vt_tstep;
        case K_PNACC: return &vt_pnacc;
        case K_NREG_X: return &vt_nreg_x;
        case K_NREG_V: return &vt_nreg_v;
        case K_SAKURA: return &vt_sakura;
        default: return nullptr;
    }
}

int kepler_run_dev(long long pairs, const real_t* const* din, double dt, real_t* const* dout, cudaStream_t st);
long long kepler_limit_take();
long long kepler_cleanup_count();

// ---- |x| minimum (fused tstep follow-up) ------------------------------------------------
__global__ void abs_min_kernel(const real_t* __restrict__ v, long long n, real_t* __restrict__ out)
{
    // single CTA; n is O(N) and this runs once per step
    __shared__ real_t sm[32];
    real_t m = (real_t)INFINITY;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        real_t a = v[i] < 0 ? -v[i] : v[i];
        m = a < m ? a : m;
    }
    for (int off = 16; off > 0; off >>= 1) {
        real_t o = __shfl_xor_sync(0xffffffffu, m, off);
        m = o < m ? o : m;
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : (real_t)INFINITY;
        for (int off = 16; off > 0; off >>= 1) {
            real_t o = __shfl_xor_sync(0xffffffffu, m, off);
            m = o < m ? o : m;
        }
        if (threadIdx.x == 0) *out = m;
    }
}

// ---- FMA peak probe ---------------------------------------------------------------------
// 8 independent FMA chains per thread, 1024 threads resident per SM: the pipe is the only
// limit.  Reports what the FMA pipe of this precision sustains on this board right now
// (power cap and clocks included) -- the roofline denominator bench.py uses.
__global__ void __launch_bounds__(256) fma_probe_kernel(real_t* out, int iters, real_t a, real_t b)
{
    real_t x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    real_t s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == (real_t)123.456) out[0] = s;  // never true; keeps the chains alive
}

// The other shapes an FP-pipe instruction can take (round 2, tools/microbench2.cu): an FP64 instruction
// holds the pipe for 2 clocks whatever it is, but one that has to fetch three distinct registers
// (the chains above: x, a, b) needs a third clock unless the operand-reuse cache serves one.
//   KIND 1: x = x + a   KIND 2: x = x * a   KIND 3: x = fma(x, y, 1)   (two registers + immediate)
template <int KIND>
__global__ void __launch_bounds__(256) pipe_probe_kernel(real_t* out, int iters, real_t a)
{
    real_t x[8], y[8];
    // y comes from memory so that it lives in registers (a launch constant would be folded into the
    // instruction as a constant-bank operand: a different operand shape)
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = (real_t)(threadIdx.x + k); y[k] = out[8 + k] * a; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (KIND == 1) x[k] = x[k] + y[k];
                if (KIND == 2) x[k] = x[k] * y[k];
                if (KIND == 3) x[k] = fma(x[k], y[k], (real_t)1);
            }
        }
    }
    real_t s = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
    if (s == (real_t)123.456) out[0] = s;
}

}  // namespace tupan

using namespace tupan;

extern "C" {

int tupan_cuda_info(int kernel, int* n_in, int* n_out, int* n_scal, int* flops)
{
    if (kernel == K_KEPLER) {
        if (n_in) *n_in = 8;
        if (n_out) *n_out = 6;
        if (n_scal) *n_scal = 1;
        if (flops) *flops = 0;
        return 0;
    }
    const KernelVTable* vt = vtable(kernel);
    if (!vt) return -1;
    if (n_in) *n_in = vt->n_in;
    if (n_out) *n_out = vt->n_out;
    if (n_scal) *n_scal = vt->n_scal;
    if (flops) *flops = vt->flops;
    return 0;
}

int tupan_cuda_run_dev(int kernel, long long ni, const void* const* iarr, long long nj, const void* const* jarr,
                       const double* scal, void* const* out, void* stream)
{
    const KernelVTable* vt = vtable(kernel);
    if (!vt) return -1;
    return vt->run_dev(ni, (const real_t* const*)iarr, nj, (const real_t* const*)jarr, scal, (real_t* const*)out,
                       (cudaStream_t)stream);
}

int tupan_cuda_kepler_dev(long long pairs, const void* const* arr, double dt, void* const* out, void* stream)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    return kepler_run_dev(pairs, (const real_t* const*)arr, dt, (real_t* const*)out, (cudaStream_t)stream);
}

long long tupan_cuda_kepler_limit_hits(void)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.init()) return -1;
    return kepler_limit_take();
}

long long tupan_cuda_kepler_cleanup_pairs(void)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.init()) return -1;
    return kepler_cleanup_count();
}

int tupan_cuda_row_width(int kernel, const double* scal)
{
    const KernelVTable* vt = vtable(kernel);
    return vt ? vt->row_width(scal) : -1;
}
int tupan_cuda_n_acc(int kernel, const double* scal)
{
    const KernelVTable* vt = vtable(kernel);
    return vt ? vt->n_acc(scal) : -1;
}
int tupan_cuda_pack_dev(int kernel, long long nj, const void* const* jarr, const double* scal, void* packed,
                        void* stream)
{
    const KernelVTable* vt = vtable(kernel);
    if (!vt) return -1;
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    return vt->pack(nj, (const real_t* const*)jarr, scal, (real_t*)packed, (cudaStream_t)stream);
}
int tupan_cuda_sweep_slots(int kernel, long long ni, long long rows, const double* scal)
{
    const KernelVTable* vt = vtable(kernel);
    if (!vt) return -1;
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.init()) return -1;
    return vt->sweep_slots(ni, rows, scal);
}
/* slots a multi-owner sweep writes (the plan is made for the logical rows: whole tiles per owner) */
int tupan_cuda_sweep_multi_slots(int kernel, long long ni, int nseg, const long long* seg_rows, const double* scal)
{
    const KernelVTable* vt = vtable(kernel);
    if (!vt || !seg_rows) return -1;
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.init()) return -1;
    return vt->sweep_slots(ni, multi_logical_rows(nseg, seg_rows), scal);
}
int tupan_cuda_sweep_multi_dev(int kernel, long long ni, const void* const* iarr, int nseg,
                               const void* const* seg_ptr, const long long* seg_rows, const double* scal,
                               void* partial, int slot0, void* stream)
{
    const KernelVTable* vt = vtable(kernel);
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    if (!vt || !partial || !seg_ptr || !seg_rows) return c.fail(cudaErrorInvalidValue, "sweep_multi arguments");
    return vt->sweep_multi(ni, (const real_t* const*)iarr, nseg, (const real_t* const*)seg_ptr, seg_rows, scal,
                           (real_t*)partial, slot0, (cudaStream_t)stream);
}
/* The launch shape choose_plan() gives a kernel for ni x nj pairs.  Needs no device: the model
 * only uses the SM count (148 until a context is bound to a GPU). */
int tupan_cuda_plan_query(int kernel, long long ni, long long nj, const double* scal, int* lane_split, int* js_log2,
                          int* jg)
{
    const KernelVTable* vt = vtable(kernel);
    if (!vt) return -1;
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    const double zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int g = vt->sweep_slots(ni, nj, scal ? scal : zero);
    if (lane_split) *lane_split = c.last_plan.lane_split;
    if (js_log2) *js_log2 = c.last_plan.js_log2;
    if (jg) *jg = c.last_plan.jg;
    return g > 0 ? 0 : -1;
}
int tupan_cuda_sweep_dev(int kernel, long long ni, const void* const* iarr, const void* packed, long long j0,
                         long long j1, const double* scal, void* partial, int slot0, void* stream)
{
    const KernelVTable* vt = vtable(kernel);
    if (!vt) return -1;
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    return vt->sweep(ni, (const real_t* const*)iarr, (const real_t*)packed, j0, j1, scal, (real_t*)partial, slot0,
                     (cudaStream_t)stream);
}
int tupan_cuda_finalize_dev(int kernel, long long ni, const void* const* iarr, const void* partial, int nslots,
                            const double* scal, void* const* out, void* stream)
{
    const KernelVTable* vt = vtable(kernel);
    if (!vt) return -1;
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    return vt->finalize(ni, (const real_t* const*)iarr, (const real_t*)partial, nslots, scal, (real_t* const*)out,
                        (cudaStream_t)stream);
}

int tupan_cuda_abs_min_dev(long long n, const void* d_values, void* d_min, void* stream)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    abs_min_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((const real_t*)d_values, n, (real_t*)d_min);
    TUPAN_CHECK(cudaGetLastError(), "abs_min_kernel");
    c.launches++;
    return 0;
}

int tupan_cuda_init(void)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    return c.init();
}
int tupan_cuda_last_error(char* msg, int msg_len)
{
    Context& c = ctx();
    if (msg && msg_len > 0) {
        strncpy(msg, c.last_msg, (size_t)msg_len - 1);
        msg[msg_len - 1] = 0;
    }
    return c.last_error;
}
void tupan_cuda_clear_error(void)
{
    Context& c = ctx();
    c.last_error = 0;
    c.last_msg[0] = 0;
}
void tupan_cuda_force_plan(int lane_split, int js_log2, int jg)
{
    Context& c = ctx();
    c.forced.lane_split = lane_split;
    c.forced.js_log2 = js_log2;
    c.forced.jg = jg;
}
void tupan_cuda_last_plan(int* lane_split, int* js_log2, int* jg)
{
    Context& c = ctx();
    if (lane_split) *lane_split = c.last_plan.lane_split;
    if (js_log2) *js_log2 = c.last_plan.js_log2;
    if (jg) *jg = c.last_plan.jg;
}
void tupan_cuda_set_timing(int enable)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (enable && !c.evr) {
        c.evr = new cudaEvent_t[Context::TIME_RING][6];
        c.evmark = new unsigned[Context::TIME_RING];
        for (int q = 0; q < Context::TIME_RING; ++q) {
            c.evmark[q] = 0;
            for (int k = 0; k < 6; ++k)
                if (cudaEventCreate(&c.evr[q][k]) != cudaSuccess) {
                    c.fail(cudaGetLastError(), "timing events");
                    c.evr = nullptr;       // leaked on purpose: timing simply stays off
                    c.timing = false;
                    return;
                }
        }
    }
    if (enable) c.nsets = 0;
    c.timing = enable != 0 && c.evr != nullptr;
}
void tupan_cuda_last_times(float* h2d, float* pack, float* pair, float* fin, float* d2h)
{
    Context& c = ctx();
    float t[5] = {0, 0, 0, 0, 0};
    if (c.evr && c.nsets > 0) c.set_times(c.cur_set(), t);
    c.last.h2d_ms = t[0]; c.last.pack_ms = t[1]; c.last.pair_ms = t[2]; c.last.finalize_ms = t[3]; c.last.d2h_ms = t[4];
    if (h2d) *h2d = c.last.h2d_ms;
    if (pack) *pack = c.last.pack_ms;
    if (pair) *pair = c.last.pair_ms;
    if (fin) *fin = c.last.finalize_ms;
    if (d2h) *d2h = c.last.d2h_ms;
}
int tupan_cuda_sum_times(float* sums5, long long* calls)
{
    Context& c = ctx();
    float acc[5] = {0, 0, 0, 0, 0};
    long long n = 0;
    if (c.evr) {
        const long long have = c.nsets < Context::TIME_RING ? c.nsets : Context::TIME_RING;
        for (long long q = 0; q < have; ++q) {
            float t[5];
            if (!c.set_times((int)q, t)) continue;
            for (int k = 0; k < 5; ++k) acc[k] += t[k];
            n++;
        }
    }
    if (sums5) for (int k = 0; k < 5; ++k) sums5[k] = acc[k];
    if (calls) *calls = n;
    return c.nsets > Context::TIME_RING ? 1 : 0;     // 1: the ring wrapped, older calls are not in the sums
}
long long tupan_cuda_launch_count(void) { return ctx().launches; }
void tupan_cuda_count_launches(long long n) { ctx().launches += n; }
int tupan_cuda_sm_count(void)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.init()) return -1;
    return c.info.sm_count;
}
int tupan_cuda_real_bytes(void) { return (int)sizeof(real_t); }

int tupan_cuda_pipe_probe(int kind, double ms, double* tera_ops)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    if (kind < 1 || kind > 3) return c.fail(cudaErrorInvalidValue, "pipe_probe: kind 1..3");
    real_t* d = static_cast<real_t*>(c.partial.ensure(256));
    if (!d) return c.fail(cudaErrorMemoryAllocation, "probe buffer");
    TUPAN_CHECK(cudaMemsetAsync(d, 0, 256, c.stream), "probe buffer");
    cudaEvent_t e0, e1;
    TUPAN_CHECK(cudaEventCreate(&e0), "event");
    TUPAN_CHECK(cudaEventCreate(&e1), "event");
    const int grid = c.info.sm_count * 4, block = 256;
    const double ops_per_iter = 8.0 * 32 * (double)grid * block;
    int iters = 1000;
    float t = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, c.stream);
        if (kind == 1) pipe_probe_kernel<1><<<grid, block, 0, c.stream>>>(d, iters, (real_t)1e-6);
        if (kind == 2) pipe_probe_kernel<2><<<grid, block, 0, c.stream>>>(d, iters, (real_t)0.999999);
        if (kind == 3) pipe_probe_kernel<3><<<grid, block, 0, c.stream>>>(d, iters, (real_t)0.999999);
        cudaEventRecord(e1, c.stream);
        TUPAN_CHECK(cudaEventSynchronize(e1), "pipe probe");
        cudaEventElapsedTime(&t, e0, e1);
        c.launches++;
        if (rep < 2 && t > 0) {
            double scale = ms / t;
            if (scale > 50) scale = 50;
            iters = (int)(iters * scale) + 1;
        }
    }
    if (tera_ops) *tera_ops = ops_per_iter * iters / (t * 1e-3) * 1e-12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

int tupan_cuda_fma_peak(double ms, double* tflops, double* sm_mhz_effective)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    real_t* d = static_cast<real_t*>(c.partial.ensure(256));
    if (!d) return c.fail(cudaErrorMemoryAllocation, "probe buffer");
    cudaEvent_t e0, e1;
    TUPAN_CHECK(cudaEventCreate(&e0), "event");
    TUPAN_CHECK(cudaEventCreate(&e1), "event");
    const int grid = c.info.sm_count * 4, block = 256;
    const double flop_per_iter = 2.0 * 8 * 16 * (double)grid * block;
    int iters = 2000;
    float t = 0;
    // warm up, then size the run to ~ms and time it
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, c.stream);
        fma_probe_kernel<<<grid, block, 0, c.stream>>>(d, iters, (real_t)0.999999, (real_t)1e-6);
        cudaEventRecord(e1, c.stream);
        TUPAN_CHECK(cudaEventSynchronize(e1), "fma probe");
        cudaEventElapsedTime(&t, e0, e1);
        c.launches++;
        if (rep < 2 && t > 0) {
            double scale = ms / t;
            if (scale > 50) scale = 50;
            iters = (int)(iters * scale) + 1;
        }
    }
    const double fl = flop_per_iter * iters / (t * 1e-3);
    if (tflops) *tflops = fl * 1e-12;
    if (sm_mhz_effective) {
        // lanes per SM per clock: fp64 64, fp32 128 on sm_100
        const double lanes = sizeof(real_t) == 8 ? 64.0 : 128.0;
        *sm_mhz_effective = fl / (2.0 * lanes * c.info.sm_count) * 1e-6;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

}  // extern "C"
