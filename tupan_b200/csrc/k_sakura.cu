// k_sakura.cu -- sakura on the pair engine, and the two-body Kepler kernel.
#include "kepler.cuh"
#include "runtime.cuh"

namespace tupan {
// scal = dt, flag  (libtupan.h:203-229)
enum { SAKURA_JOBS_CAP = 1 << 16 };
static DevBuf g_jobs;        // SAKURA_JOBS_CAP entries + the counter; lives as long as the library
static inline SakuraParams<real_t> sakura_params(const double* s)
{
    SakuraParams<real_t> p;
    p.dt = (real_t)s[0];
    p.flag = (int)s[1];
    // the list of pairs handed to the clean-up launch (kepler.cuh); allocated once, never resized,
    // so that pointers captured in a CUDA graph stay valid
    const size_t bytes = (size_t)SAKURA_JOBS_CAP * SAKURA_JOB_REALS * sizeof(real_t);
    const bool fresh = g_jobs.p == nullptr;
    char* base = static_cast<char*>(g_jobs.ensure(bytes + 256));
    if (base && fresh) cudaMemset(base + bytes, 0, 256);
    p.jobs = reinterpret_cast<real_t*>(base);
    p.njobs = base ? reinterpret_cast<unsigned*>(base + bytes) : nullptr;
    p.jobs_cap = base ? (unsigned)SAKURA_JOBS_CAP : 0u;
    return p;
}
// one straight-line variant per flag value (as for the PN levels, k_pn.cu)
typedef SakuraOp<real_t, 1> SakuraP1;
typedef SakuraOp<real_t, -1> SakuraM1;
typedef SakuraOp<real_t, 2> SakuraP2;
typedef SakuraOp<real_t, -2> SakuraM2;
typedef SakuraOp<real_t, 0> SakuraNone;
TUPAN_DEFINE_VTABLE(vt_sakura_p1, SakuraP1, "sakura_kernel", 8, 6, 2, 0, sakura_params)
TUPAN_DEFINE_VTABLE(vt_sakura_m1, SakuraM1, "sakura_kernel", 8, 6, 2, 0, sakura_params)
TUPAN_DEFINE_VTABLE(vt_sakura_p2, SakuraP2, "sakura_kernel", 8, 6, 2, 0, sakura_params)
TUPAN_DEFINE_VTABLE(vt_sakura_m2, SakuraM2, "sakura_kernel", 8, 6, 2, 0, sakura_params)
TUPAN_DEFINE_VTABLE(vt_sakura_none, SakuraNone, "sakura_kernel", 8, 6, 2, 0, sakura_params)

static const KernelVTable* sakura_vtable_for_flag(double flag)
{
    switch ((int)flag) {
        case 1: return &vt_sakura_p1;
        case -1: return &vt_sakura_m1;
        case 2: return &vt_sakura_p2;
        case -2: return &vt_sakura_m2;
        default: return &vt_sakura_none;
    }
}

// Read and reset the count of pairs whose Kepler sub-stepping hit the 2^MAX_DOUBLINGS bound
// (synchronises the device).  < 0 on a CUDA error.
long long kepler_limit_take()
{
    unsigned int hits = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&hits, kepler_limit_hits, sizeof(hits)) != cudaSuccess) return -1;
    if (hits != 0 && cudaMemcpyToSymbol(kepler_limit_hits, &zero, sizeof(zero)) != cudaSuccess) return -1;
    return (long long)hits;
}
long long kepler_cleanup_count()
{
    unsigned long long n = 0;
    if (cudaMemcpyFromSymbol(&n, kepler_cleanup_total, sizeof(n)) != cudaSuccess) return -1;
    return (long long)n;
}
static int kepler_limit_check(const char* where)
{
    const long long hits = kepler_limit_take();
    if (hits == 0) return 0;
    Context& c = ctx();
    c.last_error = hits < 0 ? (int)cudaGetLastError() : -2;
    snprintf(c.last_msg, sizeof(c.last_msg),
             "tupan_cuda: %s: %lld pair(s) needed more than 2^%d Kepler sub-steps (softened tight binary); "
             "result not converged", where, hits, (int)FULL_DOUBLINGS);
    fprintf(stderr, "%s\n", c.last_msg);
    return c.last_error;
}

// sakura: the pair engine entry points per flag, plus the sub-step-limit check on the synchronous path
namespace {
int sakura_rw(const double* s) { return vt_sakura_p1.row_width(s); }
int sakura_na(const double* s) { return vt_sakura_p1.n_acc(s); }
int sakura_host(long long ni, const real_t* const* hi, long long nj, const real_t* const* hj, const double* s,
                real_t* const* ho)
{
    int rc = sakura_vtable_for_flag(s[1])->run_host(ni, hi, nj, hj, s, ho);
    if (rc) return rc;
    return kepler_limit_check("sakura_kernel");
}
int sakura_dev(long long ni, const real_t* const* di, long long nj, const real_t* const* dj, const double* s,
               real_t* const* dout, cudaStream_t st)
{ return sakura_vtable_for_flag(s[1])->run_dev(ni, di, nj, dj, s, dout, st); }
int sakura_pack(long long nj, const real_t* const* dj, const double* s, real_t* packed, cudaStream_t st)
{ return vt_sakura_p1.pack(nj, dj, s, packed, st); }
int sakura_slots(long long ni, long long rows, const double* s)
{ return sakura_vtable_for_flag(s[1])->sweep_slots(ni, rows, s); }
int sakura_sweep(long long ni, const real_t* const* di, const real_t* packed, long long j0, long long j1,
                 const double* s, real_t* partial, int slot0, cudaStream_t st)
{ return sakura_vtable_for_flag(s[1])->sweep(ni, di, packed, j0, j1, s, partial, slot0, st); }
int sakura_fin(long long ni, const real_t* const* di, const real_t* partial, int nslots, const double* s,
               real_t* const* dout, cudaStream_t st)
{ return sakura_vtable_for_flag(s[1])->finalize(ni, di, partial, nslots, s, dout, st); }
int sakura_multi(long long ni, const real_t* const* di, int nseg, const real_t* const* sp, const long long* sr,
                 const double* s, real_t* partial, int slot0, cudaStream_t st)
{ return sakura_vtable_for_flag(s[1])->sweep_multi(ni, di, nseg, sp, sr, s, partial, slot0, st); }
}  // namespace
extern const KernelVTable vt_sakura = {"sakura_kernel", 8, 6, 2, 0, sakura_rw, sakura_na, sakura_host, sakura_dev,
                                       sakura_pack, sakura_slots, sakura_sweep, sakura_fin, sakura_multi};

// kepler_solver_kernel: arrays of 2*pairs bodies; scal = dt.  Device pointers.
int kepler_run_dev(long long pairs, const real_t* const* din, double dt, real_t* const* dout, cudaStream_t st)
{
    if (pairs <= 0) return 0;
    InRefs<real_t> in;
    OutRefs<real_t> out;
    for (int k = 0; k < MAX_IN; ++k) in.p[k] = k < 8 ? din[k] : nullptr;
    for (int k = 0; k < MAX_OUT; ++k) out.p[k] = k < 6 ? dout[k] : nullptr;
    kepler_pairs_kernel<real_t><<<(unsigned)((pairs + 63) / 64), 64, 0, st>>>(in, pairs, (real_t)dt, out);
    TUPAN_CHECK(cudaGetLastError(), "kepler_pairs_kernel");
    ctx().launches++;
    return 0;
}

// Host pointers; outputs may alias inputs (extensions.py:642-646): all inputs are copied to
// the device before any output is written back.
int kepler_run_host(long long pairs, const real_t* const* hin, double dt, real_t* const* hout)
{
    Context& c = ctx();
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.init();
    if (rc) return rc;
    if (pairs <= 0) return 0;
    const size_t n = (size_t)(2 * pairs), bytes = n * sizeof(real_t);
    const real_t* din[8];
    real_t* dout[6];
    // the reference calls this with ONE binary (16 + 12 values): one pinned block, one copy each way
    const size_t stride = (n + 3) / 4 * 4;
    if (14 * stride * sizeof(real_t) <= (size_t)(1 << 20)) {
        real_t* hs = static_cast<real_t*>(c.stage_host.ensure(14 * stride * sizeof(real_t)));
        real_t* ds = static_cast<real_t*>(c.stage_dev.ensure(14 * stride * sizeof(real_t)));
        if (!hs || !ds) return c.fail(cudaErrorMemoryAllocation, "kepler staging block");
        for (int k = 0; k < 8; ++k) {
            memcpy(hs + k * stride, hin[k], bytes);
            din[k] = ds + k * stride;
        }
        for (int k = 0; k < 6; ++k) dout[k] = ds + (8 + k) * stride;
        TUPAN_CHECK(cudaMemcpyAsync(ds, hs, 8 * stride * sizeof(real_t), cudaMemcpyHostToDevice, c.stream),
                    "H2D kepler block");
        rc = kepler_run_dev(pairs, din, dt, dout, c.stream);
        if (rc) return rc;
        TUPAN_CHECK(cudaMemcpyAsync(hs + 8 * stride, ds + 8 * stride, 6 * stride * sizeof(real_t),
                                    cudaMemcpyDeviceToHost, c.stream), "D2H kepler block");
        TUPAN_CHECK(cudaStreamSynchronize(c.stream), "synchronize");
        for (int k = 0; k < 6; ++k) memcpy(hout[k], hs + (8 + k) * stride, bytes);
        return kepler_limit_check("kepler_solver_kernel");
    }
    for (int k = 0; k < 8; ++k) {
        real_t* d = static_cast<real_t*>(c.in_i[k].ensure(bytes));
        if (!d) return c.fail(cudaErrorMemoryAllocation, "kepler in");
        TUPAN_CHECK(cudaMemcpyAsync(d, hin[k], bytes, cudaMemcpyHostToDevice, c.stream), "H2D kepler");
        din[k] = d;
    }
    for (int k = 0; k < 6; ++k) {
        dout[k] = static_cast<real_t*>(c.outb[k].ensure(bytes));
        if (!dout[k]) return c.fail(cudaErrorMemoryAllocation, "kepler out");
    }
    rc = kepler_run_dev(pairs, din, dt, dout, c.stream);
    if (rc) return rc;
    for (int k = 0; k < 6; ++k)
        TUPAN_CHECK(cudaMemcpyAsync(hout[k], dout[k], bytes, cudaMemcpyDeviceToHost, c.stream), "D2H kepler");
    TUPAN_CHECK(cudaStreamSynchronize(c.stream), "synchronize");
    return kepler_limit_check("kepler_solver_kernel");
}
}  // namespace tupan
