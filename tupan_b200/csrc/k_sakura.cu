// k_sakura.cu -- sakura on the pair engine, and the two-body Kepler kernel.
#include "kepler.cuh"
#include "runtime.cuh"

namespace tupan {
// scal = dt, flag  (libtupan.h:203-229)
static inline SakuraParams<real_t> sakura_params(const double* s)
{
    SakuraParams<real_t> p;
    p.dt = (real_t)s[0];
    p.flag = (int)s[1];
    return p;
}
TUPAN_DEFINE_VTABLE(vt_sakura, SakuraOp<real_t>, "sakura_kernel", 8, 6, 2, 0, sakura_params)

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
    const size_t bytes = (size_t)(2 * pairs) * sizeof(real_t);
    const real_t* din[8];
    real_t* dout[6];
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
    return 0;
}
}  // namespace tupan
