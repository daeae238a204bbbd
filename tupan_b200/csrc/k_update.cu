// k_update.cu -- Part 3 of include/libtupan_cuda.h: the O(N) integrator updates on device-
// resident state, so that a whole Hermite / SIA / Sakura step runs on
// one stream without the particle arrays ever leaving HBM and without a host round trip for
// the time-step.
//
// What the reference does on the host with numpy between two kernel calls (and what each
// routine here replaces):
//   Base.get_base_tstep / get_min_block_tstep   integrator/__init__.py:48-78   -> step_begin
//   H2/H4/H6/H8.epredict / ecorrect             integrator/hermite.py:25-283   -> hermite_*
//   drift_n / kick_n, sakura half drifts        integrator/sia.py:64-84, sakura.py:25-48 -> axpy
//   nreg_x / nreg_v rescaling                   integrator/nreg.py:28-30,51-53  -> scale
//   t_curr / time / nstep / tstep bookkeeping   hermite.py:398-401, sia.py:1108-1113 -> step_end
//   kinetic / potential energy, min |tstep|, the sakura step criterion
//                                              particles/body.py:262-306,364-368; sakura.py:100-117 -> reduce
//
// The step size lives in a device control block (8 doubles): step_begin derives it on the
// device from the reduced minimum time-step, every update kernel reads it from there.
//
// This translation unit is compiled with -fmad=false and evaluates each update in the
// reference's operation order (numpy evaluates `a * tau / 2 + v` as ((a*tau)/2)+v, one
// rounding per operation, in REAL), so the O(N) part of a step is bit-identical to the
// reference's and the only difference left is the summation order inside the pair kernels.
#include <float.h>

#include "runtime.cuh"
#include "../../include/libtupan_cuda.h"

namespace tupan {

enum { CTL_T_CURR = 0, CTL_TAU = 1, CTL_T_END = 2, CTL_ETA = 3, CTL_NSTEPS = 4, CTL_DONE = 5, CTL_TAU_BASE = 6,
       CTL_MIN_TS = 7 };

// ---------------------------------------------------------------------------------------
// step_begin: tau for the next step.  One thread; a few dozen flops.
// ---------------------------------------------------------------------------------------
__global__ void step_begin_kernel(double* __restrict__ ctl, const double* __restrict__ d_min, int is_real32)
{
    const double t_curr = ctl[CTL_T_CURR], t_end = ctl[CTL_T_END], eta = ctl[CTL_ETA];
    // the driver loop's condition, simulation.py:193: a step past the end is a no-op (tau = 0)
    const bool done = !(fabs(t_curr) < fabs(t_end));
    // Base.get_base_tstep, integrator/__init__.py:48-57
    double dt = fmin(fabs(t_end) - fabs(t_curr), fabs(eta));
    dt = fmax(dt, fabs(t_end) * (2.0 * DBL_EPSILON));
    double tau = copysign(dt, eta);
    ctl[CTL_TAU_BASE] = tau;
    if (d_min != nullptr && !done) {
        // Base.get_min_block_tstep, integrator/__init__.py:59-78
        const double min_ts = *d_min;
        ctl[CTL_MIN_TS] = min_ts;
        // np.log2 of a REAL scalar, minus 1 (weak python int keeps REAL), truncated toward zero
        const double l2 = is_real32 ? (double)(log2f((float)min_ts) - 1.0f) : log2(min_ts) - 1.0;
        const int power = (int)l2;
        double min_bts = ldexp(1.0, power);
        const double t_next = t_curr + min_bts;
        // python's float % takes the sign of the divisor; only "== 0" matters here
        int guard = 0;
        while (fmod(t_next, min_bts) != 0.0 && guard++ < 1100) min_bts *= 0.5;
        min_bts = copysign(min_bts, tau);
        if (fabs(min_bts) > fabs(tau)) min_bts = tau;
        tau = min_bts;
    }
    ctl[CTL_TAU] = done ? 0.0 : tau;
    ctl[CTL_DONE] = done ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------
// Hermite predictor / corrector.  ND = order/2 derivative sets (a | a j | a j s | a j s c).
// ---------------------------------------------------------------------------------------
struct HermiteRefs {
    real_t* rv[6];           // rx ry rz vx vy vz : state being advanced (ps1)
    real_t* rv0[6];          // copy of the state at the start of the step (ps0)
    const real_t* d0[12];    // derivatives of ps0: ax ay az jx jy jz sx sy sz cx cy cz
    const real_t* d1[12];    // derivatives of ps1 (corrector only)
    const double* ctl;
    long long n;
};

// hermite.py:25-43 (H2), 75-91 (H4), 127-158 (H6), 202-242 (H8): Taylor series in Horner form,
//   r += ((((c tau/5 + s) tau/4 + j) tau/3 + a) tau/2 + v) tau ,  v likewise one order lower.
template <int ND>
__global__ void __launch_bounds__(256) hermite_predict_kernel(const HermiteRefs a)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const real_t tau = (real_t)a.ctl[CTL_TAU];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const real_t r = a.rv[c][i], v = a.rv[3 + c][i];
        a.rv0[c][i] = r;
        a.rv0[3 + c][i] = v;
        real_t L[ND + 1];
        L[0] = v;
#pragma unroll
        for (int q = 0; q < ND; ++q) L[q + 1] = a.d0[3 * q + c][i];
        real_t x = L[ND];
#pragma unroll
        for (int k = ND; k >= 1; --k) x = x * tau / (real_t)(k + 1) + L[k - 1];
        real_t y = L[ND];
#pragma unroll
        for (int k = ND - 1; k >= 1; --k) y = y * tau / (real_t)(k + 1) + L[k];
        a.rv[c][i] = r + x * tau;
        a.rv[3 + c][i] = v + y * tau;
    }
}

// hermite.py:45-57 (H2), 93-121 (H4), 160-196 (H6), 244-283 (H8): v from the derivative
// pairs, then r from the same formula one order lower using the NEW v.
template <int ND> TUPAN_DEV real_t hermite_corr(const real_t (&p0)[ND + 1], const real_t (&p1)[ND + 1], real_t tau);
// p[0] = the quantity being integrated (v or r) at the start; p[1..ND] = its derivatives.
template <> TUPAN_DEV real_t hermite_corr<1>(const real_t (&p0)[2], const real_t (&p1)[2], real_t tau)
{
    return (p0[1] + p1[1]) * tau / (real_t)2 + p0[0];
}
template <> TUPAN_DEV real_t hermite_corr<2>(const real_t (&p0)[3], const real_t (&p1)[3], real_t tau)
{
    return ((p0[2] - p1[2]) * tau / (real_t)6 + (p0[1] + p1[1])) * tau / (real_t)2 + p0[0];
}
template <> TUPAN_DEV real_t hermite_corr<3>(const real_t (&p0)[4], const real_t (&p1)[4], real_t tau)
{
    return (((p0[3] + p1[3]) * tau / (real_t)12 + (p0[2] - p1[2])) * tau / (real_t)5 + (p0[1] + p1[1])) * tau
               / (real_t)2
           + p0[0];
}
template <> TUPAN_DEV real_t hermite_corr<4>(const real_t (&p0)[5], const real_t (&p1)[5], real_t tau)
{
    return ((((p0[4] - p1[4]) * tau / (real_t)20 + (p0[3] + p1[3])) * tau / (real_t)3
             + (real_t)3 * (p0[2] - p1[2])) * tau / (real_t)14
            + (p0[1] + p1[1])) * tau / (real_t)2
           + p0[0];
}

template <int ND>
__global__ void __launch_bounds__(256) hermite_correct_kernel(const HermiteRefs a)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const real_t tau = (real_t)a.ctl[CTL_TAU];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const real_t r0 = a.rv0[c][i], v0 = a.rv0[3 + c][i];
        real_t p0[ND + 1], p1[ND + 1];
        p0[0] = v0;
        p1[0] = v0;
#pragma unroll
        for (int q = 0; q < ND; ++q) {
            p0[q + 1] = a.d0[3 * q + c][i];
            p1[q + 1] = a.d1[3 * q + c][i];
        }
        const real_t v1 = hermite_corr<ND>(p0, p1, tau);
        // r: same weights, one derivative lower: (r; v, a, j, ...) with the new v
        real_t q0[ND + 1], q1[ND + 1];
        q0[0] = r0;
        q1[0] = r0;
        q0[1] = v0;
        q1[1] = v1;
#pragma unroll
        for (int q = 1; q < ND; ++q) {
            q0[q + 1] = p0[q];
            q1[q + 1] = p1[q];
        }
        const real_t r1 = hermite_corr<ND>(q0, q1, tau);
        a.rv[c][i] = r1;
        a.rv[3 + c][i] = v1;
    }
}

// ---------------------------------------------------------------------------------------
// Individual block time-steps (absent from the reference, whose adaptive
// Hermite advances every particle with the shared minimum block step, hermite.py:343-401).
// Every particle carries its own time and power-of-two step.  block_predict brings ALL particles
// to the next block time (Taylor series of the derivatives they hold, Horner form); the force
// kernels then run rectangular -- active particles against everybody's predicted state --, and
// block_correct applies the Hermite corrector (the same hermite_corr<ND> as above) to the active
// ones, each with its own step.
//   state[3*k + c]: k = 0 r, 1 v, 2 a, 3 j, 4 s  (ND + 2 levels, component c)
//   pred [3*m + c]: m = 0 r, 1 v and, for ND >= 3 (snap_crackle needs them), 2 a, 3 j
// ---------------------------------------------------------------------------------------
struct BlockPredictRefs {
    const real_t* x[15];
    real_t* p[12];
    const real_t* time;
    double t_next;
    long long n;
};

template <int ND>
__global__ void __launch_bounds__(256) block_predict_kernel(const BlockPredictRefs a)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    constexpr int NL = ND + 2;                  // levels held: r v a j (s)
    constexpr int NP = ND >= 3 ? 4 : 2;         // levels predicted
    const real_t dt = (real_t)(a.t_next - (double)a.time[i]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        real_t L[NL];
#pragma unroll
        for (int k = 0; k < NL; ++k) L[k] = a.x[3 * k + c][i];
#pragma unroll
        for (int m = 0; m < NP; ++m) {
            real_t x = L[NL - 1];
#pragma unroll
            for (int k = NL - 1; k > m; --k) x = x * dt / (real_t)(k - m) + L[k - 1];
            a.p[3 * m + c][i] = x;
        }
    }
}

struct BlockCorrectRefs {
    const real_t* tau;       // per active particle
    const real_t* rv0[6];    // state at the start of the particle's own step
    const real_t* d0[9];     // its derivatives there: a j (s)
    const real_t* d1[9];     // derivatives at the block time
    real_t* rv[6];           // corrected state out
    long long n;
};

template <int ND>
__global__ void __launch_bounds__(256) block_correct_kernel(const BlockCorrectRefs a)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const real_t tau = a.tau[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const real_t r0 = a.rv0[c][i], v0 = a.rv0[3 + c][i];
        real_t p0[ND + 1], p1[ND + 1];
        p0[0] = v0;
        p1[0] = v0;
#pragma unroll
        for (int q = 0; q < ND; ++q) {
            p0[q + 1] = a.d0[3 * q + c][i];
            p1[q + 1] = a.d1[3 * q + c][i];
        }
        const real_t v1 = hermite_corr<ND>(p0, p1, tau);
        real_t q0[ND + 1], q1[ND + 1];
        q0[0] = r0;
        q1[0] = r0;
        q0[1] = v0;
        q1[1] = v1;
#pragma unroll
        for (int q = 1; q < ND; ++q) {
            q0[q + 1] = p0[q];
            q1[q + 1] = p1[q];
        }
        a.rv[c][i] = hermite_corr<ND>(q0, q1, tau);
        a.rv[3 + c][i] = v1;
    }
}

// New step of the particles that just arrived at block time t_next: the largest power of two
// not above the criterion ts, at most dt_max, at most twice the old step tau -- and twice only if
// t_next is a multiple of it (block steps stay commensurate).  Also stamps their new time.
// block_select: next block time t = min_i(time_i + dt_i) and the ordered list of the particles
// that reach it, in ONE launch of one CTA (N is a few 10^5: two passes over 2 arrays), so that a
// block step reads back 16 bytes once instead of synchronising for a minimum, a nonzero() and a
// count.  out2[0] = t, out2[1] = number of active particles; idx = their indices, ascending.
__global__ void __launch_bounds__(1024) block_select_kernel(long long n, const real_t* __restrict__ time,
                                                            const real_t* __restrict__ dt, double* __restrict__ out2,
                                                            long long* __restrict__ idx)
{
    __shared__ double smin[32];
    __shared__ int scount[32];
    __shared__ long long sbase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double m = INFINITY;
    for (long long i = tid; i < n; i += 1024) {
        const double t = (double)(time[i] + dt[i]);
        m = t < m ? t : m;
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, m, off);
        m = o < m ? o : m;
    }
    if (lane == 0) smin[warp] = m;
    __syncthreads();
    m = smin[lane];
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, m, off);
        m = o < m ? o : m;
    }
    const real_t tn = (real_t)m;          // every thread holds the minimum
    if (tid == 0) sbase = 0;
    __syncthreads();
    // ordered compaction, 1024 particles per round
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + tid;
        const bool act = i < n && (time[i] + dt[i]) == tn;
        const unsigned bal = __ballot_sync(0xffffffffu, act);
        if (lane == 0) scount[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            const int c = scount[w];
            if (w < warp) before += c;
            total += c;
        }
        const long long b0 = sbase;
        if (act) idx[b0 + before + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (tid == 0) sbase = b0 + total;
        __syncthreads();
    }
    if (tid == 0) {
        out2[0] = m;
        out2[1] = (double)sbase;
    }
}

__global__ void __launch_bounds__(256) block_quantize_kernel(long long n, const real_t* __restrict__ ts,
                                                             const real_t* __restrict__ tau, double t_next,
                                                             double dt_max, real_t* __restrict__ dt_new,
                                                             real_t* __restrict__ time_new)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int e;
    frexp((double)ts[i], &e);                     // ts = m 2^e, m in [0.5, 1)
    double cand = ldexp(1.0, e - 1);
    cand = cand < dt_max ? cand : dt_max;
    const double old = (double)tau[i], twice = 2.0 * old;
    const bool up = cand >= twice && fmod(t_next, twice) == 0.0;
    dt_new[i] = (real_t)(up ? twice : (cand < old ? cand : old));
    time_new[i] = (real_t)t_next;
}

// ---------------------------------------------------------------------------------------
// axpy: y[k] += x[k] * REAL(c_inner * (c_outer * tau)), k < narr  (drift, kick, += dr)
// scale: y[k] = x[k] / REAL(denom)
// ---------------------------------------------------------------------------------------
struct VecRefs {
    real_t* y[6];
    const real_t* x[6];
    int narr;
    long long n;
};

__global__ void __launch_bounds__(256) axpy_kernel(const VecRefs a, double c_outer, double c_inner,
                                                   const double* __restrict__ ctl)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    // the reference forms the sub-step in python doubles (e.g. d[0] * (d[0] * tau), sia.py:316,347)
    // and numpy casts it to REAL when it meets the array
    const double tau = ctl ? ctl[CTL_TAU] : 1.0;
    const real_t f = (real_t)(c_inner * (c_outer * tau));
    for (int k = 0; k < a.narr; ++k) a.y[k][i] = a.y[k][i] + a.x[k][i] * f;
}

__global__ void __launch_bounds__(256) scale_kernel(const VecRefs a, double denom)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const real_t d = (real_t)denom;
    for (int k = 0; k < a.narr; ++k) a.y[k][i] = a.x[k][i] / d;
}

// ---------------------------------------------------------------------------------------
// Post-Newtonian kick of the SIA integrators (kick_pn, sia.py:136-157) around the pnacc
// evaluation, with the bookkeeping of PNbodyMethods (particles/body.py:471-527):
//   phase 0:  v += (a tau + w)/2
//   phase 1:  pn_ke -= (v . m pna) tau;  pn_mv -= m pna tau;  pn_am -= (r x m pna) tau;
//             w = 2 pna tau - w;  v += (a tau + w)/2
// ---------------------------------------------------------------------------------------
struct PnRefs {
    real_t* v[3];
    const real_t* a[3];
    real_t* w[3];
    const real_t* pna[3];
    const real_t* mass;
    const real_t* r[3];
    real_t* pn_ke;
    real_t* pn_mv[3];
    real_t* pn_am[3];
    long long n;
};

template <int PHASE>
__global__ void __launch_bounds__(256) pn_kick_kernel(const PnRefs p, double c_outer, double c_inner,
                                                      const double* __restrict__ ctl)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const double tau_d = ctl ? ctl[CTL_TAU] : 1.0;
    const real_t tau = (real_t)(c_inner * (c_outer * tau_d));
    real_t v[3], a[3], w[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { v[c] = p.v[c][i]; a[c] = p.a[c][i]; w[c] = p.w[c][i]; }
    if (PHASE == 1) {
        const real_t m = p.mass[i];
        real_t f[3], r[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { f[c] = m * p.pna[c][i]; r[c] = p.r[c][i]; }
        p.pn_ke[i] = p.pn_ke[i] - ((v[0] * f[0] + v[1] * f[1]) + v[2] * f[2]) * tau;
#pragma unroll
        for (int c = 0; c < 3; ++c) p.pn_mv[c][i] = p.pn_mv[c][i] - f[c] * tau;
        p.pn_am[0][i] = p.pn_am[0][i] - (r[1] * f[2] - r[2] * f[1]) * tau;
        p.pn_am[1][i] = p.pn_am[1][i] - (r[2] * f[0] - r[0] * f[2]) * tau;
        p.pn_am[2][i] = p.pn_am[2][i] - (r[0] * f[1] - r[1] * f[0]) * tau;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            w[c] = (real_t)2 * p.pna[c][i] * tau - w[c];
            p.w[c][i] = w[c];
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) p.v[c][i] = v[c] + (a[c] * tau + w[c]) / (real_t)2;
}

// ---------------------------------------------------------------------------------------
// step_end: t_curr += tau; tstep[:] = tau; time += tau; nstep += 1
// (hermite.py:398-401, sia.py:1105-1113, sakura.py:136-139)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) step_end_kernel(long long n, real_t* __restrict__ time,
                                                       abi_uint* __restrict__ nstep, real_t* __restrict__ tstep,
                                                       double* ctl)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // every thread reads TAU and DONE only; thread 0 alone writes T_CURR and NSTEPS
    const bool done = ctl[CTL_DONE] != 0.0;
    const double tau = ctl[CTL_TAU];
    if (done) return;
    if (i == 0) {
        ctl[CTL_T_CURR] = ctl[CTL_T_CURR] + tau;
        ctl[CTL_NSTEPS] = ctl[CTL_NSTEPS] + 1.0;
    }
    if (i >= n) return;
    const real_t t = (real_t)tau;
    if (tstep) tstep[i] = t;
    if (time) time[i] = time[i] + t;
    if (nstep) nstep[i] = nstep[i] + 1;
}

// The same bookkeeping for a sub-system of the hierarchical SIA recursion, where the step of each
// level is known on the host (sia.py:1108-1113: slow.tstep[...] = tau; slow.time += tau; ...)
__global__ void __launch_bounds__(256) stamp_kernel(long long n, real_t* __restrict__ time,
                                                    abi_uint* __restrict__ nstep, real_t* __restrict__ tstep,
                                                    double tau)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const real_t t = (real_t)tau;
    if (tstep) tstep[i] = t;
    if (time) time[i] = time[i] + t;
    if (nstep) nstep[i] = nstep[i] + 1;
}

// ---------------------------------------------------------------------------------------
// Reductions over particles (deterministic: fixed grid, fixed tree; double accumulation).
// ---------------------------------------------------------------------------------------
enum { RED_SUM = 0, RED_KINETIC = 1, RED_HALF_DOT = 2, RED_SAKURA_DT = 3, RED_ABS_MIN = 4, RED_ABS_MAX = 5,
       RED_DOT = 6, RED_MOMENT = 7, RED_COUNT = 8 };

struct RedRefs {
    const real_t* x[5];
    long long n;
    double param;
};

TUPAN_DEV bool red_is_sum(int what) { return what <= RED_HALF_DOT || what >= RED_DOT; }

TUPAN_DEV double red_term(int what, const RedRefs& a, long long i)
{
    switch (what) {
        case RED_SUM: return (double)a.x[0][i];
        case RED_KINETIC: {   // 0.5 m (vx^2 + vy^2 + vz^2), body.py:262-275
            const real_t vx = a.x[1][i], vy = a.x[2][i], vz = a.x[3][i];
            return (double)((real_t)0.5 * a.x[0][i] * (vx * vx + vy * vy + vz * vz));
        }
        case RED_HALF_DOT: return (double)(a.x[0][i] * a.x[1][i]);   // m * phi, body.py:294-299
        case RED_SAKURA_DT: {  // (eta/tstep)^2 - (eta/tstepij)^2, sakura.py:105-108
            const real_t eta = (real_t)a.param;
            const real_t wa = eta / a.x[0][i], wb = eta / a.x[1][i];
            return (double)(wa * wa - wb * wb);
        }
        case RED_DOT: return (double)(a.x[0][i] * a.x[1][i]);        // m * r, m * v: body.py:60-72, 88-126
        case RED_MOMENT:   // m (ra vb - rb va): one component of the angular momentum, body.py:175-186
            return (double)(a.x[0][i] * (a.x[1][i] * a.x[4][i] - a.x[2][i] * a.x[3][i]));
        default: return fabs((double)a.x[0][i]);
    }
}
TUPAN_DEV double red_identity(int what)
{
    if (red_is_sum(what)) return 0.0;
    return what == RED_ABS_MIN ? (double)INFINITY : -(double)INFINITY;
}
TUPAN_DEV double red_op(int what, double p, double q)
{
    if (red_is_sum(what)) return p + q;
    return what == RED_ABS_MIN ? fmin(p, q) : fmax(p, q);
}

TUPAN_DEV double red_block(int what, double v, double* sm)
{
    for (int off = 16; off > 0; off >>= 1) v = red_op(what, v, __shfl_xor_sync(0xffffffffu, v, off));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : red_identity(what);
    if (threadIdx.x < 32)
        for (int off = 16; off > 0; off >>= 1) v = red_op(what, v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

__global__ void __launch_bounds__(256) reduce_stage1_kernel(int what, const RedRefs a, double* __restrict__ part)
{
    __shared__ double sm[8];
    double v = red_identity(what);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n;
         i += (long long)gridDim.x * blockDim.x)
        v = red_op(what, v, red_term(what, a, i));
    v = red_block(what, v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
}

__global__ void __launch_bounds__(256) reduce_stage2_kernel(int what, const double* __restrict__ part, int nparts,
                                                            double param, double* __restrict__ out)
{
    __shared__ double sm[8];
    double v = red_identity(what);
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) v = red_op(what, v, part[i]);
    v = red_block(what, v, sm);
    if (threadIdx.x == 0) {
        if (what == RED_HALF_DOT) v = 0.5 * v;
        if (what == RED_SAKURA_DT) {
            // dt_sakura = eta / (1 + max)^0.5 in REAL (sakura.py:109-110)
            // (a shard without particles reduces to the identity -inf: it must not turn the
            // all-reduced minimum of the ranks into NaN -> dt = eta, which never wins the minimum)
            const real_t eta = (real_t)param;
            v = nparts > 0 && v > -(double)INFINITY ? (double)(eta / sqrt((real_t)1 + (real_t)v)) : (double)eta;
        }
        *out = v;
    }
}

static DevBuf g_red_scratch;

}  // namespace tupan

using namespace tupan;

namespace {
inline unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

int begin_call(Context*& c)
{
    c = &ctx();
    return c->init();
}
}  // namespace

extern "C" {

int tupan_cuda_step_begin_dev(void* d_ctl, const void* d_min_tstep, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    step_begin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((double*)d_ctl, (const double*)d_min_tstep,
                                                        sizeof(real_t) == 4 ? 1 : 0);
    TUPAN_CHECK(cudaGetLastError(), "step_begin_kernel");
    c->launches++;
    return 0;
}

static int hermite_launch(bool predict, int order, long long n, void* const* rv, void* const* rv0,
                          const void* const* d0, const void* const* d1, const void* d_ctl, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    if (order != 2 && order != 4 && order != 6 && order != 8) return c->fail(cudaErrorInvalidValue, "hermite order");
    if (n <= 0) return 0;
    const int nd = order / 2;
    HermiteRefs a;
    for (int k = 0; k < 6; ++k) {
        a.rv[k] = (real_t*)rv[k];
        a.rv0[k] = (real_t*)rv0[k];
    }
    for (int k = 0; k < 12; ++k) {
        a.d0[k] = k < 3 * nd ? (const real_t*)d0[k] : nullptr;
        a.d1[k] = (d1 && k < 3 * nd) ? (const real_t*)d1[k] : nullptr;
    }
    a.ctl = (const double*)d_ctl;
    a.n = n;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = blocks_for(n);
    if (predict) {
        switch (nd) {
            case 1: hermite_predict_kernel<1><<<g, 256, 0, s>>>(a); break;
            case 2: hermite_predict_kernel<2><<<g, 256, 0, s>>>(a); break;
            case 3: hermite_predict_kernel<3><<<g, 256, 0, s>>>(a); break;
            default: hermite_predict_kernel<4><<<g, 256, 0, s>>>(a); break;
        }
    } else {
        switch (nd) {
            case 1: hermite_correct_kernel<1><<<g, 256, 0, s>>>(a); break;
            case 2: hermite_correct_kernel<2><<<g, 256, 0, s>>>(a); break;
            case 3: hermite_correct_kernel<3><<<g, 256, 0, s>>>(a); break;
            default: hermite_correct_kernel<4><<<g, 256, 0, s>>>(a); break;
        }
    }
    TUPAN_CHECK(cudaGetLastError(), predict ? "hermite_predict_kernel" : "hermite_correct_kernel");
    c->launches++;
    return 0;
}

int tupan_cuda_hermite_predict_dev(int order, long long n, void* const* rv, void* const* rv0, const void* const* d0,
                                   const void* d_ctl, void* stream)
{
    return hermite_launch(true, order, n, rv, rv0, d0, nullptr, d_ctl, stream);
}

int tupan_cuda_hermite_correct_dev(int order, long long n, void* const* rv, void* const* rv0, const void* const* d0,
                                   const void* const* d1, const void* d_ctl, void* stream)
{
    return hermite_launch(false, order, n, rv, rv0, d0, d1, d_ctl, stream);
}

int tupan_cuda_block_predict_dev(int order, long long n, const void* const* state, const void* time, double t_next,
                                 void* const* pred, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    if (order != 4 && order != 6) return c->fail(cudaErrorInvalidValue, "block_predict: order 4 or 6");
    if (n <= 0) return 0;
    const int nd = order / 2;
    BlockPredictRefs a;
    for (int k = 0; k < 15; ++k) a.x[k] = k < 3 * (nd + 2) ? (const real_t*)state[k] : nullptr;
    for (int k = 0; k < 12; ++k) a.p[k] = k < (nd >= 3 ? 12 : 6) ? (real_t*)pred[k] : nullptr;
    a.time = (const real_t*)time;
    a.t_next = t_next;
    a.n = n;
    cudaStream_t s = (cudaStream_t)stream;
    if (nd == 2) block_predict_kernel<2><<<blocks_for(n), 256, 0, s>>>(a);
    else block_predict_kernel<3><<<blocks_for(n), 256, 0, s>>>(a);
    TUPAN_CHECK(cudaGetLastError(), "block_predict_kernel");
    c->launches++;
    return 0;
}

int tupan_cuda_block_correct_dev(int order, long long n, const void* tau, const void* const* rv0, const void* const* d0,
                                 const void* const* d1, void* const* rv, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    if (order != 4 && order != 6) return c->fail(cudaErrorInvalidValue, "block_correct: order 4 or 6");
    if (n <= 0) return 0;
    const int nd = order / 2;
    BlockCorrectRefs a;
    a.tau = (const real_t*)tau;
    for (int k = 0; k < 6; ++k) {
        a.rv0[k] = (const real_t*)rv0[k];
        a.rv[k] = (real_t*)rv[k];
    }
    for (int k = 0; k < 9; ++k) {
        a.d0[k] = k < 3 * nd ? (const real_t*)d0[k] : nullptr;
        a.d1[k] = k < 3 * nd ? (const real_t*)d1[k] : nullptr;
    }
    a.n = n;
    cudaStream_t s = (cudaStream_t)stream;
    if (nd == 2) block_correct_kernel<2><<<blocks_for(n), 256, 0, s>>>(a);
    else block_correct_kernel<3><<<blocks_for(n), 256, 0, s>>>(a);
    TUPAN_CHECK(cudaGetLastError(), "block_correct_kernel");
    c->launches++;
    return 0;
}

int tupan_cuda_block_select_dev(long long n, const void* time, const void* dt, void* d_out2, void* d_idx, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    if (n <= 0) return c->fail(cudaErrorInvalidValue, "block_select: no particles");
    block_select_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n, (const real_t*)time, (const real_t*)dt,
                                                              (double*)d_out2, (long long*)d_idx);
    TUPAN_CHECK(cudaGetLastError(), "block_select_kernel");
    c->launches++;
    return 0;
}

int tupan_cuda_block_quantize_dev(long long n, const void* ts, const void* tau, double t_next, double dt_max,
                                  void* dt_new, void* time_new, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    if (n <= 0) return 0;
    block_quantize_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(n, (const real_t*)ts, (const real_t*)tau,
                                                                          t_next, dt_max, (real_t*)dt_new,
                                                                          (real_t*)time_new);
    TUPAN_CHECK(cudaGetLastError(), "block_quantize_kernel");
    c->launches++;
    return 0;
}

int tupan_cuda_axpy_dev(int narr, long long n, void* const* y, const void* const* x, double c_outer, double c_inner,
                        const void* d_ctl, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    if (narr < 1 || narr > 6) return c->fail(cudaErrorInvalidValue, "axpy: 1..6 arrays");
    if (n <= 0) return 0;
    VecRefs a;
    for (int k = 0; k < 6; ++k) {
        a.y[k] = k < narr ? (real_t*)y[k] : nullptr;
        a.x[k] = k < narr ? (const real_t*)x[k] : nullptr;
    }
    a.narr = narr;
    a.n = n;
    axpy_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(a, c_outer, c_inner, (const double*)d_ctl);
    TUPAN_CHECK(cudaGetLastError(), "axpy_kernel");
    c->launches++;
    return 0;
}

int tupan_cuda_scale_dev(int narr, long long n, void* const* y, const void* const* x, double denom, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    if (narr < 1 || narr > 6) return c->fail(cudaErrorInvalidValue, "scale: 1..6 arrays");
    if (n <= 0) return 0;
    VecRefs a;
    for (int k = 0; k < 6; ++k) {
        a.y[k] = k < narr ? (real_t*)y[k] : nullptr;
        a.x[k] = k < narr ? (const real_t*)x[k] : nullptr;
    }
    a.narr = narr;
    a.n = n;
    scale_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(a, denom);
    TUPAN_CHECK(cudaGetLastError(), "scale_kernel");
    c->launches++;
    return 0;
}

int tupan_cuda_step_end_dev(long long n, void* d_time, void* d_nstep, void* d_tstep, void* d_ctl, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    const long long m = n > 0 ? n : 1;
    step_end_kernel<<<blocks_for(m), 256, 0, (cudaStream_t)stream>>>(n, (real_t*)d_time, (abi_uint*)d_nstep,
                                                                     (real_t*)d_tstep, (double*)d_ctl);
    TUPAN_CHECK(cudaGetLastError(), "step_end_kernel");
    c->launches++;
    return 0;
}

int tupan_cuda_pn_kick_dev(int phase, long long n, void* const* arr, double c_outer, double c_inner,
                           const void* d_ctl, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    if (phase != 0 && phase != 1) return c->fail(cudaErrorInvalidValue, "pn_kick phase");
    if (n <= 0) return 0;
    PnRefs p;
    for (int k = 0; k < 3; ++k) {
        p.v[k] = (real_t*)arr[k];
        p.a[k] = (const real_t*)arr[3 + k];
        p.w[k] = (real_t*)arr[6 + k];
        p.pna[k] = (const real_t*)arr[9 + k];
        p.r[k] = (const real_t*)arr[13 + k];
        p.pn_mv[k] = (real_t*)arr[17 + k];
        p.pn_am[k] = (real_t*)arr[20 + k];
    }
    p.mass = (const real_t*)arr[12];
    p.pn_ke = (real_t*)arr[16];
    p.n = n;
    cudaStream_t s = (cudaStream_t)stream;
    if (phase == 0)
        pn_kick_kernel<0><<<blocks_for(n), 256, 0, s>>>(p, c_outer, c_inner, (const double*)d_ctl);
    else
        pn_kick_kernel<1><<<blocks_for(n), 256, 0, s>>>(p, c_outer, c_inner, (const double*)d_ctl);
    TUPAN_CHECK(cudaGetLastError(), "pn_kick_kernel");
    c->launches++;
    return 0;
}

int tupan_cuda_stamp_dev(long long n, void* d_time, void* d_nstep, void* d_tstep, double tau, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    if (n <= 0) return 0;
    stamp_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(n, (real_t*)d_time, (abi_uint*)d_nstep,
                                                                  (real_t*)d_tstep, tau);
    TUPAN_CHECK(cudaGetLastError(), "stamp_kernel");
    c->launches++;
    return 0;
}

int tupan_cuda_reduce_dev(int what, long long n, const void* const* arrays, double param, void* d_out, void* stream)
{
    Context* c;
    std::lock_guard<std::mutex> lock(ctx().mu);
    int rc = begin_call(c);
    if (rc) return rc;
    if (what < RED_SUM || what >= RED_COUNT) return c->fail(cudaErrorInvalidValue, "reduce: unknown reduction");
    static const int nin[] = {1, 4, 2, 2, 1, 1, 2, 5};
    RedRefs a;
    for (int k = 0; k < 5; ++k) a.x[k] = k < nin[what] ? (const real_t*)arrays[k] : nullptr;
    a.n = n > 0 ? n : 0;
    a.param = param;
    int parts = (int)((a.n + 1023) / 1024);
    if (parts > 4 * c->info.sm_count) parts = 4 * c->info.sm_count;
    if (parts < 1) parts = 1;
    double* scratch = static_cast<double*>(g_red_scratch.ensure(sizeof(double) * 4 * 1024));
    if (!scratch) return c->fail(cudaErrorMemoryAllocation, "reduce scratch");
    cudaStream_t s = (cudaStream_t)stream;
    reduce_stage1_kernel<<<parts, 256, 0, s>>>(what, a, scratch);
    TUPAN_CHECK(cudaGetLastError(), "reduce_stage1_kernel");
    reduce_stage2_kernel<<<1, 256, 0, s>>>(what, scratch, parts, param, (double*)d_out);
    TUPAN_CHECK(cudaGetLastError(), "reduce_stage2_kernel");
    c->launches += 2;
    return 0;
}

}  // extern "C"
