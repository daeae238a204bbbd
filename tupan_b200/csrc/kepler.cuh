// kepler.cuh -- universal-variable two-body propagator on the device, and the Sakura pair op.
//
// Replaces, for the GPU path:
//   universal_kepler_solver.h:17-614   Stumpff functions, Kepler's equation in universal
//                                      variables, Laguerre (order 5) root find, Lagrange f/g
//                                      map, sub-stepping on failure, softened-orbit energy check
//   sakura_kernel_common.h:8-243       leapfrog / Kepler choice per pair, drift removal by flag
//   kepler_solver_kernel_common.h:8-81 two-body wrapper in the centre-of-mass frame
//
// Behaviour that has to match the reference: tolerance 2^-42 (fp64) / 2^-16 (fp32) on the
// Laguerre step, at most 64 iterations, failure codes trigger a restart with twice as many
// sub-steps, period reduction for bound orbits, log-based first guess for hyperbolic ones,
// r2 > 0 mask returns the state unchanged.
//
// Sub-step doubling.  The reference doubles the number of sub-steps without bound, both on a
// solver failure (:481-520) and in the energy check of softened orbits (:523-606); for SOFTENED
// tight binaries the latter needs 2^14 .. 2^27 SEQUENTIAL sub-steps in the reference's own golden
// vectors (tests/golden/kepler_fp64.npz; 43 s on a host core for the worst).  Here the doubling
// takes a bound as an argument:
//   * inside the pair sweep (sakura) it is 2^SWEEP_DOUBLINGS, so that one pathological pair cannot
//     hold 255 other threads of its CTA for minutes; a pair that runs into it is not failed but
//     handed, with its description, to a clean-up launch (one thread per such pair, see
//     sakura_cleanup_kernel) whose result is added to the owner's outputs afterwards;
//   * the clean-up launch and the two-body entry point use 2^FULL_DOUBLINGS = 2^27 sub-steps -- the
//     deepest case of the reference's golden vectors, 43 s on a host core -- i.e. the reference's
//     behaviour for anything it finishes within minutes.  Only beyond that a pair is counted in
//     kepler_limit_hits and the host entry points fail loudly (tupan_cuda_last_error).
#pragma once
#include "ops.cuh"

namespace tupan {

TUPAN_DEV double k_sqrt(double x) { return sqrt(x); }
TUPAN_DEV float k_sqrt(float x) { return sqrtf(x); }
TUPAN_DEV double k_abs(double x) { return fabs(x); }
TUPAN_DEV float k_abs(float x) { return fabsf(x); }
TUPAN_DEV double k_cos(double x) { return cos(x); }
TUPAN_DEV float k_cos(float x) { return cosf(x); }
TUPAN_DEV double k_sin(double x) { return sin(x); }
TUPAN_DEV float k_sin(float x) { return sinf(x); }
TUPAN_DEV double k_cosh(double x) { return cosh(x); }
TUPAN_DEV float k_cosh(float x) { return coshf(x); }
TUPAN_DEV double k_sinh(double x) { return sinh(x); }
TUPAN_DEV float k_sinh(float x) { return sinhf(x); }
TUPAN_DEV double k_log(double x) { return log(x); }
TUPAN_DEV float k_log(float x) { return logf(x); }

template <typename T> struct KeplerTol;
template <> struct KeplerTol<double> { static TUPAN_DEV double value() { return 2.2737367544323205948e-13; } };  // 2^-42
template <> struct KeplerTol<float>  { static TUPAN_DEV float value() { return 1.52587890625e-5f; } };           // 2^-16

// SOLVER_DOUBLINGS bounds the restart-on-solver-failure loop (:481-520), which in practice ends after
// 0-3 doublings; it is what keeps a pair with a state the solver cannot digest from spinning.
enum { KEPLER_MAXITER = 64, SWEEP_DOUBLINGS = 12, FULL_DOUBLINGS = 27, SOLVER_DOUBLINGS = 16 };

__device__ unsigned int kepler_limit_hits = 0;   // pairs that ran into FULL_DOUBLINGS (this TU only)
__device__ unsigned long long kepler_cleanup_total = 0;   // pairs the sweeps handed to the clean-up launch so far

template <typename T> TUPAN_DEV int sgn(T x) { return (x > T(0)) - (x < T(0)); }

template <typename T> struct State { T x, y, z, vx, vy, vz; };

// Stumpff-type functions of z = alpha s^2.  c0..c3 are evaluated together where a caller
// needs several of them (sin/cos or sinh/cosh of the same argument).
template <typename T> struct Stumpff { T c0, c1, c2, c3; };

template <typename T> TUPAN_DEV Stumpff<T> stumpff_plain(T z)
{
    Stumpff<T> o;
    if (z < T(0)) {
        T s = k_sqrt(-z);
        T cs = k_cos(s), sn = k_sin(s);
        o.c0 = cs;
        o.c1 = sn / s;
        o.c2 = (cs - T(1)) / z;
        o.c3 = (sn / s - T(1)) / z;
    } else if (z > T(0)) {
        T s = k_sqrt(z);
        T ch = k_cosh(s), sh = k_sinh(s);
        o.c0 = ch;
        o.c1 = sh / s;
        o.c2 = (ch - T(1)) / z;
        o.c3 = (sh / s - T(1)) / z;
    } else {
        o.c0 = T(1);
        o.c1 = T(1);
        o.c2 = T(1) / T(2);
        o.c3 = T(1) / T(6);
    }
    return o;
}

template <typename T> struct KeplerEq { T dt, r0, rv0, m, alpha; };

// fp64: the closed forms, exactly as the reference writes them (universal_kepler_solver.h:17-83).
// fp32: (cos s - 1)/z and (sin s / s - 1)/z cancel catastrophically for the small arguments the
// sub-stepping produces -- one ulp of cosf is 1e-3 of c2 at s = 0.01 -- which is the size of the
// reference's own energy-check tolerance (64 x 2^-16): whether a softened orbit passes the check
// then depends on the last bit of the libm in use (glibc's cosf on the host, CUDA's on the
// device), and with CUDA's the doubling did not terminate for 9 of the 18 golden cases.  The fp32
// library therefore evaluates the four functions in double and rounds once: the value the
// reference's formula means, to float accuracy.
TUPAN_DEV Stumpff<double> stumpff(double z) { return stumpff_plain<double>(z); }
TUPAN_DEV Stumpff<float> stumpff(float z)
{
    const Stumpff<double> d = stumpff_plain<double>((double)z);
    Stumpff<float> o;
    o.c0 = (float)d.c0; o.c1 = (float)d.c1; o.c2 = (float)d.c2; o.c3 = (float)d.c3;
    return o;
}

// Laguerre iteration, order 5 (universal_kepler_solver.h:219-254).
template <typename T> TUPAN_DEV int laguerre5(T x0, T& x, const KeplerEq<T>& e)
{
    const T tol = KeplerTol<T>::value();
    const T mar = e.m + e.alpha * e.r0;
    int it = 0;
    T delta;
    x = x0;
    do {
        const T s = x, s2 = s * s;
        const Stumpff<T> c = stumpff(e.alpha * s2);
        const T S1 = s * c.c1, S2 = s2 * c.c2, S3 = (s * s2) * c.c3;
        const T fv = (e.r0 * s + e.rv0 * S2 + mar * S3) - e.dt;
        const T dfv = e.r0 + e.rv0 * S1 + mar * S2;
        const T ddfv = e.rv0 * c.c0 + mar * S1;
        const T a = dfv, a2 = a * a;
        const T b = a2 - fv * ddfv;
        const T g = T(5) * fv;
        const T h = a + T(sgn(a)) * k_sqrt(k_abs(T(4) * (T(5) * b - a2)));
        if (h == T(0)) return -1;
        delta = -g / h;
        x += delta;
        it += 1;
        if (it > KEPLER_MAXITER) return -2;
    } while (k_abs(delta) > tol);
    if (sgn(x) != sgn(x0)) return -3;
    return 0;
}

// One attempt over dt0 (universal_kepler_solver.h:368-478); the state moves only on success.
template <typename T> TUPAN_DEV int kepler_try(T dt0, T m, T e2, State<T>& p)
{
    const T PI = T(3.141592653589793);
    T r2 = p.x * p.x + p.y * p.y + p.z * p.z;
    if (!(r2 > T(0))) return 0;
    r2 += e2;
    const T r = k_sqrt(r2);
    const T v2 = p.vx * p.vx + p.vy * p.vy + p.vz * p.vz;
    const T rv = p.x * p.vx + p.y * p.vy + p.z * p.vz;
    const T beta = T(2) - (e2 / r2);
    const T alpha = v2 - beta * m / r;

    T dt = dt0;
    if (alpha < T(0)) {  // bound orbit: remove whole periods
        const T a = m / k_abs(alpha);
        const T P = T(2) * PI * a * k_sqrt(a / m);
        const T ratio = dt0 / P;
        dt = (ratio - T((long long)ratio)) * P;
    }
    T s0 = dt / r;
    if (alpha > T(0)) {  // hyperbolic first guess (Bate et al. 1971, 4.5.11)
        const T sa = k_sqrt(alpha);
        const T ss = k_abs(T(2) * alpha * dt / (rv + (m + alpha * r) / sa));
        if (ss > T(1)) s0 = T(sgn(dt)) * k_log(ss) / sa;
    }
    KeplerEq<T> e = {dt, r, rv, m, alpha};
    T s;
    const int err = laguerre5(s0, s, e);
    if (err != 0) return err;

    // Lagrange coefficients (universal_kepler_solver.h:317-365, '+e2' under the sqrt :347)
    const T s2 = s * s;
    const Stumpff<T> c = stumpff(alpha * s2);
    const T S1 = s * c.c1, S2 = s2 * c.c2, S3 = (s * s2) * c.c3;
    const T lf = T(1) - m * S2 / r;
    const T lg = dt - m * S3;
    const T x1 = p.x * lf + p.vx * lg;
    const T y1 = p.y * lf + p.vy * lg;
    const T z1 = p.z * lf + p.vz * lg;
    const T r1 = k_sqrt(x1 * x1 + y1 * y1 + z1 * z1 + e2);
    const T ldf = -m * S1 / (r * r1);
    const T ldg = (T(1) + lg * ldf) / lf;
    const T u1 = p.x * ldf + p.vx * ldg;
    const T v1 = p.y * ldf + p.vy * ldg;
    const T w1 = p.z * ldf + p.vz * ldg;
    p.x = x1; p.y = y1; p.z = z1;
    p.vx = u1; p.vy = v1; p.vz = w1;
    return 0;
}

// Restart with twice as many equal sub-steps on any failure (:481-520).  `ok` is cleared when
// 2^max_doublings sub-steps were not enough.
template <typename T> TUPAN_DEV State<T> kepler_substep(T dt, T m, T e2, const State<T>& p0, int max_doublings, bool& ok)
{
    State<T> p = p0;
    int n = 1;
    bool bad = false;
    if (max_doublings > SOLVER_DOUBLINGS) max_doublings = SOLVER_DOUBLINGS;
    for (int level = 0; level <= max_doublings; ++level) {
        bad = false;
        p = p0;
        const T h = dt / T(n);
        for (int i = 0; i < n; ++i) {
            if (kepler_try(h, m, e2, p) != 0) { bad = true; break; }
        }
        if (!bad) break;
        n *= 2;
    }
    if (bad) ok = false;
    return p;
}

template <typename T> TUPAN_DEV T kepler_energy(T m, T e2, const State<T>& p, T& u)
{
    const T r = k_sqrt(p.x * p.x + p.y * p.y + p.z * p.z + e2);
    const T v2 = p.vx * p.vx + p.vy * p.vy + p.vz * p.vz;
    u = T(2) * m / r;
    return v2 - u;
}

// Driver with the energy check for softened orbits (:523-614).  Returns false when the bound on
// the doubling was hit (the state returned is then meaningless).
template <typename T>
__device__ __noinline__ bool kepler_propagate(T dt, T m, T e2, const State<T>& p0, int max_doublings, State<T>& out)
{
    bool ok = true;
    State<T> p = kepler_substep(dt, m, e2, p0, max_doublings, ok);
    out = p;
    if (!ok) return false;
    if (e2 == T(0)) return true;
    const T r2 = p0.x * p0.x + p0.y * p0.y + p0.z * p0.z;
    if (!(r2 > T(0))) { out = p0; return true; }
    T u0, u1;
    const T e0 = kepler_energy(m, e2, p0, u0);
    T e1 = kepler_energy(m, e2, p, u1);
    const T tol = T(64) * KeplerTol<T>::value();
    if (T(2) * k_abs(e1 - e0) < tol * (u1 + u0)) return true;
    int n = 1;
    bool bad = true;
    for (int level = 0; level < max_doublings; ++level) {
        n *= 2;
        bad = false;
        p = p0;
        const T h = dt / T(n);
        for (int i = 0; i < n; ++i) {
            p = kepler_substep(h, m, e2, p, max_doublings, ok);
            if (!ok) return false;
            e1 = kepler_energy(m, e2, p, u1);
            if (T(2) * k_abs(e1 - e0) > tol * (u1 + u0)) { bad = true; break; }
        }
        if (!bad) break;
    }
    out = p;
    return !bad;
}

// Leapfrog (drift-kick-drift) with the softened pair force (sakura_kernel_common.h:45-91).
template <typename T> TUPAN_DEV void twobody_leapfrog(T dt, T m, T e2, State<T>& p)
{
    const T half = dt / T(2);
    p.x = fma(p.vx, half, p.x); p.y = fma(p.vy, half, p.y); p.z = fma(p.vz, half, p.z);
    T r2 = p.x * p.x; r2 = fma(p.y, p.y, r2); r2 = fma(p.z, p.z, r2);
    const InvR<T> w = soft_inv<true>(r2 + e2, r2);
    const T g = -(m * w.r3) * dt;
    p.vx = fma(g, p.x, p.vx); p.vy = fma(g, p.y, p.vy); p.vz = fma(g, p.z, p.vz);
    p.x = fma(p.vx, half, p.x); p.y = fma(p.vy, half, p.y); p.z = fma(p.vz, half, p.z);
}

// Wide or fast pairs take the leapfrog, the rest the Kepler map (:94-123):
//     R = 64 (m / v2);  r2 > R^2 ? leapfrog : universal_kepler_solver
// FAST = true is the in-line half of the deferring sweep (pair_engine.cuh).  It takes the
// leapfrog only when the choice is beyond doubt -- r2 v2^2 > 4096 m^2 (1 + 2^-36) (fp32: 2^-18), a
// test without the division, whose margin covers any rounding of either form --, leaves a masked
// pair (r2 > 0 false) unchanged as kepler_propagate would, and returns false for everything else; the caller parks such a pair and re-runs it from its initial state with
// FAST = false, i.e. with the reference's own test.  Parking is always safe.
// FAST = false: returns false when the Kepler sub-stepping ran into its bound (2^max_doublings).
template <bool FAST, typename T> TUPAN_DEV bool twobody_step(T dt, T m, T e2, State<T>& p, int max_doublings = 0)
{
    const T r2 = p.x * p.x + p.y * p.y + p.z * p.z;
    const T v2 = p.vx * p.vx + p.vy * p.vy + p.vz * p.vz;
    if (FAST) {
        const T margin = sizeof(T) == 8 ? T(1.4551915228366852e-11) : T(3.814697265625e-6);  // 2^-36, 2^-18
        const T lhs = r2 * (v2 * v2);
        const T rhs = (m * m) * (T(4096) * (T(1) + margin));
        if (lhs > rhs) {
            twobody_leapfrog(dt, m, e2, p);
            return true;
        }
        return !(r2 > T(0));   // masked pair: unchanged, as kepler_propagate would return it
    }
    const T R = T(64) * (m / v2);
    if (r2 > R * R) {
        twobody_leapfrog(dt, m, e2, p);
        return true;
    }
    State<T> q;
    const bool ok = kepler_propagate(dt, m, e2, p, max_doublings, q);
    p = q;
    return ok;
}

// flag in {-2,-1,1,2}: where the free drift is taken out (:126-191); other values: no-op.
template <bool FAST, typename T> TUPAN_DEV bool twobody_evolve(T dt, int flag, T m, T e2, State<T>& p, int md = 0)
{
    if (flag == -1) {
        p.x -= p.vx * dt; p.y -= p.vy * dt; p.z -= p.vz * dt;
        return twobody_step<FAST>(dt, m, e2, p, md);
    } else if (flag == 1) {
        if (!twobody_step<FAST>(dt, m, e2, p, md)) return false;
        p.x -= p.vx * dt; p.y -= p.vy * dt; p.z -= p.vz * dt;
    } else if (flag == -2) {
        const T h = dt / T(2);
        p.x -= p.vx * h; p.y -= p.vy * h; p.z -= p.vz * h;
        if (!twobody_step<FAST>(dt, m, e2, p, md)) return false;
        p.x -= p.vx * h; p.y -= p.vy * h; p.z -= p.vz * h;
    } else if (flag == 2) {
        if (!twobody_step<FAST>(dt / T(2), m, e2, p, md)) return false;
        p.x -= p.vx * dt; p.y -= p.vy * dt; p.z -= p.vz * dt;
        return twobody_step<FAST>(dt / T(2), m, e2, p, md);
    }
    return true;
}

// =======================================================================================
// sakura -- replaces sakura_kernel (sakura_kernel.c:5-64, core sakura_kernel_common.h:194-243)
// =======================================================================================
// jobs: pairs whose Kepler sub-stepping hit the sweep's bound; each entry is JOB_REALS reals --
// the pair description d[ND], the owner's particle index (as a real) -- filled by the sweep,
// replaced by the pair's contribution c[NA] by sakura_cleanup_kernel, applied by sakura_apply_kernel.
template <typename T> struct SakuraParams { T dt; int flag; T* jobs; unsigned* njobs; unsigned jobs_cap; };
enum { SAKURA_JOB_REALS = 12 };
// FLAG is a template parameter: with the flag tested at run time inside the pair, the four
// variants met at a join point and every pair paid ~40 register moves and 6 branches for it
// (profiles/r01_sakura_defer_ncu_summary.txt: 139 instructions per pair, 66 of them FP64).
// FLAG = 0 stands for every other value (no-op, as in the reference).
template <typename T, int FLAG> __global__ void sakura_cleanup_kernel(SakuraParams<T> p);
template <typename T> __global__ void sakura_apply_kernel(SakuraParams<T> p, OutRefs<T> out, long long ni);

template <typename T, int FLAG> struct SakuraOp {
    typedef T real;
    typedef SakuraParams<T> Params;
    enum { NI = 8, NJ = 8, NA = 6, NO = 6, WPT = 1, UNROLL = 1 };
    enum { NJP = round_up(NJ, Vec16<T>::N) };
    enum { IX, IY, IZ, IE, IVX, IVY, IVZ, IM };
    static TUPAN_DEV void load_i(const T* const* a, long long i, T (&s)[NI])
    {
        s[IM] = a[0][i];
        s[IX] = a[1][i]; s[IY] = a[2][i]; s[IZ] = a[3][i]; s[IE] = a[4][i];
        s[IVX] = a[5][i]; s[IVY] = a[6][i]; s[IVZ] = a[7][i];
    }
    static TUPAN_DEV void pack_j(const T* const* j, long long r, T (&row)[NJP]) { pack_row8(j, r, row); }
    static TUPAN_DEV void zero(T (&a)[NA]) { zero_all(a); }
    // Deferring sweep (pair_engine.cuh): leapfrog pairs are finished in line, pairs that need
    // the Kepler solver are described by D_* and solved later with full warps.
    enum { D_X, D_Y, D_Z, D_VX, D_VY, D_VZ, D_M, D_E2, D_MJ, ND };
    static_assert(ND + 1 <= SAKURA_JOB_REALS, "job entry");
    static TUPAN_DEV bool pair_fast(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA], const Params& p, T (&d)[ND])
    {
        State<T> p0;
        p0.x = s[IX] - row[JX]; p0.y = s[IY] - row[JY]; p0.z = s[IZ] - row[JZ];
        p0.vx = s[IVX] - row[J8_VX]; p0.vy = s[IVY] - row[J8_VY]; p0.vz = s[IVZ] - row[J8_VZ];
        const T e2 = s[IE] + row[J8_E2];
        const T m = s[IM] + row[JM];
        // d is written on BOTH paths: left untouched on the common one it is a loop-carried value for
        // the compiler, and every pair paid 36 register moves to keep the previous (dead) copy alive
        d[D_X] = p0.x; d[D_Y] = p0.y; d[D_Z] = p0.z; d[D_VX] = p0.vx; d[D_VY] = p0.vy; d[D_VZ] = p0.vz;
        d[D_M] = m; d[D_E2] = e2; d[D_MJ] = row[JM];
        State<T> q = p0;
        if (twobody_evolve<true>(p.dt, FLAG, m, e2, q)) {
            const T mu = row[JM] * rcp_fast(m);
            a[0] = fma(mu, q.x - p0.x, a[0]); a[1] = fma(mu, q.y - p0.y, a[1]); a[2] = fma(mu, q.z - p0.z, a[2]);
            a[3] = fma(mu, q.vx - p0.vx, a[3]); a[4] = fma(mu, q.vy - p0.vy, a[4]); a[5] = fma(mu, q.vz - p0.vz, a[5]);
            return false;
        }
        return true;
    }
    // false: the sub-stepping hit 2^max_doublings; c is zero and the pair has to be redone
    static TUPAN_DEV bool pair_slow(const T (&d)[ND], const Params& p, T (&c)[NA], int max_doublings = SWEEP_DOUBLINGS)
    {
        const State<T> p0 = {d[D_X], d[D_Y], d[D_Z], d[D_VX], d[D_VY], d[D_VZ]};
        State<T> q = p0;
        const bool ok = twobody_evolve<false>(p.dt, FLAG, d[D_M], d[D_E2], q, max_doublings);
        const T mu = ok ? d[D_MJ] / d[D_M] : T(0);
        c[0] = mu * (q.x - p0.x); c[1] = mu * (q.y - p0.y); c[2] = mu * (q.z - p0.z);
        c[3] = mu * (q.vx - p0.vx); c[4] = mu * (q.vy - p0.vy); c[5] = mu * (q.vz - p0.vz);
        if (!ok) {
#pragma unroll
            for (int k = 0; k < NA; ++k) c[k] = T(0);
        }
        return ok;
    }
    // hand a pair the sweep could not finish to the clean-up launch
    static TUPAN_DEV void defer_to_cleanup(const T (&d)[ND], long long i, const Params& p)
    {
        const unsigned k = atomicAdd(p.njobs, 1u);
        if (k >= p.jobs_cap) { atomicAdd(&kepler_limit_hits, 1u); return; }   // list full: fail loudly
        T* e = p.jobs + (size_t)k * SAKURA_JOB_REALS;
#pragma unroll
        for (int q = 0; q < ND; ++q) e[q] = d[q];
        e[ND] = (T)i;
    }
    static TUPAN_DEV void combine(T (&a)[NA], const T (&b)[NA]) { sum_combine(a, b); }
    static TUPAN_DEV void finish(const T* const*, long long i, const T (&a)[NA], const Params&, T* const* out)
    {
#pragma unroll
        for (int k = 0; k < NO; ++k) out[k][i] = a[k];
    }
    // Host side, after the outputs of a call have been written (pair kernel or finalize): finish
    // the pairs the sweep handed over and add them to the owners' outputs.  Two small launches,
    // stream-ordered, no host round trip; they find an empty list in the normal case.
    static cudaError_t after_sweeps(long long ni, const Params& p, const OutRefs<T>& out, cudaStream_t st, long long* launches)
    {
        if (FLAG == 0 || ni <= 0 || !p.jobs) return cudaSuccess;
        sakura_cleanup_kernel<T, FLAG><<<64, 32, 0, st>>>(p);
        sakura_apply_kernel<T><<<1, 32, 0, st>>>(p, out, ni);
        if (launches) *launches += 2;
        return cudaGetLastError();
    }
};

template <typename T, int FLAG> struct Defers<SakuraOp<T, FLAG>> { enum { value = 1 }; };
template <typename T, int FLAG> struct OpCost<SakuraOp<T, FLAG>> { enum { value = (FLAG == 2 || FLAG == -2) ? 90 : 62 }; };

// The pairs the sweep handed over: one thread each, the reference's unbounded sub-stepping
// (2^FULL_DOUBLINGS).  The entry's description is replaced by the pair's contribution.
template <typename T, int FLAG>
__global__ void sakura_cleanup_kernel(SakuraParams<T> p)
{
    typedef SakuraOp<T, FLAG> Op;
    unsigned n = *p.njobs;
    if (n > p.jobs_cap) n = p.jobs_cap;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        T* e = p.jobs + (size_t)k * SAKURA_JOB_REALS;
        T d[Op::ND], c[Op::NA];
#pragma unroll
        for (int q = 0; q < Op::ND; ++q) d[q] = e[q];
        if (!Op::pair_slow(d, p, c, FULL_DOUBLINGS)) atomicAdd(&kepler_limit_hits, 1u);
#pragma unroll
        for (int q = 0; q < Op::NA; ++q) e[q] = c[q];
    }
}
// Add the contributions to the owners' outputs in a fixed order (by owner, then by value), so that
// the result does not depend on the order in which the sweep appended them; resets the list.
template <typename T>
__global__ void sakura_apply_kernel(SakuraParams<T> p, OutRefs<T> out, long long ni)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned n = *p.njobs;
    if (n > p.jobs_cap) n = p.jobs_cap;
    enum { OWNER = 9 };                       // SakuraOp::ND
    for (unsigned done = 0; done < n; ++done) {
        // selection: the smallest (owner, first component) not applied yet; applied entries get owner -1
        int best = -1;
        for (unsigned k = 0; k < n; ++k) {
            const T* e = p.jobs + (size_t)k * SAKURA_JOB_REALS;
            if (e[OWNER] < T(0)) continue;
            if (best < 0) { best = (int)k; continue; }
            const T* b = p.jobs + (size_t)best * SAKURA_JOB_REALS;
            if (e[OWNER] < b[OWNER] || (e[OWNER] == b[OWNER] && e[0] < b[0])) best = (int)k;
        }
        if (best < 0) break;
        T* e = p.jobs + (size_t)best * SAKURA_JOB_REALS;
        const long long i = (long long)e[OWNER];
        if (i >= 0 && i < ni) {
#pragma unroll
            for (int q = 0; q < 6; ++q) out.p[q][i] += e[q];
        }
        e[OWNER] = T(-1);
    }
    kepler_cleanup_total += n;
    *p.njobs = 0u;
}

// =======================================================================================
// Two-body Kepler kernel -- replaces kepler_solver_kernel (kepler_solver_kernel.c:5-50).
// `pairs` independent binaries per launch (the reference ABI is pairs == 1); outputs may
// alias inputs (extensions.py:642-646): each thread reads its two bodies before writing.
// Arrays hold 2*pairs bodies, binary b = elements (2b, 2b+1).
// =======================================================================================
template <typename T>
__global__ void kepler_pairs_kernel(InRefs<T> in, long long pairs, T dt, OutRefs<T> out)
{
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= pairs) return;
    const long long i0 = 2 * b, i1 = 2 * b + 1;
    const T m0 = in.p[0][i0], m1 = in.p[0][i1];
    const T x0 = in.p[1][i0], y0 = in.p[2][i0], z0 = in.p[3][i0];
    const T x1 = in.p[1][i1], y1 = in.p[2][i1], z1 = in.p[3][i1];
    const T e2 = in.p[4][i0] + in.p[4][i1];
    const T u0 = in.p[5][i0], v0 = in.p[6][i0], w0 = in.p[7][i0];
    const T u1 = in.p[5][i1], v1 = in.p[6][i1], w1 = in.p[7][i1];
    State<T> rel = {x0 - x1, y0 - y1, z0 - z1, u0 - u1, v0 - v1, w0 - w1};
    const T m = m0 + m1;
    const T imu = m0 / m, jmu = m1 / m;
    T cx = imu * x0 + jmu * x1, cy = imu * y0 + jmu * y1, cz = imu * z0 + jmu * z1;
    const T cu = imu * u0 + jmu * u1, cv = imu * v0 + jmu * v1, cw = imu * w0 + jmu * w1;
    cx += cu * dt; cy += cv * dt; cz += cw * dt;
    State<T> q;
    if (!kepler_propagate(dt, m, e2, rel, FULL_DOUBLINGS, q)) atomicAdd(&kepler_limit_hits, 1u);
    out.p[0][i0] = cx + jmu * q.x;  out.p[1][i0] = cy + jmu * q.y;  out.p[2][i0] = cz + jmu * q.z;
    out.p[3][i0] = cu + jmu * q.vx; out.p[4][i0] = cv + jmu * q.vy; out.p[5][i0] = cw + jmu * q.vz;
    out.p[0][i1] = cx - imu * q.x;  out.p[1][i1] = cy - imu * q.y;  out.p[2][i1] = cz - imu * q.z;
    out.p[3][i1] = cu - imu * q.vx; out.p[4][i1] = cv - imu * q.vy; out.p[5][i1] = cw - imu * q.vz;
}

}  // namespace tupan
