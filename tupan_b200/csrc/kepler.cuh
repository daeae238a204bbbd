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
// r2 > 0 mask returns the state unchanged.  One deliberate difference: the two doubling
// loops are bounded (at most 2^MAX_DOUBLINGS sub-steps) so a pathological pair cannot hang
// the GPU.  The reference doubles without bound: for SOFTENED tight binaries its energy
// check (:533-606) drives it to 10^6 .. 10^9 sequential sub-steps (measured with the oracle:
// eps2 = 1e-8, a ~ 1e-3, dt = 0.37 -> n = 2^30, 6 minutes on a host core).  When the bound is
// hit the pair is counted in kepler_limit_hits and the host entry points fail loudly
// (tupan_cuda_last_error) instead of returning an unconverged state silently.
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

enum { KEPLER_MAXITER = 64, MAX_DOUBLINGS = 16 };

__device__ unsigned int kepler_limit_hits = 0;   // pairs that ran into MAX_DOUBLINGS (this TU only)

template <typename T> TUPAN_DEV int sgn(T x) { return (x > T(0)) - (x < T(0)); }

template <typename T> struct State { T x, y, z, vx, vy, vz; };

// Stumpff-type functions of z = alpha s^2.  c0..c3 are evaluated together where a caller
// needs several of them (sin/cos or sinh/cosh of the same argument).
template <typename T> struct Stumpff { T c0, c1, c2, c3; };

template <typename T> TUPAN_DEV Stumpff<T> stumpff(T z)
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

// Restart with twice as many equal sub-steps on any failure (:481-520).
template <typename T> TUPAN_DEV State<T> kepler_substep(T dt, T m, T e2, const State<T>& p0)
{
    State<T> p = p0;
    int n = 1;
    bool bad = false;
    for (int level = 0; level <= MAX_DOUBLINGS; ++level) {
        bad = false;
        p = p0;
        const T h = dt / T(n);
        for (int i = 0; i < n; ++i) {
            if (kepler_try(h, m, e2, p) != 0) { bad = true; break; }
        }
        if (!bad) break;
        n *= 2;
    }
    if (bad) atomicAdd(&kepler_limit_hits, 1u);
    return p;
}

template <typename T> TUPAN_DEV T kepler_energy(T m, T e2, const State<T>& p, T& u)
{
    const T r = k_sqrt(p.x * p.x + p.y * p.y + p.z * p.z + e2);
    const T v2 = p.vx * p.vx + p.vy * p.vy + p.vz * p.vz;
    u = T(2) * m / r;
    return v2 - u;
}

// Driver with the energy check for softened orbits (:523-614).
template <typename T> __device__ __noinline__ State<T> kepler_propagate(T dt, T m, T e2, const State<T>& p0)
{
    State<T> p = kepler_substep(dt, m, e2, p0);
    if (e2 == T(0)) return p;
    const T r2 = p0.x * p0.x + p0.y * p0.y + p0.z * p0.z;
    if (!(r2 > T(0))) return p0;
    T u0, u1;
    const T e0 = kepler_energy(m, e2, p0, u0);
    T e1 = kepler_energy(m, e2, p, u1);
    const T tol = T(64) * KeplerTol<T>::value();
    if (T(2) * k_abs(e1 - e0) < tol * (u1 + u0)) return p;
    int n = 1;
    bool bad = true;
    for (int level = 0; level < MAX_DOUBLINGS; ++level) {
        n *= 2;
        bad = false;
        p = p0;
        const T h = dt / T(n);
        for (int i = 0; i < n; ++i) {
            p = kepler_substep(h, m, e2, p);
            e1 = kepler_energy(m, e2, p, u1);
            if (T(2) * k_abs(e1 - e0) > tol * (u1 + u0)) { bad = true; break; }
        }
        if (!bad) break;
    }
    if (bad) atomicAdd(&kepler_limit_hits, 1u);
    return p;
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
template <bool FAST, typename T> TUPAN_DEV bool twobody_step(T dt, T m, T e2, State<T>& p)
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
    if (r2 > R * R) twobody_leapfrog(dt, m, e2, p);
    else p = kepler_propagate(dt, m, e2, p);
    return true;
}

// flag in {-2,-1,1,2}: where the free drift is taken out (:126-191); other values: no-op.
template <bool FAST, typename T> TUPAN_DEV bool twobody_evolve(T dt, int flag, T m, T e2, State<T>& p)
{
    if (flag == -1) {
        p.x -= p.vx * dt; p.y -= p.vy * dt; p.z -= p.vz * dt;
        return twobody_step<FAST>(dt, m, e2, p);
    } else if (flag == 1) {
        if (!twobody_step<FAST>(dt, m, e2, p)) return false;
        p.x -= p.vx * dt; p.y -= p.vy * dt; p.z -= p.vz * dt;
    } else if (flag == -2) {
        const T h = dt / T(2);
        p.x -= p.vx * h; p.y -= p.vy * h; p.z -= p.vz * h;
        if (!twobody_step<FAST>(dt, m, e2, p)) return false;
        p.x -= p.vx * h; p.y -= p.vy * h; p.z -= p.vz * h;
    } else if (flag == 2) {
        if (!twobody_step<FAST>(dt / T(2), m, e2, p)) return false;
        p.x -= p.vx * dt; p.y -= p.vy * dt; p.z -= p.vz * dt;
        return twobody_step<FAST>(dt / T(2), m, e2, p);
    }
    return true;
}

// =======================================================================================
// sakura -- replaces sakura_kernel (sakura_kernel.c:5-64, core sakura_kernel_common.h:194-243)
// =======================================================================================
template <typename T> struct SakuraParams { T dt; int flag; };
// FLAG is a template parameter: with the flag tested at run time inside the pair, the four
// variants met at a join point and every pair paid ~40 register moves and 6 branches for it
// (profiles/r01_sakura_defer_ncu_summary.txt: 139 instructions per pair, 66 of them FP64).
// FLAG = 0 stands for every other value (no-op, as in the reference).
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
    static TUPAN_DEV void pair_slow(const T (&d)[ND], const Params& p, T (&c)[NA])
    {
        const State<T> p0 = {d[D_X], d[D_Y], d[D_Z], d[D_VX], d[D_VY], d[D_VZ]};
        State<T> q = p0;
        twobody_evolve<false>(p.dt, FLAG, d[D_M], d[D_E2], q);
        const T mu = d[D_MJ] / d[D_M];
        c[0] = mu * (q.x - p0.x); c[1] = mu * (q.y - p0.y); c[2] = mu * (q.z - p0.z);
        c[3] = mu * (q.vx - p0.vx); c[4] = mu * (q.vy - p0.vy); c[5] = mu * (q.vz - p0.vz);
    }
    static TUPAN_DEV void combine(T (&a)[NA], const T (&b)[NA]) { sum_combine(a, b); }
    static TUPAN_DEV void finish(const T* const*, long long i, const T (&a)[NA], const Params&, T* const* out)
    {
#pragma unroll
        for (int k = 0; k < NO; ++k) out[k][i] = a[k];
    }
};

template <typename T, int FLAG> struct Defers<SakuraOp<T, FLAG>> { enum { value = 1 }; };
template <typename T, int FLAG> struct OpCost<SakuraOp<T, FLAG>> { enum { value = (FLAG == 2 || FLAG == -2) ? 90 : 62 }; };

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
    const State<T> q = kepler_propagate(dt, m, e2, rel);
    out.p[0][i0] = cx + jmu * q.x;  out.p[1][i0] = cy + jmu * q.y;  out.p[2][i0] = cz + jmu * q.z;
    out.p[3][i0] = cu + jmu * q.vx; out.p[4][i0] = cv + jmu * q.vy; out.p[5][i0] = cw + jmu * q.vz;
    out.p[0][i1] = cx - imu * q.x;  out.p[1][i1] = cy - imu * q.y;  out.p[2][i1] = cz - imu * q.z;
    out.p[3][i1] = cu - imu * q.vx; out.p[4][i1] = cv - imu * q.vy; out.p[5][i1] = cw - imu * q.vz;
}

}  // namespace tupan
