// ops.cuh -- per-pair physics of the Newtonian kernels, written for the pair engine.
//
// Each Op states which reference kernel it replaces (file:line in tupan/lib/src) and the
// reference's own flops-per-pair convention (its "Total flop count" comments), which is the
// convention bench.py reports rooflines in.  The arithmetic is arranged for the GPU (FMA
// chains, rsqrt seed + one cubic step, mask read off the exponent of r2) and therefore differs
// from the reference in rounding, not in meaning: results agree to ~1e-15 per pair (fp64).
//
// Every masked kernel computes r2 = rx^2 + ry^2 + rz^2 on its own and x = r2 + (ie2 + je2):
// one FP64 instruction more than folding e2 into the FMA chain, but the mask then costs two
// integer instructions instead of seven (common.cuh) -- a net gain of ~8 % on B200.
//
// Caller array order (libtupan.h): 8-array kernels pass  m, rx, ry, rz, e2, vx, vy, vz.
#pragma once
#include "pair_engine.cuh"

namespace tupan {

struct NoParams {};

// Packed row layouts (shared by several kernels so one packed buffer / one all-gather can
// feed them all).
//   Row5 : rx ry rz m  e2                      (phi, acc)
//   Row8 : rx ry rz m  vx vy vz e2             (acc_jerk, tstep, pnacc, nreg_X, sakura)
//   Row14: Row8 + ax ay az jx jy jz            (snap_crackle)
//   RowV : vx vy vz m  ax ay az                (nreg_V)
enum { JX = 0, JY = 1, JZ = 2, JM = 3 };
enum { J5_E2 = 4 };
enum { J8_VX = 4, J8_VY = 5, J8_VZ = 6, J8_E2 = 7 };
enum { J14_AX = 8, J14_AY = 9, J14_AZ = 10, J14_JX = 11, J14_JY = 12, J14_JZ = 13 };

template <typename T, int NJP>
TUPAN_DEV void pack_row5(const T* const* j, long long r, T (&row)[NJP])
{
    row[JM] = j[0][r]; row[JX] = j[1][r]; row[JY] = j[2][r]; row[JZ] = j[3][r]; row[J5_E2] = j[4][r];
}
template <typename T, int NJP>
TUPAN_DEV void pack_row8(const T* const* j, long long r, T (&row)[NJP])
{
    row[JM] = j[0][r]; row[JX] = j[1][r]; row[JY] = j[2][r]; row[JZ] = j[3][r];
    row[J8_E2] = j[4][r]; row[J8_VX] = j[5][r]; row[J8_VY] = j[6][r]; row[J8_VZ] = j[7][r];
}

template <typename T, int NA>
TUPAN_DEV void sum_combine(T (&a)[NA], const T (&b)[NA])
{
#pragma unroll
    for (int k = 0; k < NA; ++k) a[k] += b[k];
}
template <typename T, int NA>
TUPAN_DEV void zero_all(T (&a)[NA])
{
#pragma unroll
    for (int k = 0; k < NA; ++k) a[k] = T(0);
}

// =======================================================================================
// phi  -- replaces phi_kernel (phi_kernel.c:5-35, core phi_kernel_common.h:7-32); 14 flop/pair
// =======================================================================================
template <typename T> struct PhiOp {
    typedef T real;
    typedef NoParams Params;
    enum { NI = 4, NJ = 5, NA = 1, NO = 1, WPT = 4, UNROLL = 4 };
    enum { NJP = round_up(NJ, Vec16<T>::N) };
    enum { IX, IY, IZ, IE };
    static TUPAN_DEV void load_i(const T* const* a, long long i, T (&s)[NI])
    {
        s[IX] = a[1][i]; s[IY] = a[2][i]; s[IZ] = a[3][i]; s[IE] = a[4][i];
    }
    static TUPAN_DEV void pack_j(const T* const* j, long long r, T (&row)[NJP]) { pack_row5(j, r, row); }
    static TUPAN_DEV void zero(T (&a)[NA]) { zero_all(a); }
    static TUPAN_DEV void pair(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA], const Params&)
    {
        T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
        T r2 = rx * rx; r2 = fma(ry, ry, r2); r2 = fma(rz, rz, r2);
        T x = r2 + (s[IE] + row[J5_E2]);
        T r1 = rsqrt_scaled<true>(x, r2, T(1), T(0.5), T(0.375));
        a[0] = fma(-row[JM], r1, a[0]);
    }
    static TUPAN_DEV void combine(T (&a)[NA], const T (&b)[NA]) { sum_combine(a, b); }
    static TUPAN_DEV void finish(const T* const*, long long i, const T (&a)[NA], const Params&, T* const* out)
    {
        out[0][i] = a[0];
    }

    // ---- grouped form (pair_kernel_grouped, fp64; see AccJerkOp / AccOp): 13 FP64 instructions per pair
    // instead of 14 (e2_i in the r2 chain, the mask tested per group); the accumulations of the W pairs
    // of a row share the row's mass and run back to back.
    // Measured (profiles/r02_kernel_lab5_phi.txt, tools/kernel_lab3.cu -DLAB_OP=2): ungrouped 1072 Gpair/s (34.7 clocks per
    // pair), 3 x 2 1200 (31.0; 110 registers, two CTAs per SM), 8 x 2 1192, 6 x 2 1130.
#ifndef TUPAN_PHI_GROUPED
#define TUPAN_PHI_GROUPED 1
#define TUPAN_PHI_GW 3
#define TUPAN_PHI_GU 2
#endif
    enum { GROUPED = (TUPAN_PHI_GROUPED != 0 && sizeof(T) == 8), GW = TUPAN_PHI_GW, GU = TUPAN_PHI_GU, GNT = 256,
           GMODE = 8, GALT = 0 };
    struct PV { T r1, m; };
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase1(const T (*s)[NI], const T (*rows)[NJP], PV (&o)[W * U], const Params&, int)
    {
        constexpr int G = W * U;
        T rx[G], ry[G], rz[G], r2[G], x[G], y0[G], t[G];
#pragma unroll
        for (int p = 0; p < G; ++p) {
            const T(&si)[NI] = s[p % W];
            const T(&rw)[NJP] = rows[p / W];
            rx[p] = si[IX] - rw[JX]; ry[p] = si[IY] - rw[JY]; rz[p] = si[IZ] - rw[JZ];
        }
#pragma unroll
        for (int p = 0; p < G; ++p) r2[p] = fma(rx[p], rx[p], s[p % W][IE]);
#pragma unroll
        for (int p = 0; p < G; ++p) r2[p] = fma(ry[p], ry[p], r2[p]);
#pragma unroll
        for (int p = 0; p < G; ++p) r2[p] = fma(rz[p], rz[p], r2[p]);
        group_fold_seeds<G>(r2, [&](int p) { return __double2hiint(s[p % W][IE]); }, [&](int p) { return rows[p / W][J5_E2]; },
                            [&](int p) { T q = rx[p] * rx[p]; q = fma(ry[p], ry[p], q); return fma(rz[p], rz[p], q); }, x, y0);
        group_rsqrt_step<G>(x, y0, T(1), T(0.5), T(0.375), t);
#pragma unroll
        for (int p = 0; p < G; ++p) { o[p].r1 = t[p]; o[p].m = rows[p / W][JM]; }
    }
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase2(PV (&o)[W * U], T (*a)[NA], const Params&)
    {
#pragma unroll
        for (int p = 0; p < W * U; ++p) a[p % W][0] = fma(-o[p].m, o[p].r1, a[p % W][0]);   // p / W = row: m shared
    }
};

// =======================================================================================
// acc  -- replaces acc_kernel (acc_kernel.c:5-41, core acc_kernel_common.h:7-38); 20 flop/pair
// =======================================================================================
template <typename T> struct AccOp {
    typedef T real;
    typedef NoParams Params;
    enum { NI = 4, NJ = 5, NA = 3, NO = 3, WPT = 4, UNROLL = 4 };
    enum { NJP = round_up(NJ, Vec16<T>::N) };
    enum { IX, IY, IZ, IE };
    static TUPAN_DEV void load_i(const T* const* a, long long i, T (&s)[NI])
    {
        s[IX] = a[1][i]; s[IY] = a[2][i]; s[IZ] = a[3][i]; s[IE] = a[4][i];
    }
    static TUPAN_DEV void pack_j(const T* const* j, long long r, T (&row)[NJP]) { pack_row5(j, r, row); }
    static TUPAN_DEV void zero(T (&a)[NA]) { zero_all(a); }
    static TUPAN_DEV void pair(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA], const Params&)
    {
        T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
        T r2 = rx * rx; r2 = fma(ry, ry, r2); r2 = fma(rz, rz, r2);
        T x = r2 + (s[IE] + row[J5_E2]);
        InvR<T> w = soft_inv<false>(x, r2);
        T g = -(row[JM] * w.r3);
        a[0] = fma(g, rx, a[0]); a[1] = fma(g, ry, a[1]); a[2] = fma(g, rz, a[2]);
    }
    static TUPAN_DEV void combine(T (&a)[NA], const T (&b)[NA]) { sum_combine(a, b); }
    static TUPAN_DEV void finish(const T* const*, long long i, const T (&a)[NA], const Params&, T* const* out)
    {
        out[0][i] = a[0]; out[1][i] = a[1]; out[2][i] = a[2];
    }

    // ---- grouped form (pair_kernel_grouped, fp64; see AccJerkOp for the why): G = W x U pairs operation
    // by operation, 18 FP64 instructions per pair instead of 19 -- e2_i rides in the r2 chain and the mask
    // (r2 zero or denormal) is tested once per group through a necessary condition, as in
    // AccJerkOp::group_phase1 -- and the three accumulations of a pair, which share g, in a block of
    // their own.
    // Measured (profiles/r02_kernel_lab4_acc.txt, tools/kernel_lab3.cu -DLAB_OP=1): ungrouped 779.7 Gpair/s (47.7 clocks
    // per pair), 6 x 2 902.5 (41.3; 238 registers, 1536 particles per CTA), 3 x 2 873.3 (42.6; 768 per CTA:
    // the second shape, for small ni).
#ifndef TUPAN_ACC_GROUPED
#define TUPAN_ACC_GROUPED 1
#define TUPAN_ACC_GW 6
#define TUPAN_ACC_GU 2
#endif
    enum { GROUPED = (TUPAN_ACC_GROUPED != 0 && sizeof(T) == 8), GW = TUPAN_ACC_GW, GU = TUPAN_ACC_GU, GNT = 256,
           GMODE = 8 };
    enum { GALT = 1, GW2 = 3, GU2 = 2, GMODE2 = 8, GCOST2_PERMILLE = 1033 };
    struct PV { T rx, ry, rz, g; };
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase1(const T (*s)[NI], const T (*rows)[NJP], PV (&o)[W * U], const Params&, int)
    {
        constexpr int G = W * U;
        T r2[G], x[G], y0[G], t[G], h[G];
#pragma unroll
        for (int p = 0; p < G; ++p) {
            const T(&si)[NI] = s[p % W];
            const T(&rw)[NJP] = rows[p / W];
            o[p].rx = si[IX] - rw[JX]; o[p].ry = si[IY] - rw[JY]; o[p].rz = si[IZ] - rw[JZ];
        }
#pragma unroll
        for (int p = 0; p < G; ++p) r2[p] = fma(o[p].rx, o[p].rx, s[p % W][IE]);
#pragma unroll
        for (int p = 0; p < G; ++p) r2[p] = fma(o[p].ry, o[p].ry, r2[p]);
#pragma unroll
        for (int p = 0; p < G; ++p) r2[p] = fma(o[p].rz, o[p].rz, r2[p]);
        group_fold_seeds<G>(r2, [&](int p) { return __double2hiint(s[p % W][IE]); }, [&](int p) { return rows[p / W][J5_E2]; },
                            [&](int p) { T q = o[p].rx * o[p].rx; q = fma(o[p].ry, o[p].ry, q); return fma(o[p].rz, o[p].rz, q); },
                            x, y0);
        group_rsqrt_step<G>(x, y0, T(1), T(0.5), T(0.375), t);                      // 1/sqrt(x)
#pragma unroll
        for (int p = 0; p < G; ++p) h[p] = t[p] * t[p];
#pragma unroll
        for (int p = 0; p < G; ++p) h[p] = h[p] * t[p];                  // x^-3/2
#pragma unroll
        for (int p = 0; p < G; ++p) o[p].g = -(rows[p / W][JM] * h[p]);
    }
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase2(PV (&o)[W * U], T (*a)[NA], const Params&)
    {
#pragma unroll
        for (int p = 0; p < W * U; ++p) {
            T(&ac)[NA] = a[p % W];
            ac[0] = fma(o[p].rx, o[p].g, ac[0]); ac[1] = fma(o[p].ry, o[p].g, ac[1]); ac[2] = fma(o[p].rz, o[p].g, ac[2]);
        }
    }
};

// =======================================================================================
// acc_jerk -- replaces acc_jerk_kernel (acc_jerk_kernel.c:5-61, core
// acc_jerk_kernel_common.h:7-56); 42 flop/pair by the reference's count.
// FP64-pipe instructions per pair here: 6 (differences) + 3 (r2) + 2 (e2, x) + 3 (r.v) + 5
// (rsqrt step) + 2 (3/x, 3 sqrt3 x^-3/2) + 1 (alpha) + 1 (g) + 9 (FMA updates) = 32, plus
// 2 LDS.128, MUFU, ISETP, SEL.  Two constant factors ride for free: the rsqrt step returns
// sqrt(3/x), so its square is the 3/x of alpha = 3 (r.v)/x, and the 3 sqrt 3 that its cube
// carries is divided out of the mass once, in pack_j (row[JM] = mj / (3 sqrt 3)).
// =======================================================================================
template <typename T> struct AccJerkOp {
    typedef T real;
    typedef NoParams Params;
    enum { NI = 7, NJ = 8, NA = 6, NO = 6, WPT = 2, UNROLL = 8 };
    enum { NJP = round_up(NJ, Vec16<T>::N) };
    enum { IX, IY, IZ, IE, IVX, IVY, IVZ };
    static TUPAN_DEV void load_i(const T* const* a, long long i, T (&s)[NI])
    {
        s[IX] = a[1][i]; s[IY] = a[2][i]; s[IZ] = a[3][i]; s[IE] = a[4][i];
        s[IVX] = a[5][i]; s[IVY] = a[6][i]; s[IVZ] = a[7][i];
    }
    static TUPAN_DEV void pack_j(const T* const* j, long long r, T (&row)[NJP])
    {
        pack_row8(j, r, row);
        row[JM] = row[JM] * T(0.19245008972987526);          // 1 / (3 sqrt 3)
    }
    static TUPAN_DEV void zero(T (&a)[NA]) { zero_all(a); }
    static TUPAN_DEV void pair(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA], const Params&)
    {
        T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
        T vx = s[IVX] - row[J8_VX], vy = s[IVY] - row[J8_VY], vz = s[IVZ] - row[J8_VZ];
        T r2 = rx * rx; r2 = fma(ry, ry, r2); r2 = fma(rz, rz, r2);
        T x = r2 + (s[IE] + row[J8_E2]);
        T rv = rx * vx; rv = fma(ry, vy, rv); rv = fma(rz, vz, rv);
        // sqrt(3/x): k = sqrt 3, k/2, 3k/8
        T r1 = rsqrt_scaled<false>(x, r2, T(1.7320508075688772), T(0.86602540378443865), T(0.64951905283832900));
        T q2 = r1 * r1;                    // 3/x
        T q3 = q2 * r1;                    // 3 sqrt3 x^-3/2
        T alpha = q2 * rv;                 // 3 (r.v)/x
        T g = -(row[JM] * q3);             // -mj x^-3/2
        vx = fma(-alpha, rx, vx); vy = fma(-alpha, ry, vy); vz = fma(-alpha, rz, vz);
        a[0] = fma(g, rx, a[0]); a[1] = fma(g, ry, a[1]); a[2] = fma(g, rz, a[2]);
        a[3] = fma(g, vx, a[3]); a[4] = fma(g, vy, a[4]); a[5] = fma(g, vz, a[5]);
    }
    static TUPAN_DEV void combine(T (&a)[NA], const T (&b)[NA]) { sum_combine(a, b); }
    static TUPAN_DEV void finish(const T* const*, long long i, const T (&a)[NA], const Params&, T* const* out)
    {
#pragma unroll
        for (int k = 0; k < NO; ++k) out[k][i] = a[k];
    }

    // ---- grouped form (pair_kernel_grouped, fp64): the same 32 FP64 operations per pair, written for
    // the G = W x U pairs of a row group operation by operation.  Three basic blocks per group:
    //   1a  differences, then the r2 and r.v chains step by step -- fma(ry, ry, r2) next to
    //       fma(ry, vy, rv): the second finds ry in the operand-reuse cache;
    //   1b  x, rsqrt seed + cubic step, 3/x, -alpha, 3 sqrt3 x^-3/2 (no three-register DFMA at all);
    //   2   v' = v - alpha r for every pair, then per pair g = -(m q3) and the six accumulations,
    //       all with g as the LAST-defined multiplicand (ptxas puts the operand it shares between
    //       consecutive DFMAs in one slot when it is the later-defined one).
    // Measured (profiles/r02_kernel_lab_grouped.txt, r02_kernel_lab3_w3.txt, r02_kernel_lab3_fold.txt):
    // W x U = 3 x 2 leaves 2.0 uncached three-register DFMAs per pair (the first of each run of DFMAs
    // that share an operand) instead of 9.3, and 5.8 instead of 6.9 non-FP64 instructions per pair: 71.3
    // clocks per pair in the lab kernel (round-1 kernel 75.1, 2 x 4 with one-trip loops 72.9); with e2_i
    // folded into the r2 chain (31 FP64 instructions per pair, bit 4) 68.6 = 0.612 of the nominal peak.
#ifndef TUPAN_AJ_GW
#define TUPAN_AJ_GW 3
#define TUPAN_AJ_GU 2
#define TUPAN_AJ_GNT 256
#endif
    // GMODE bit 0: block 1a on its own; bit 1: g formed in block 2 (see above); bit 2: block 2 pair by
    // pair; bit 3: the blocks are fenced by `if (one != 0)` instead of one-trip loops (two uniform
    // branches per group instead of two loop headers with their counters); bit 4: e2_i folded into the
    // r2 chain, the mask tested per group (group_phase1)
    enum { GROUPED = (sizeof(T) == 8), GW = TUPAN_AJ_GW, GU = TUPAN_AJ_GU, GNT = TUPAN_AJ_GNT, GMODE = 27 };
    // second shape, 2 x 4 (512 instead of 768 particles per CTA; 71.9 clocks per pair): small and medium ni
    enum { GALT = 1, GW2 = 2, GU2 = 4, GMODE2 = 27, GCOST2_PERMILLE = 1047 };
    struct PV { T rx, ry, rz, vx, vy, vz, na, q3, mj; };
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase1(const T (*s)[NI], const T (*rows)[NJP], PV (&o)[W * U], const Params&, int one)
    {
        constexpr int G = W * U;
        constexpr bool FOLD = (MODE & 16) != 0;      // e2_i rides in the r2 chain (one DADD per pair less)
        T r2[G], rv[G], e[G], y0[G];
#pragma unroll
        for (int p = 0; p < G; ++p) {
            const T(&si)[NI] = s[p % W];
            const T(&rw)[NJP] = rows[p / W];
            o[p].rx = si[IX] - rw[JX]; o[p].ry = si[IY] - rw[JY]; o[p].rz = si[IZ] - rw[JZ];
            o[p].vx = si[IVX] - rw[J8_VX]; o[p].vy = si[IVY] - rw[J8_VY]; o[p].vz = si[IVZ] - rw[J8_VZ];
            e[p] = FOLD ? rw[J8_E2] : si[IE] + rw[J8_E2];
            o[p].mj = rw[JM];
        }
        auto chains = [&]() {
#pragma unroll
            for (int p = 0; p < G; ++p) {
                r2[p] = FOLD ? fma(o[p].rx, o[p].rx, s[p % W][IE]) : o[p].rx * o[p].rx;
                rv[p] = o[p].rx * o[p].vx;
            }
#pragma unroll
            for (int p = 0; p < G; ++p) { r2[p] = fma(o[p].ry, o[p].ry, r2[p]); rv[p] = fma(o[p].ry, o[p].vy, rv[p]); }
#pragma unroll
            for (int p = 0; p < G; ++p) { r2[p] = fma(o[p].rz, o[p].rz, r2[p]); rv[p] = fma(o[p].rz, o[p].vz, rv[p]); }
        };
        if (MODE & 8) {                 // block 1a behind a never-skipped branch (cheaper than a loop)
            if (one != 0) chains();
        } else {
#pragma unroll 1
            for (int z = 0; z < ((MODE & 1) ? one : 1); ++z) chains();       // block 1a
        }
        T x[G], t[G], h[G];
        if (FOLD) {
            // x = ((e2_i + rx^2) + ry^2 + rz^2) + e2_j: 31 FP64 instructions per pair; the mask per group
            // (group_fold_seeds, common.cuh)
            group_fold_seeds<G>(r2, [&](int p) { return __double2hiint(s[p % W][IE]); }, [&](int p) { return e[p]; },
                                [&](int p) { T q = o[p].rx * o[p].rx; q = fma(o[p].ry, o[p].ry, q); return fma(o[p].rz, o[p].rz, q); },
                                x, y0);
        } else {
#pragma unroll
            for (int p = 0; p < G; ++p) x[p] = r2[p] + e[p];                 // x = r2 + e2
#pragma unroll
            for (int p = 0; p < G; ++p) y0[p] = rsqrt_seed_masked<false>(x[p], r2[p]);
        }
        // sqrt(3/x): k = sqrt 3, k/2, 3k/8
        group_rsqrt_step<G>(x, y0, T(1.7320508075688772), T(0.86602540378443865), T(0.64951905283832900), t);
#pragma unroll
        for (int p = 0; p < G; ++p) h[p] = t[p] * t[p];                  // 3/x
#pragma unroll
        for (int p = 0; p < G; ++p) { o[p].na = -(h[p] * rv[p]); o[p].q3 = h[p] * t[p]; }   // -alpha; 3 sqrt3 x^-3/2
        if (!(MODE & 2)) {
#pragma unroll
            for (int p = 0; p < G; ++p) o[p].q3 = -(o[p].mj * o[p].q3);      // g
        }
    }
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase2(PV (&o)[W * U], T (*a)[NA], const Params&)
    {
        constexpr int G = W * U;
        if (MODE & 4) {          // pair by pair
#pragma unroll
            for (int p = 0; p < G; ++p) {
                o[p].vx = fma(o[p].na, o[p].rx, o[p].vx);
                o[p].vy = fma(o[p].na, o[p].ry, o[p].vy);
                o[p].vz = fma(o[p].na, o[p].rz, o[p].vz);
                T g = o[p].q3;
                if (MODE & 2) asm volatile("mul.f64 %0, %1, %2;" : "=d"(g) : "d"(-o[p].mj), "d"(o[p].q3));
                T(&ac)[NA] = a[p % W];
                ac[0] = fma(g, o[p].rx, ac[0]); ac[1] = fma(g, o[p].ry, ac[1]); ac[2] = fma(g, o[p].rz, ac[2]);
                ac[3] = fma(g, o[p].vx, ac[3]); ac[4] = fma(g, o[p].vy, ac[4]); ac[5] = fma(g, o[p].vz, ac[5]);
            }
            return;
        }
#pragma unroll
        for (int p = 0; p < G; ++p) {
            o[p].vx = fma(o[p].na, o[p].rx, o[p].vx);
            o[p].vy = fma(o[p].na, o[p].ry, o[p].vy);
            o[p].vz = fma(o[p].na, o[p].rz, o[p].vz);
        }
#pragma unroll
        for (int p = 0; p < G; ++p) {
            // the product is formed HERE, after v', on purpose (see above); volatile keeps it in
            // this block: hoisted out of the one-trip loop it would be defined before v' again
            T g = o[p].q3;
            if (MODE & 2) asm volatile("mul.f64 %0, %1, %2;" : "=d"(g) : "d"(-o[p].mj), "d"(o[p].q3));
            T(&ac)[NA] = a[p % W];
            ac[0] = fma(o[p].rx, g, ac[0]); ac[1] = fma(o[p].ry, g, ac[1]); ac[2] = fma(o[p].rz, g, ac[2]);
            ac[3] = fma(o[p].vx, g, ac[3]); ac[4] = fma(o[p].vy, g, ac[4]); ac[5] = fma(o[p].vz, g, ac[5]);
        }
    }
};

// =======================================================================================
// snap_crackle -- replaces snap_crackle_kernel (snap_crackle_kernel.c:5-83, core
// snap_crackle_kernel_common.h:7-95); 114 flop/pair.
// Caller arrays: m rx ry rz e2 vx vy vz ax ay az jx jy jz.
// =======================================================================================
template <typename T> struct SnapCrackleOp {
    typedef T real;
    typedef NoParams Params;
    enum { NI = 13, NJ = 14, NA = 6, NO = 6, WPT = 1, UNROLL = 2 };
    enum { NJP = round_up(NJ, Vec16<T>::N) };
    enum { IX, IY, IZ, IE, IVX, IVY, IVZ, IAX, IAY, IAZ, IJX, IJY, IJZ };
    static TUPAN_DEV void load_i(const T* const* a, long long i, T (&s)[NI])
    {
        s[IX] = a[1][i]; s[IY] = a[2][i]; s[IZ] = a[3][i]; s[IE] = a[4][i];
        s[IVX] = a[5][i]; s[IVY] = a[6][i]; s[IVZ] = a[7][i];
        s[IAX] = a[8][i]; s[IAY] = a[9][i]; s[IAZ] = a[10][i];
        s[IJX] = a[11][i]; s[IJY] = a[12][i]; s[IJZ] = a[13][i];
    }
    static TUPAN_DEV void pack_j(const T* const* j, long long r, T (&row)[NJP])
    {
        pack_row8(j, r, row);
        row[J14_AX] = j[8][r]; row[J14_AY] = j[9][r]; row[J14_AZ] = j[10][r];
        row[J14_JX] = j[11][r]; row[J14_JY] = j[12][r]; row[J14_JZ] = j[13][r];
    }
    static TUPAN_DEV void zero(T (&a)[NA]) { zero_all(a); }
    static TUPAN_DEV void pair(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA], const Params&)
    {
        T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
        T vx = s[IVX] - row[J8_VX], vy = s[IVY] - row[J8_VY], vz = s[IVZ] - row[J8_VZ];
        T ax = s[IAX] - row[J14_AX], ay = s[IAY] - row[J14_AY], az = s[IAZ] - row[J14_AZ];
        T jx = s[IJX] - row[J14_JX], jy = s[IJY] - row[J14_JY], jz = s[IJZ] - row[J14_JZ];
        T r2 = rx * rx; r2 = fma(ry, ry, r2); r2 = fma(rz, rz, r2);
        T x = r2 + (s[IE] + row[J8_E2]);
        T rv = rx * vx; rv = fma(ry, vy, rv); rv = fma(rz, vz, rv);
        T v2 = vx * vx; v2 = fma(vy, vy, v2); v2 = fma(vz, vz, v2);
        T rj = rx * jx; rj = fma(ry, jy, rj); rj = fma(rz, jz, rj);
        T ra = rx * ax; ra = fma(ry, ay, ra); ra = fma(rz, az, ra);
        T va = vx * ax; va = fma(vy, ay, va); va = fma(vz, az, va);
        InvR<T> w = soft_inv<false>(x, r2);

        // same recurrences as snap_crackle_kernel_common.h:63-84; the vector updates are chains of
        // FMAs (a' = a - 2 alpha v' - beta r as two FMAs per component instead of mul + fma + sub,
        // j' likewise three instead of four operations): 78 instead of 84 FP64 instructions per pair
        T alpha = rv * w.r2;
        T alpha2 = alpha * alpha;
        T beta = T(3) * fma(v2 + ra, w.r2, alpha2);
        T gamma = fma(fma(T(3), va, rj), w.r2, alpha * fma(T(-4), alpha2, beta));
        alpha *= T(3);
        gamma *= T(3);
        vx = fma(-alpha, rx, vx); vy = fma(-alpha, ry, vy); vz = fma(-alpha, rz, vz);
        T a2 = T(2) * alpha;
        ax = fma(-a2, vx, ax); ay = fma(-a2, vy, ay); az = fma(-a2, vz, az);
        ax = fma(-beta, rx, ax); ay = fma(-beta, ry, ay); az = fma(-beta, rz, az);
        alpha *= T(3);
        beta *= T(3);
        jx = fma(-alpha, ax, jx); jy = fma(-alpha, ay, jy); jz = fma(-alpha, az, jz);
        jx = fma(-beta, vx, jx); jy = fma(-beta, vy, jy); jz = fma(-beta, vz, jz);
        jx = fma(-gamma, rx, jx); jy = fma(-gamma, ry, jy); jz = fma(-gamma, rz, jz);
        T g = row[JM] * w.r3;
        a[0] = fma(-g, ax, a[0]); a[1] = fma(-g, ay, a[1]); a[2] = fma(-g, az, a[2]);
        a[3] = fma(-g, jx, a[3]); a[4] = fma(-g, jy, a[4]); a[5] = fma(-g, jz, a[5]);
    }
    static TUPAN_DEV void combine(T (&a)[NA], const T (&b)[NA]) { sum_combine(a, b); }
    static TUPAN_DEV void finish(const T* const*, long long i, const T (&a)[NA], const Params&, T* const* out)
    {
#pragma unroll
        for (int k = 0; k < NO; ++k) out[k][i] = a[k];
    }


    // ---- grouped form (pair_kernel_grouped, fp64; see AccJerkOp and pair_engine.cuh).  23.5 of the
    // 84 FP64 instructions of the plain body are DFMAs that collect three registers: the dot products
    // r.v r.j r.a v.a, and the vector updates.  Here the dot products of a group are written component
    // by component -- fma(ry, ry, r2), fma(ry, vy, rv), fma(ry, jy, rj), fma(ry, ay, ra) share ry --
    // in a block of their own, and the vector updates + accumulations scalar by scalar in another.
#ifndef TUPAN_SC_GW
#define TUPAN_SC_GW 1
#define TUPAN_SC_GU 4
#define TUPAN_SC_GNT 256
#endif
#ifndef TUPAN_SC_GROUPED
#define TUPAN_SC_GROUPED 1
#endif
    enum { GROUPED = (TUPAN_SC_GROUPED != 0 && sizeof(T) == 8), GW = TUPAN_SC_GW, GU = TUPAN_SC_GU, GNT = TUPAN_SC_GNT,
           GMODE = 1, GALT = 0 };
    struct PV { T rx, ry, rz, vx, vy, vz, ax, ay, az, jx, jy, jz, al, a2, be, al3, be3, ga, g; };
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase1(const T (*s)[NI], const T (*rows)[NJP], PV (&o)[W * U], const Params&, int one)
    {
        constexpr int G = W * U;
        T r2[G], rv[G], v2[G], rj[G], ra[G], va[G], e[G];
#pragma unroll
        for (int p = 0; p < G; ++p) {
            const T(&si)[NI] = s[p % W];
            const T(&rw)[NJP] = rows[p / W];
            o[p].rx = si[IX] - rw[JX]; o[p].ry = si[IY] - rw[JY]; o[p].rz = si[IZ] - rw[JZ];
            o[p].vx = si[IVX] - rw[J8_VX]; o[p].vy = si[IVY] - rw[J8_VY]; o[p].vz = si[IVZ] - rw[J8_VZ];
            o[p].ax = si[IAX] - rw[J14_AX]; o[p].ay = si[IAY] - rw[J14_AY]; o[p].az = si[IAZ] - rw[J14_AZ];
            o[p].jx = si[IJX] - rw[J14_JX]; o[p].jy = si[IJY] - rw[J14_JY]; o[p].jz = si[IJZ] - rw[J14_JZ];
            e[p] = si[IE] + rw[J8_E2];
        }
#pragma unroll 1
        for (int z = 0; z < ((MODE & 1) ? one : 1); ++z) {       // the dot products, component by component
#pragma unroll
            for (int p = 0; p < G; ++p) {
                r2[p] = o[p].rx * o[p].rx; rv[p] = o[p].rx * o[p].vx; rj[p] = o[p].rx * o[p].jx; ra[p] = o[p].rx * o[p].ax;
                v2[p] = o[p].vx * o[p].vx; va[p] = o[p].vx * o[p].ax;
            }
#pragma unroll
            for (int p = 0; p < G; ++p) {
                r2[p] = fma(o[p].ry, o[p].ry, r2[p]); rv[p] = fma(o[p].ry, o[p].vy, rv[p]);
                rj[p] = fma(o[p].ry, o[p].jy, rj[p]); ra[p] = fma(o[p].ry, o[p].ay, ra[p]);
                v2[p] = fma(o[p].vy, o[p].vy, v2[p]); va[p] = fma(o[p].vy, o[p].ay, va[p]);
            }
#pragma unroll
            for (int p = 0; p < G; ++p) {
                r2[p] = fma(o[p].rz, o[p].rz, r2[p]); rv[p] = fma(o[p].rz, o[p].vz, rv[p]);
                rj[p] = fma(o[p].rz, o[p].jz, rj[p]); ra[p] = fma(o[p].rz, o[p].az, ra[p]);
                v2[p] = fma(o[p].vz, o[p].vz, v2[p]); va[p] = fma(o[p].vz, o[p].az, va[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < G; ++p) {
            const InvR<T> w = soft_inv<false>(r2[p] + e[p], r2[p]);
            T alpha = rv[p] * w.r2;
            const T alpha2 = alpha * alpha;
            T beta = T(3) * fma(v2[p] + ra[p], w.r2, alpha2);
            T gamma = fma(fma(T(3), va[p], rj[p]), w.r2, alpha * fma(T(-4), alpha2, beta));
            alpha *= T(3);
            o[p].al = -alpha;
            o[p].a2 = T(-2) * alpha;
            o[p].be = -beta;
            o[p].al3 = T(-3) * alpha;
            o[p].be3 = T(-3) * beta;
            o[p].ga = T(-3) * gamma;
            o[p].g = -(rows[p / W][JM] * w.r3);
        }
    }
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase2(PV (&o)[W * U], T (*a)[NA], const Params&)
    {
        constexpr int G = W * U;
#pragma unroll
        for (int p = 0; p < G; ++p) {
            o[p].vx = fma(o[p].al, o[p].rx, o[p].vx); o[p].vy = fma(o[p].al, o[p].ry, o[p].vy); o[p].vz = fma(o[p].al, o[p].rz, o[p].vz);
        }
#pragma unroll
        for (int p = 0; p < G; ++p) {
            o[p].ax = fma(o[p].be, o[p].rx, o[p].ax); o[p].ay = fma(o[p].be, o[p].ry, o[p].ay); o[p].az = fma(o[p].be, o[p].rz, o[p].az);
        }
#pragma unroll
        for (int p = 0; p < G; ++p) {
            o[p].jx = fma(o[p].ga, o[p].rx, o[p].jx); o[p].jy = fma(o[p].ga, o[p].ry, o[p].jy); o[p].jz = fma(o[p].ga, o[p].rz, o[p].jz);
        }
#pragma unroll
        for (int p = 0; p < G; ++p) {
            o[p].ax = fma(o[p].a2, o[p].vx, o[p].ax); o[p].ay = fma(o[p].a2, o[p].vy, o[p].ay); o[p].az = fma(o[p].a2, o[p].vz, o[p].az);
        }
#pragma unroll
        for (int p = 0; p < G; ++p) {
            o[p].jx = fma(o[p].be3, o[p].vx, o[p].jx); o[p].jy = fma(o[p].be3, o[p].vy, o[p].jy); o[p].jz = fma(o[p].be3, o[p].vz, o[p].jz);
        }
#pragma unroll
        for (int p = 0; p < G; ++p) {
            o[p].jx = fma(o[p].al3, o[p].ax, o[p].jx); o[p].jy = fma(o[p].al3, o[p].ay, o[p].jy); o[p].jz = fma(o[p].al3, o[p].az, o[p].jz);
        }
#pragma unroll
        for (int p = 0; p < G; ++p) {
            T(&ac)[NA] = a[p % W];
            ac[0] = fma(o[p].g, o[p].ax, ac[0]); ac[1] = fma(o[p].g, o[p].ay, ac[1]); ac[2] = fma(o[p].g, o[p].az, ac[2]);
            ac[3] = fma(o[p].g, o[p].jx, ac[3]); ac[4] = fma(o[p].g, o[p].jy, ac[4]); ac[5] = fma(o[p].g, o[p].jz, ac[5]);
        }
    }
};

// =======================================================================================
// tstep -- replaces tstep_kernel (tstep_kernel.c:5-50, core tstep_kernel_common.h:7-60);
// 42 flop/pair.  Two accumulators: sum and max of the pair frequency w2; the epilogue
// eta/sqrt(1+.) is tstep_kernel.c:46-47.  The lane/chunk reduction of the second
// accumulator is a max, as is the fused minimum time-step (see finish_min in api).
// =======================================================================================
// FP64-pipe instructions per pair: 6 (differences) + 1 (2m) + 3 (r2) + 2 (e2, x) + 3 (r.v) + 3 (v2)
// + 6 (1/r, 1/r2) + 1 (2 phi) + 2 (w2) + 2 (gamma) + 5 (eta/sqrt w2) + 1 + 1 (w2 -= ..) + 2 (sum, max)
// = 39.  Two factors ride for free: the masses are stored doubled (i-state and packed row; an
// exact scaling, so 2 phi = (2m) / r bit for bit), and eta is folded into the constants of the
// second rsqrt step.
template <typename T> struct TstepParams { T eta, eta_k1, eta_k2; };   // eta, eta/2, 3 eta/8
template <typename T> struct TstepOp {
    typedef T real;
    typedef TstepParams<T> Params;
    enum { NI = 8, NJ = 8, NA = 2, NO = 2, WPT = 2, UNROLL = 4 };
    enum { NJP = round_up(NJ, Vec16<T>::N) };
    enum { IX, IY, IZ, IE, IVX, IVY, IVZ, IM };
    static TUPAN_DEV void load_i(const T* const* a, long long i, T (&s)[NI])
    {
        s[IM] = T(2) * a[0][i];
        s[IX] = a[1][i]; s[IY] = a[2][i]; s[IZ] = a[3][i]; s[IE] = a[4][i];
        s[IVX] = a[5][i]; s[IVY] = a[6][i]; s[IVZ] = a[7][i];
    }
    static TUPAN_DEV void pack_j(const T* const* j, long long r, T (&row)[NJP])
    {
        pack_row8(j, r, row);
        row[JM] = T(2) * row[JM];
    }
    static TUPAN_DEV void zero(T (&a)[NA]) { zero_all(a); }
    static TUPAN_DEV void pair(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA], const Params& p)
    {
        T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
        T vx = s[IVX] - row[J8_VX], vy = s[IVY] - row[J8_VY], vz = s[IVZ] - row[J8_VZ];
        T m2 = s[IM] + row[JM];                                   // 2 (mi + mj)
        T r2 = rx * rx; r2 = fma(ry, ry, r2); r2 = fma(rz, rz, r2);
        T x = r2 + (s[IE] + row[J8_E2]);
        T rv = rx * vx; rv = fma(ry, vy, rv); rv = fma(rz, vz, rv);
        T v2 = vx * vx; v2 = fma(vy, vy, v2); v2 = fma(vz, vz, v2);
        // CLEAN = false (the seed's low word is not zeroed: one MOV less per rsqrt): a masked pair gets
        // r1 ~ 1e-314 instead of 0, whose square r2inv underflows to exactly 0, so w2 = 0 * (v2 + phi2) = 0
        // and gamma = 0 * (...) = 0 as before; the second seed, of w2 = 0, is clamped to a finite value,
        // and 0 * (a finite number) = 0.
        InvR<T> w = soft_inv<false>(x, r2);
        // w2 = (v2 + 2 phi)/r2 ; gamma = (w2 + 2 phi/r2)/r2 * eta/sqrt(w2) ; w2 -= gamma*rv
        T phi2 = m2 * w.r1;
        T w2 = w.r2 * (v2 + phi2);
        T gamma = w.r2 * fma(w.r2, phi2, w2);
        // masked pair: w2 is exactly 0 and so is gamma; the seed of 0 is clamped to a finite value
        // (rsqrt_scaled_clamped), gamma stays 0 -> w2 stays 0
        gamma *= rsqrt_scaled_clamped(w2, p.eta, p.eta_k1, p.eta_k2);
        w2 = fma(-gamma, rv, w2);
        a[0] += w2;
        a[1] = rmax_nonneg(a[1], w2);
    }
    static TUPAN_DEV void combine(T (&a)[NA], const T (&b)[NA])
    {
        a[0] += b[0];
        a[1] = rmax(a[1], b[1]);
    }
    static TUPAN_DEV void finish(const T* const*, long long i, const T (&a)[NA], const Params& p, T* const* out)
    {
        out[0][i] = p.eta * rsqrt_full(T(1) + a[0]);
        out[1][i] = p.eta * rsqrt_full(T(1) + a[1]);
    }

    // ---- grouped form (pair_kernel_grouped, fp64; see AccJerkOp): 36 FP64 instructions per pair instead
    // of 37 (e2_i in the r2 chain, the r2 mask tested per group); the r2 and r.v chains step by step in a
    // block of their own -- fma(ry, ry, r2) next to fma(ry, vy, rv), which finds ry in the operand-reuse
    // cache.  The second seed needs no mask, only a clamp (rsqrt_scaled_clamped).
    // Measured (profiles/r02_kernel_lab6_tstep.txt, tools/kernel_lab3.cu -DLAB_OP=3): ungrouped 394.8 Gpair/s (94.3 clocks
    // per pair), 3 x 2 429.2 (86.7), 4 x 2 427.1, 2 x 2 406.2 (the second shape, 512 particles per CTA).
#ifndef TUPAN_TSTEP_GROUPED
#define TUPAN_TSTEP_GROUPED 1
#define TUPAN_TSTEP_GW 3
#define TUPAN_TSTEP_GU 2
#endif
    enum { GROUPED = (TUPAN_TSTEP_GROUPED != 0 && sizeof(T) == 8), GW = TUPAN_TSTEP_GW, GU = TUPAN_TSTEP_GU, GNT = 256,
           GMODE = 8 };
    enum { GALT = 1, GW2 = 2, GU2 = 2, GMODE2 = 8, GCOST2_PERMILLE = 1056 };
    struct PV { T w2; };
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase1(const T (*s)[NI], const T (*rows)[NJP], PV (&o)[W * U], const Params& prm, int one)
    {
        constexpr int G = W * U;
        T rx[G], ry[G], rz[G], vx[G], vy[G], vz[G], m2[G], r2[G], rv[G], v2[G], x[G], y0[G], t[G], h[G];
#pragma unroll
        for (int p = 0; p < G; ++p) {
            const T(&si)[NI] = s[p % W];
            const T(&rw)[NJP] = rows[p / W];
            rx[p] = si[IX] - rw[JX]; ry[p] = si[IY] - rw[JY]; rz[p] = si[IZ] - rw[JZ];
            vx[p] = si[IVX] - rw[J8_VX]; vy[p] = si[IVY] - rw[J8_VY]; vz[p] = si[IVZ] - rw[J8_VZ];
            m2[p] = si[IM] + rw[JM];
        }
#pragma unroll
        for (int p = 0; p < G; ++p) v2[p] = vx[p] * vx[p];
#pragma unroll
        for (int p = 0; p < G; ++p) v2[p] = fma(vy[p], vy[p], v2[p]);
#pragma unroll
        for (int p = 0; p < G; ++p) v2[p] = fma(vz[p], vz[p], v2[p]);
        if (one != 0) {                 // the r2 and r.v chains, fenced (AccJerkOp's block 1a)
#pragma unroll
            for (int p = 0; p < G; ++p) { r2[p] = fma(rx[p], rx[p], s[p % W][IE]); rv[p] = rx[p] * vx[p]; }
#pragma unroll
            for (int p = 0; p < G; ++p) { r2[p] = fma(ry[p], ry[p], r2[p]); rv[p] = fma(ry[p], vy[p], rv[p]); }
#pragma unroll
            for (int p = 0; p < G; ++p) { r2[p] = fma(rz[p], rz[p], r2[p]); rv[p] = fma(rz[p], vz[p], rv[p]); }
        }
        group_fold_seeds<G>(r2, [&](int p) { return __double2hiint(s[p % W][IE]); }, [&](int p) { return rows[p / W][J8_E2]; },
                            [&](int p) { T q = rx[p] * rx[p]; q = fma(ry[p], ry[p], q); return fma(rz[p], rz[p], q); }, x, y0);
        group_rsqrt_step<G>(x, y0, T(1), T(0.5), T(0.375), t);                      // 1/r
#pragma unroll
        for (int p = 0; p < G; ++p) h[p] = t[p] * t[p];                  // 1/r^2
        // w2 = (v2 + 2 phi)/r2 ; gamma = (w2 + 2 phi/r2)/r2 * eta/sqrt(w2) ; w2 -= gamma*rv   (pair())
#pragma unroll
        for (int p = 0; p < G; ++p) t[p] = m2[p] * t[p];                 // phi2
#pragma unroll
        for (int p = 0; p < G; ++p) v2[p] = v2[p] + t[p];
#pragma unroll
        for (int p = 0; p < G; ++p) v2[p] = h[p] * v2[p];                // w2
#pragma unroll
        for (int p = 0; p < G; ++p) t[p] = fma(h[p], t[p], v2[p]);
#pragma unroll
        for (int p = 0; p < G; ++p) t[p] = h[p] * t[p];                  // gamma without the eta/sqrt(w2)
#pragma unroll
        for (int p = 0; p < G; ++p) h[p] = rsqrt_scaled_clamped(v2[p], prm.eta, prm.eta_k1, prm.eta_k2);
#pragma unroll
        for (int p = 0; p < G; ++p) t[p] = t[p] * h[p];
#pragma unroll
        for (int p = 0; p < G; ++p) o[p].w2 = fma(-t[p], rv[p], v2[p]);
    }
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase2(PV (&o)[W * U], T (*a)[NA], const Params&)
    {
#pragma unroll
        for (int p = 0; p < W * U; ++p) {
            a[p % W][0] += o[p].w2;
            a[p % W][1] = rmax_nonneg(a[p % W][1], o[p].w2);
        }
    }
};

// =======================================================================================
// nreg_X -- replaces nreg_Xkernel (nreg_kernels.c:5-65, core nreg_kernels_common.h:7-61);
// 37 flop/pair.  Outputs: mrx mry mrz ax ay az u, with u = im * sum (nreg_kernels.c:63).
// =======================================================================================
template <typename T> struct DtParams { T dt; };
template <typename T> struct NregXOp {
    typedef T real;
    typedef DtParams<T> Params;
    enum { NI = 7, NJ = 8, NA = 7, NO = 7, WPT = 2, UNROLL = 4 };
    enum { NJP = round_up(NJ, Vec16<T>::N) };
    enum { IX, IY, IZ, IE, IVX, IVY, IVZ };
    static TUPAN_DEV void load_i(const T* const* a, long long i, T (&s)[NI])
    {
        s[IX] = a[1][i]; s[IY] = a[2][i]; s[IZ] = a[3][i]; s[IE] = a[4][i];
        s[IVX] = a[5][i]; s[IVY] = a[6][i]; s[IVZ] = a[7][i];
    }
    static TUPAN_DEV void pack_j(const T* const* j, long long r, T (&row)[NJP]) { pack_row8(j, r, row); }
    static TUPAN_DEV void zero(T (&a)[NA]) { zero_all(a); }
    static TUPAN_DEV void pair(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA], const Params& p)
    {
        T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
        T vx = s[IVX] - row[J8_VX], vy = s[IVY] - row[J8_VY], vz = s[IVZ] - row[J8_VZ];
        rx = fma(vx, p.dt, rx); ry = fma(vy, p.dt, ry); rz = fma(vz, p.dt, rz);
        T r2 = rx * rx; r2 = fma(ry, ry, r2); r2 = fma(rz, rz, r2);
        T x = r2 + (s[IE] + row[J8_E2]);
        InvR<T> w = soft_inv<true>(x, r2);
        T mj = row[JM];
        T g = mj * w.r3;
        a[0] = fma(mj, rx, a[0]); a[1] = fma(mj, ry, a[1]); a[2] = fma(mj, rz, a[2]);
        a[3] = fma(-g, rx, a[3]); a[4] = fma(-g, ry, a[4]); a[5] = fma(-g, rz, a[5]);
        a[6] = fma(mj, w.r1, a[6]);
    }
    static TUPAN_DEV void combine(T (&a)[NA], const T (&b)[NA]) { sum_combine(a, b); }
    static TUPAN_DEV void finish(const T* const* ia, long long i, const T (&a)[NA], const Params&, T* const* out)
    {
#pragma unroll
        for (int k = 0; k < 6; ++k) out[k][i] = a[k];
        out[6][i] = ia[0][i] * a[6];
    }

    // ---- grouped form (pair_kernel_grouped, fp64; see AccJerkOp): 28 FP64 instructions per pair instead
    // of 29 (e2_i in the r2 chain, the mask tested per group); the four accumulations that share the row's
    // mass run back to back for the W pairs of a row, then the three that share g for each pair.
    // Measured (profiles/r02_kernel_lab7_nregx.txt, tools/kernel_lab3.cu -DLAB_OP=4): ungrouped 538.2 Gpair/s (69.2 clocks
    // per pair), 4 x 2 583.0 (63.9), 3 x 2 568.8, 2 x 2 572.1 (the second shape, 512 particles per CTA).
#ifndef TUPAN_NREGX_GROUPED
#define TUPAN_NREGX_GROUPED 1
#define TUPAN_NREGX_GW 4
#define TUPAN_NREGX_GU 2
#endif
    enum { GROUPED = (TUPAN_NREGX_GROUPED != 0 && sizeof(T) == 8), GW = TUPAN_NREGX_GW, GU = TUPAN_NREGX_GU, GNT = 256,
           GMODE = 8 };
    enum { GALT = 1, GW2 = 2, GU2 = 2, GMODE2 = 8, GCOST2_PERMILLE = 1019 };
    struct PV { T rx, ry, rz, r1, g, m; };
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase1(const T (*s)[NI], const T (*rows)[NJP], PV (&o)[W * U], const Params& prm, int)
    {
        constexpr int G = W * U;
        T r2[G], x[G], y0[G], t[G], h[G];
#pragma unroll
        for (int p = 0; p < G; ++p) {
            const T(&si)[NI] = s[p % W];
            const T(&rw)[NJP] = rows[p / W];
            o[p].rx = si[IX] - rw[JX]; o[p].ry = si[IY] - rw[JY]; o[p].rz = si[IZ] - rw[JZ];
            t[p] = si[IVX] - rw[J8_VX]; h[p] = si[IVY] - rw[J8_VY]; x[p] = si[IVZ] - rw[J8_VZ];
            o[p].m = rw[JM];
        }
#pragma unroll
        for (int p = 0; p < G; ++p) {
            o[p].rx = fma(t[p], prm.dt, o[p].rx); o[p].ry = fma(h[p], prm.dt, o[p].ry); o[p].rz = fma(x[p], prm.dt, o[p].rz);
        }
#pragma unroll
        for (int p = 0; p < G; ++p) r2[p] = fma(o[p].rx, o[p].rx, s[p % W][IE]);
#pragma unroll
        for (int p = 0; p < G; ++p) r2[p] = fma(o[p].ry, o[p].ry, r2[p]);
#pragma unroll
        for (int p = 0; p < G; ++p) r2[p] = fma(o[p].rz, o[p].rz, r2[p]);
        group_fold_seeds<G>(r2, [&](int p) { return __double2hiint(s[p % W][IE]); }, [&](int p) { return rows[p / W][J8_E2]; },
                            [&](int p) { T q = o[p].rx * o[p].rx; q = fma(o[p].ry, o[p].ry, q); return fma(o[p].rz, o[p].rz, q); },
                            x, y0);
        group_rsqrt_step<G>(x, y0, T(1), T(0.5), T(0.375), t);
#pragma unroll
        for (int p = 0; p < G; ++p) o[p].r1 = t[p];
#pragma unroll
        for (int p = 0; p < G; ++p) h[p] = o[p].r1 * o[p].r1;
#pragma unroll
        for (int p = 0; p < G; ++p) h[p] = h[p] * o[p].r1;
#pragma unroll
        for (int p = 0; p < G; ++p) o[p].g = -(o[p].m * h[p]);
    }
    template <int W, int U, int MODE>
    static TUPAN_DEV void group_phase2(PV (&o)[W * U], T (*a)[NA], const Params&)
    {
#pragma unroll
        for (int p = 0; p < W * U; ++p) {             // p / W = row: m shared by 4 W consecutive DFMAs
            T(&ac)[NA] = a[p % W];
            ac[0] = fma(o[p].rx, o[p].m, ac[0]); ac[1] = fma(o[p].ry, o[p].m, ac[1]); ac[2] = fma(o[p].rz, o[p].m, ac[2]);
            ac[6] = fma(o[p].r1, o[p].m, ac[6]);
        }
#pragma unroll
        for (int p = 0; p < W * U; ++p) {
            T(&ac)[NA] = a[p % W];
            ac[3] = fma(o[p].rx, o[p].g, ac[3]); ac[4] = fma(o[p].ry, o[p].g, ac[4]); ac[5] = fma(o[p].rz, o[p].g, ac[5]);
        }
    }
};

// =======================================================================================
// nreg_V -- replaces nreg_Vkernel (nreg_kernels.c:68-116, core nreg_kernels_common.h:63-102);
// 25 flop/pair; no mask, no 1/r.  Caller arrays: m vx vy vz ax ay az.
// Outputs: mvx mvy mvz mk with mk = im * sum (nreg_kernels.c:113).
// =======================================================================================
template <typename T> struct NregVOp {
    typedef T real;
    typedef DtParams<T> Params;
    enum { NI = 6, NJ = 7, NA = 4, NO = 4, WPT = 4, UNROLL = 4 };
    enum { NJP = round_up(NJ, Vec16<T>::N) };
    enum { IVX, IVY, IVZ, IAX, IAY, IAZ };
    enum { RV_VX = 0, RV_VY = 1, RV_VZ = 2, RV_M = 3, RV_AX = 4, RV_AY = 5, RV_AZ = 6 };
    static TUPAN_DEV void load_i(const T* const* a, long long i, T (&s)[NI])
    {
        s[IVX] = a[1][i]; s[IVY] = a[2][i]; s[IVZ] = a[3][i];
        s[IAX] = a[4][i]; s[IAY] = a[5][i]; s[IAZ] = a[6][i];
    }
    static TUPAN_DEV void pack_j(const T* const* j, long long r, T (&row)[NJP])
    {
        row[RV_M] = j[0][r];
        row[RV_VX] = j[1][r]; row[RV_VY] = j[2][r]; row[RV_VZ] = j[3][r];
        row[RV_AX] = j[4][r]; row[RV_AY] = j[5][r]; row[RV_AZ] = j[6][r];
    }
    static TUPAN_DEV void zero(T (&a)[NA]) { zero_all(a); }
    static TUPAN_DEV void pair(const T (&s)[NI], const T (&row)[NJP], T (&a)[NA], const Params& p)
    {
        T vx = s[IVX] - row[RV_VX], vy = s[IVY] - row[RV_VY], vz = s[IVZ] - row[RV_VZ];
        T ax = s[IAX] - row[RV_AX], ay = s[IAY] - row[RV_AY], az = s[IAZ] - row[RV_AZ];
        vx = fma(ax, p.dt, vx); vy = fma(ay, p.dt, vy); vz = fma(az, p.dt, vz);
        T v2 = vx * vx; v2 = fma(vy, vy, v2); v2 = fma(vz, vz, v2);
        T mj = row[RV_M];
        a[0] = fma(mj, vx, a[0]); a[1] = fma(mj, vy, a[1]); a[2] = fma(mj, vz, a[2]);
        a[3] = fma(mj, v2, a[3]);
    }
    static TUPAN_DEV void combine(T (&a)[NA], const T (&b)[NA]) { sum_combine(a, b); }
    static TUPAN_DEV void finish(const T* const* ia, long long i, const T (&a)[NA], const Params&, T* const* out)
    {
        out[0][i] = a[0]; out[1][i] = a[1]; out[2][i] = a[2];
        out[3][i] = ia[0][i] * a[3];
    }
};

// FP-pipe instructions per pair (launch-plan model, pair_engine.cuh)
template <typename T> struct OpCost<PhiOp<T>> { enum { value = 14 }; };
template <typename T> struct OpCost<AccOp<T>> { enum { value = 19 }; };
template <typename T> struct OpCost<AccJerkOp<T>> { enum { value = 32 }; };
template <typename T> struct OpCost<SnapCrackleOp<T>> { enum { value = 78 }; };
template <typename T> struct OpCost<TstepOp<T>> { enum { value = 39 }; };
template <typename T> struct OpCost<NregXOp<T>> { enum { value = 30 }; };
template <typename T> struct OpCost<NregVOp<T>> { enum { value = 16 }; };

}  // namespace tupan
