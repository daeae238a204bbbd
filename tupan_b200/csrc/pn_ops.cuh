// pn_ops.cuh -- post-Newtonian pair acceleration (1PN, 2PN, 2.5PN, 3PN, 3.5PN) for the pair
// engine.  Replaces pnacc_kernel (pnacc_kernel.c:5-61, core pnacc_kernel_common.h:8-70) and
// the pair terms of pn_terms.h:126-567.
//
//   a_i += sum_j  mj/r^3 * A * r_ij  +  mj/r^2 * B * v_ij ,     A = sum_k A_k, B = sum_k B_k
//
// with k over the enabled orders.  The reference gates the orders with nested
// `order > 1, > 3, > 4, > 5, > 6` tests (pn_terms.h:509-546; orders 1 and 3 add nothing);
// here the gate is the template parameter LEVEL in {0, 2, 4, 5, 6, 7}, chosen on the host,
// so each variant is straight-line code.
//
// WHAT THIS FILE IS: the polynomials pn_1 .. pn_35 below RESTATE the reference's p2p_pn2 .. p2p_pn7
// (pn_terms.h:126-470) term by term -- the same coefficients (those of the harmonic-coordinate
// two-body equations of motion through 3.5PN, Blanchet, Living Rev. Relativity, general frame,
// n = r_ij/r, vi, vj the two velocities, v = vi - vj), in the reference author's factorisation and
// term order, with the scalar products renamed (njv2 -> nj2, ivjv -> vij, ...).  They were written
// next to the reference source so that every term could be checked against it; they are not an
// independent derivation.  What is new here is what surrounds them: one compiled variant per PN
// level instead of run-time gates, the scalar products formed once per pair with FMAs, the masked
// rsqrt of common.cuh, and the pair engine.  A re-derivation that shares sub-polynomials between
// A and B and keeps the ~90 rational constants out of UMOV (107 of them per pair at level 7,
// profiles/r01_static_instruction_mix_fp64.txt) is the open performance item of this kernel.
//
// flops/pair (binary operations counted in the reference source): 33 + 72 +
// {16, 72, 16, 252, 171} for {1PN, 2PN, 2.5PN, 3PN, 3.5PN} -> 632 at order 7.
#pragma once
#include "ops.cuh"

namespace tupan {

template <typename T> struct PNParams { T c2, c4, c5, c6, c7; };

template <typename T> TUPAN_DEV constexpr T fr(int n, int d) { return T(n) / T(d); }

// The rational coefficients whose denominators are not powers of two (x/3, /5, /7, /15, /21, /35,
// /105 ...) need all 64 bits: as literals ptxas materialises each of them with two UMOVs into a
// uniform register, per pair (107 UMOVs per pair at level 7, profiles/r01_static_instruction_mix_fp64.txt).
// As __constant__ objects they are constant-bank operands of the DFMA / DMUL that uses them.
template <int N, int D> __constant__ double pnk_d = double(N) / double(D);
template <int N, int D> __constant__ float pnk_f = float(N) / float(D);
template <typename T, int N, int D> struct PNK;
template <int N, int D> struct PNK<double, N, D> { static TUPAN_DEV double v() { return pnk_d<N, D>; } };
template <int N, int D> struct PNK<float, N, D> { static TUPAN_DEV float v() { return pnk_f<N, D>; } };
#define FRK(n, d) (PNK<T, (n), (d)>::v())

template <typename T> struct PNScalars {
    T mi, mj, mi2, mj2, mimj, ir, ir2;
    T v2, vi2, vj2, vi4, vj4, vij, vij2;
    T nv, nv2, ni, nj, ni2, nj2, ninj;
};

// ---- 1PN --------------------------------------------------------------------------------
template <typename T> TUPAN_DEV void pn_1(const PNScalars<T>& q, T c2, T& A, T& B)
{
    T a = -q.vi2 - T(2) * q.vj2 + T(4) * q.vij + fr<T>(3, 2) * q.nj2 + q.ir * (T(5) * q.mi + T(4) * q.mj);
    T b = T(4) * q.ni - T(3) * q.nj;
    A += a * c2;
    B += b * c2;
}

// ---- 2PN --------------------------------------------------------------------------------
template <typename T> TUPAN_DEV void pn_2(const PNScalars<T>& q, T c4, T& A, T& B)
{
    T a = T(-2) * (q.vj4 + q.vij2) + T(4) * q.vj2 * q.vij
        + q.nj2 * (fr<T>(3, 2) * q.vi2 + fr<T>(9, 2) * q.vj2 - T(6) * q.vij - fr<T>(15, 8) * q.nj2)
        - q.ir2 * (fr<T>(57, 4) * q.mi2 + T(9) * q.mj2 + fr<T>(69, 2) * q.mimj)
        + q.ir * (q.mi * (-fr<T>(15, 4) * q.vi2 + fr<T>(5, 4) * q.vj2 - fr<T>(5, 2) * q.vij
                          + fr<T>(39, 2) * q.ni2 - T(39) * q.ninj + fr<T>(17, 2) * q.nj2)
                  + q.mj * (T(4) * q.vj2 - T(8) * q.vij + T(2) * q.ni2 - T(4) * q.ninj - T(6) * q.nj2));
    T b = q.ir * (q.mi * (-fr<T>(63, 4) * q.ni + fr<T>(55, 4) * q.nj) - q.mj * T(2) * (q.ni + q.nj))
        + q.vi2 * q.nj - q.vij * q.nv * T(4)
        + q.vj2 * (T(4) * q.ni - T(5) * q.nj)
        + q.nj2 * (T(-6) * q.ni + fr<T>(9, 2) * q.nj);
    A += a * c4;
    B += b * c4;
}

// ---- 2.5PN ------------------------------------------------------------------------------
template <typename T> TUPAN_DEV void pn_25(const PNScalars<T>& q, T c5, T& A, T& B)
{
    T a = q.nv * (q.ir * (-FRK(24, 5) * q.mi + FRK(208, 15) * q.mj) + FRK(12, 5) * q.v2);
    T b = -q.v2 + q.ir * (FRK(8, 5) * q.mi - FRK(32, 5) * q.mj);
    T s = (q.mi * q.ir) * c5;
    A += a * s;
    B += b * s;
}

// ---- 3PN --------------------------------------------------------------------------------
template <typename T> TUPAN_DEV void pn_3(const PNScalars<T>& q, T c6, T& A, T& B)
{
    const T PI2 = T(9.869604401089358);
    const T mi = q.mi, mj = q.mj, mi2 = q.mi2, mj2 = q.mj2, mimj = q.mimj, ir = q.ir, ir2 = q.ir2;
    const T v2 = q.v2, vi2 = q.vi2, vj2 = q.vj2, vj4 = q.vj4, vij = q.vij, vij2 = q.vij2;
    const T nv = q.nv, nv2 = q.nv2, ni = q.ni, nj = q.nj, ni2 = q.ni2, nj2 = q.nj2, ninj = q.ninj;

    T a = nj2 * (T(3) * vij2 + fr<T>(3, 2) * vi2 * vj2 - T(12) * vij * vj2 + fr<T>(15, 2) * vj4
                 + nj2 * (fr<T>(15, 2) * (vij - vj2 - fr<T>(1, 4) * vi2) + fr<T>(35, 16) * nj2))
        + T(2) * vj2 * (-vij2 + vj2 * (T(2) * vij - vj2))
        + mi * ir * (ni * (nj * (T(244) * vij - fr<T>(205, 2) * vi2 - fr<T>(283, 2) * vj2 + fr<T>(383, 2) * nj2)
                           + ni * (fr<T>(229, 4) * (vi2 + vj2 - T(2) * vij) - fr<T>(723, 4) * nj2
                                   + ni * (fr<T>(171, 2) * (nj - fr<T>(1, 4) * ni))))
                     + nj2 * (fr<T>(191, 4) * vi2 + fr<T>(259, 4) * vj2 - fr<T>(225, 2) * vij - fr<T>(455, 8) * nj2)
                     + vij * (fr<T>(91, 2) * vi2 + T(43) * vj2 - fr<T>(177, 4) * vij)
                     - fr<T>(91, 8) * vi2 * (vi2 + T(2) * vj2)
                     - fr<T>(81, 8) * vj4)
        + mj * ir * T(4) * (vj4
                            + nj * (ni * (vij - vj2)
                                    + nj * (T(3) * (vij - vj2) - fr<T>(3, 2) * ni2
                                            + nj * (T(3) * ni + fr<T>(3, 2) * nj)))
                            + vij * (vij - T(2) * vj2))
        + mj2 * ir2 * (-ni2 + T(2) * ninj + fr<T>(43, 2) * nj2 + T(18) * vij - T(9) * vj2)
        + mimj * ir2 * (fr<T>(415, 8) * ni2 - fr<T>(375, 4) * ninj + fr<T>(1113, 8) * nj2 + T(18) * vi2
                        + PI2 * (fr<T>(123, 64) * v2 - fr<T>(615, 64) * nv2)
                        + T(33) * (vij - fr<T>(1, 2) * vj2))
        + mi2 * ir2 * (-fr<T>(2069, 8) * ni2 + T(543) * ninj - fr<T>(939, 4) * nj2 + fr<T>(471, 8) * vi2
                       + fr<T>(357, 8) * (vj2 - T(2) * vij))
        + ir * ir2 * (T(16) * mj * mj2
                      + mi2 * mj * (FRK(547, 3) - fr<T>(41, 16) * PI2)
                      - FRK(13, 12) * mi * mi2
                      + mi * mj2 * (FRK(545, 3) - fr<T>(41, 16) * PI2));

    T b = nj * (vj2 * (vi2 + T(8) * vij - T(7) * vj2) - T(2) * vij2
                + nj * (T(6) * ni * (vij - T(2) * vj2)
                        + nj * (T(6) * (T(2) * vj2 - vij - fr<T>(1, 4) * vi2)
                                + nj * (fr<T>(15, 2) * (ni - fr<T>(3, 4) * nj)))))
        + T(4) * ni * (vj4 - vij * vj2)
        + mj * ir * (nj * (T(4) * (vij - vj2 - fr<T>(1, 2) * ni2) + nj * (T(2) * (T(4) * ni + nj)))
                     + T(2) * ni * (vij - vj2))
        + mi * ir * (ni * (fr<T>(207, 8) * vi2 + fr<T>(81, 8) * vj2 - T(36) * vij - fr<T>(269, 4) * nj2
                           + ni * (fr<T>(565, 4) * nj - fr<T>(243, 4) * ni))
                     + nj * (fr<T>(83, 8) * vj2 + fr<T>(27, 4) * vij - fr<T>(137, 8) * vi2 - FRK(95, 12) * nj2))
        + ir2 * (mj2 * (T(4) * ni + T(5) * nj)
                 + mi2 * (fr<T>(311, 4) * ni - fr<T>(357, 4) * nj)
                 + mimj * (fr<T>(479, 8) * nj - fr<T>(307, 8) * ni + fr<T>(123, 32) * PI2 * nv));
    A += a * c6;
    B += b * c6;
}

// ---- 3.5PN ------------------------------------------------------------------------------
template <typename T> TUPAN_DEV void pn_35(const PNScalars<T>& q, T c7, T& A, T& B)
{
    const T mi = q.mi, mj = q.mj, mi2 = q.mi2, mj2 = q.mj2, mimj = q.mimj, ir = q.ir, ir2 = q.ir2;
    const T v2 = q.v2, vi2 = q.vi2, vj2 = q.vj2, vi4 = q.vi4, vj4 = q.vj4, vij = q.vij;
    const T nv = q.nv, nv2 = q.nv2, ni = q.ni, nj = q.nj, ni2 = q.ni2, nj2 = q.nj2, ninj = q.ninj;

    T a = mi2 * ir2 * (FRK(3992, 105) * ni - FRK(4328, 105) * nj)
        + mimj * ir * ir2 * (-FRK(13576, 105) * ni + FRK(2872, 21) * nj)
        + mj2 * ir * ir2 * (-FRK(3172, 21) * nv)
        + mi * ir * (ni * (T(48) * ni2 - FRK(4888, 105) * vi2 + FRK(2056, 21) * vij - FRK(1028, 21) * vj2)
                     + ninj * (-FRK(696, 5) * ni + FRK(744, 5) * nj)
                     + nj * (-FRK(288, 5) * nj2 + FRK(5056, 105) * vi2 - FRK(2224, 21) * vij
                             + FRK(5812, 105) * vj2))
        + mj * ir * (ni * (-FRK(582, 5) * ni2 - FRK(2864, 35) * vij + FRK(1432, 35) * vj2)
                     + ninj * (FRK(1746, 5) * ni - FRK(1954, 5) * nj)
                     + FRK(3568, 105) * nv * vi2
                     + nj * (T(158) * nj2 - FRK(5752, 105) * vj2 + FRK(10048, 105) * vij))
        + (nv * (T(-56) * nv2 * nv2 - FRK(246, 35) * vi4)
           + ni * (v2 * (T(60) * ni2 - T(180) * ninj + T(174) * nj2)
                   + vij * (FRK(1068, 35) * (vi2 - vij) + FRK(984, 35) * vj2)
                   - FRK(534, 35) * vi2 * vj2 - FRK(204, 35) * vj4)
           + nj * (T(-54) * nj2 * v2
                   + vij * (-FRK(984, 35) * vi2 + FRK(180, 7) * vij - FRK(732, 35) * vj2)
                   + FRK(90, 7) * vi2 * vj2 + FRK(24, 7) * vj4));

    T b = -mi2 * ir2 * FRK(184, 21) + mimj * ir2 * FRK(6224, 105) + mj2 * ir2 * FRK(6388, 105)
        + mi * ir * (FRK(52, 15) * ni2 - FRK(56, 15) * ninj - FRK(44, 15) * nj2 - FRK(132, 35) * vi2
                     + FRK(152, 35) * vij - FRK(48, 35) * vj2)
        + mj * ir * (FRK(454, 15) * ni2 - FRK(372, 5) * ninj + FRK(854, 15) * nj2 - FRK(152, 21) * vi2
                     + FRK(2864, 105) * vij - FRK(1768, 105) * vj2)
        + (T(60) * nv2 * nv2 + v2 * (-FRK(348, 5) * ni2 + FRK(684, 5) * ninj - T(66) * nj2)
           + FRK(334, 35) * vi4
           + vij * (-FRK(1336, 35) * vi2 + FRK(1308, 35) * vij - FRK(1252, 35) * vj2)
           + FRK(654, 35) * vi2 * vj2 + FRK(292, 35) * vj4);
    T s = (mi * ir) * c7;
    A += a * s;
    B += b * s;
}

// =======================================================================================
// The Op.  Caller arrays: m rx ry rz e2 vx vy vz; packed rows: Row8.
// =======================================================================================
template <typename T, int LEVEL> struct PNAccOp {
    typedef T real;
    typedef PNParams<T> Params;
    enum { NI = 8, NJ = 8, NA = 3, NO = 3, WPT = 1, UNROLL = 1 };
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
    static TUPAN_DEV void pair(const T (&s)[NI], const T (&row)[NJP], T (&acc)[NA], const Params& p)
    {
        if (LEVEL < 2) return;
        T rx = s[IX] - row[JX], ry = s[IY] - row[JY], rz = s[IZ] - row[JZ];
        T vx = s[IVX] - row[J8_VX], vy = s[IVY] - row[J8_VY], vz = s[IVZ] - row[J8_VZ];
        T r2 = rx * rx; r2 = fma(ry, ry, r2); r2 = fma(rz, rz, r2);
        T x = r2 + (s[IE] + row[J8_E2]);
        InvR<T> w = soft_inv<true>(x, r2);

        PNScalars<T> q;
        q.mi = s[IM]; q.mj = row[JM]; q.ir = w.r1; q.ir2 = w.r2;
        const T nx = rx * w.r1, ny = ry * w.r1, nz = rz * w.r1;
        const T uix = s[IVX], uiy = s[IVY], uiz = s[IVZ];
        const T ujx = row[J8_VX], ujy = row[J8_VY], ujz = row[J8_VZ];
        q.v2 = fma(vz, vz, fma(vy, vy, vx * vx));
        q.vi2 = fma(uiz, uiz, fma(uiy, uiy, uix * uix));
        q.vj2 = fma(ujz, ujz, fma(ujy, ujy, ujx * ujx));
        q.vij = fma(uiz, ujz, fma(uiy, ujy, uix * ujx));
        q.ni = fma(nz, uiz, fma(ny, uiy, nx * uix));
        q.nj = fma(nz, ujz, fma(ny, ujy, nx * ujx));
        q.nv = fma(nz, vz, fma(ny, vy, nx * vx));
        q.nj2 = q.nj * q.nj;
        q.ni2 = q.ni * q.ni;
        q.ninj = q.ni * q.nj;
        q.nv2 = q.nv * q.nv;
        q.mi2 = q.mi * q.mi; q.mj2 = q.mj * q.mj; q.mimj = q.mi * q.mj;
        q.vi4 = q.vi2 * q.vi2; q.vj4 = q.vj2 * q.vj2; q.vij2 = q.vij * q.vij;

        T A = T(0), B = T(0);          // summed highest order first, as pn_terms.h:560-561
        if (LEVEL >= 7) pn_35(q, p.c7, A, B);
        if (LEVEL >= 6) pn_3(q, p.c6, A, B);
        if (LEVEL >= 5) pn_25(q, p.c5, A, B);
        if (LEVEL >= 4) pn_2(q, p.c4, A, B);
        pn_1(q, p.c2, A, B);
        A *= q.mj * w.r3;
        B *= q.mj * w.r2;
        acc[0] += fma(A, rx, B * vx);
        acc[1] += fma(A, ry, B * vy);
        acc[2] += fma(A, rz, B * vz);
    }
    static TUPAN_DEV void combine(T (&a)[NA], const T (&b)[NA]) { sum_combine(a, b); }
    static TUPAN_DEV void finish(const T* const*, long long i, const T (&a)[NA], const Params&, T* const* out)
    {
        out[0][i] = a[0]; out[1][i] = a[1]; out[2][i] = a[2];
    }
};

// FP-pipe instructions per pair (launch-plan model): ~45 for the Newtonian frame + ~60 per level
template <typename T, int LEVEL> struct OpCost<PNAccOp<T, LEVEL>> { enum { value = 45 + 60 * LEVEL }; };

}  // namespace tupan
