// abi.cu -- Part 1 of include/libtupan_cuda.h: the ten entry points of the reference's
// tupan/lib/src/libtupan.h:2-246, same names and argument order, host pointers, synchronous.
// Pure marshalling; the work happens in the kernel families behind the vtable.
#include <math.h>

#include "runtime.cuh"
#include "../../include/libtupan_cuda.h"

namespace tupan {
int kepler_run_host(long long pairs, const real_t* const* hin, double dt, real_t* const* hout);
}
using namespace tupan;

// The reference's entry points are void and so are these; a caller that does not look at
// tupan_cuda_last_error() must still not mistake untouched output arrays for a result: on any
// failure every output is filled with NaN.
static void poison(int rc, long long n, REAL* const* out, int n_out)
{
    if (rc == 0 || n <= 0) return;
    const REAL bad = (REAL)NAN;
    for (int k = 0; k < n_out; ++k)
        if (out[k])
            for (long long i = 0; i < n; ++i) out[k][i] = bad;
}

extern "C" {

void phi_kernel(const UINT ni, const REAL* im, const REAL* irx, const REAL* iry, const REAL* irz, const REAL* ie2,
                const UINT nj, const REAL* jm, const REAL* jrx, const REAL* jry, const REAL* jrz, const REAL* je2,
                REAL* iphi)
{
    const REAL* hi[] = {im, irx, iry, irz, ie2};
    const REAL* hj[] = {jm, jrx, jry, jrz, je2};
    REAL* ho[] = {iphi};
    poison(vtable(K_PHI)->run_host((long long)ni, hi, (long long)nj, hj, nullptr, ho), (long long)ni, ho, (int)(sizeof(ho) / sizeof(ho[0])));
}

void acc_kernel(const UINT ni, const REAL* im, const REAL* irx, const REAL* iry, const REAL* irz, const REAL* ie2,
                const UINT nj, const REAL* jm, const REAL* jrx, const REAL* jry, const REAL* jrz, const REAL* je2,
                REAL* iax, REAL* iay, REAL* iaz)
{
    const REAL* hi[] = {im, irx, iry, irz, ie2};
    const REAL* hj[] = {jm, jrx, jry, jrz, je2};
    REAL* ho[] = {iax, iay, iaz};
    poison(vtable(K_ACC)->run_host((long long)ni, hi, (long long)nj, hj, nullptr, ho), (long long)ni, ho, (int)(sizeof(ho) / sizeof(ho[0])));
}

void acc_jerk_kernel(const UINT ni, const REAL* im, const REAL* irx, const REAL* iry, const REAL* irz,
                     const REAL* ie2, const REAL* ivx, const REAL* ivy, const REAL* ivz, const UINT nj,
                     const REAL* jm, const REAL* jrx, const REAL* jry, const REAL* jrz, const REAL* je2,
                     const REAL* jvx, const REAL* jvy, const REAL* jvz, REAL* iax, REAL* iay, REAL* iaz, REAL* ijx,
                     REAL* ijy, REAL* ijz)
{
    const REAL* hi[] = {im, irx, iry, irz, ie2, ivx, ivy, ivz};
    const REAL* hj[] = {jm, jrx, jry, jrz, je2, jvx, jvy, jvz};
    REAL* ho[] = {iax, iay, iaz, ijx, ijy, ijz};
    poison(vtable(K_ACC_JERK)->run_host((long long)ni, hi, (long long)nj, hj, nullptr, ho), (long long)ni, ho, (int)(sizeof(ho) / sizeof(ho[0])));
}

void snap_crackle_kernel(const UINT ni, const REAL* im, const REAL* irx, const REAL* iry, const REAL* irz,
                         const REAL* ie2, const REAL* ivx, const REAL* ivy, const REAL* ivz, const REAL* iax,
                         const REAL* iay, const REAL* iaz, const REAL* ijx, const REAL* ijy, const REAL* ijz,
                         const UINT nj, const REAL* jm, const REAL* jrx, const REAL* jry, const REAL* jrz,
                         const REAL* je2, const REAL* jvx, const REAL* jvy, const REAL* jvz, const REAL* jax,
                         const REAL* jay, const REAL* jaz, const REAL* jjx, const REAL* jjy, const REAL* jjz,
                         REAL* isx, REAL* isy, REAL* isz, REAL* icx, REAL* icy, REAL* icz)
{
    const REAL* hi[] = {im, irx, iry, irz, ie2, ivx, ivy, ivz, iax, iay, iaz, ijx, ijy, ijz};
    const REAL* hj[] = {jm, jrx, jry, jrz, je2, jvx, jvy, jvz, jax, jay, jaz, jjx, jjy, jjz};
    REAL* ho[] = {isx, isy, isz, icx, icy, icz};
    poison(vtable(K_SNAP_CRACKLE)->run_host((long long)ni, hi, (long long)nj, hj, nullptr, ho), (long long)ni, ho, (int)(sizeof(ho) / sizeof(ho[0])));
}

void tstep_kernel(const UINT ni, const REAL* im, const REAL* irx, const REAL* iry, const REAL* irz, const REAL* ie2,
                  const REAL* ivx, const REAL* ivy, const REAL* ivz, const UINT nj, const REAL* jm, const REAL* jrx,
                  const REAL* jry, const REAL* jrz, const REAL* je2, const REAL* jvx, const REAL* jvy,
                  const REAL* jvz, const REAL eta, REAL* idt_a, REAL* idt_b)
{
    const REAL* hi[] = {im, irx, iry, irz, ie2, ivx, ivy, ivz};
    const REAL* hj[] = {jm, jrx, jry, jrz, je2, jvx, jvy, jvz};
    REAL* ho[] = {idt_a, idt_b};
    const double scal[] = {(double)eta};
    poison(vtable(K_TSTEP)->run_host((long long)ni, hi, (long long)nj, hj, scal, ho), (long long)ni, ho, (int)(sizeof(ho) / sizeof(ho[0])));
}

void pnacc_kernel(const UINT ni, const REAL* im, const REAL* irx, const REAL* iry, const REAL* irz, const REAL* ie2,
                  const REAL* ivx, const REAL* ivy, const REAL* ivz, const UINT nj, const REAL* jm, const REAL* jrx,
                  const REAL* jry, const REAL* jrz, const REAL* je2, const REAL* jvx, const REAL* jvy,
                  const REAL* jvz, UINT order, const REAL inv1, const REAL inv2, const REAL inv3, const REAL inv4,
                  const REAL inv5, const REAL inv6, const REAL inv7, REAL* ipnax, REAL* ipnay, REAL* ipnaz)
{
    const REAL* hi[] = {im, irx, iry, irz, ie2, ivx, ivy, ivz};
    const REAL* hj[] = {jm, jrx, jry, jrz, je2, jvx, jvy, jvz};
    REAL* ho[] = {ipnax, ipnay, ipnaz};
    const double scal[] = {(double)order, (double)inv1, (double)inv2, (double)inv3,
                           (double)inv4,  (double)inv5, (double)inv6, (double)inv7};
    poison(vtable(K_PNACC)->run_host((long long)ni, hi, (long long)nj, hj, scal, ho), (long long)ni, ho, (int)(sizeof(ho) / sizeof(ho[0])));
}

void nreg_Xkernel(const UINT ni, const REAL* im, const REAL* irx, const REAL* iry, const REAL* irz, const REAL* ie2,
                  const REAL* ivx, const REAL* ivy, const REAL* ivz, const UINT nj, const REAL* jm, const REAL* jrx,
                  const REAL* jry, const REAL* jrz, const REAL* je2, const REAL* jvx, const REAL* jvy,
                  const REAL* jvz, const REAL dt, REAL* idrx, REAL* idry, REAL* idrz, REAL* iax, REAL* iay,
                  REAL* iaz, REAL* iu)
{
    const REAL* hi[] = {im, irx, iry, irz, ie2, ivx, ivy, ivz};
    const REAL* hj[] = {jm, jrx, jry, jrz, je2, jvx, jvy, jvz};
    REAL* ho[] = {idrx, idry, idrz, iax, iay, iaz, iu};
    const double scal[] = {(double)dt};
    poison(vtable(K_NREG_X)->run_host((long long)ni, hi, (long long)nj, hj, scal, ho), (long long)ni, ho, (int)(sizeof(ho) / sizeof(ho[0])));
}

void nreg_Vkernel(const UINT ni, const REAL* im, const REAL* ivx, const REAL* ivy, const REAL* ivz, const REAL* iax,
                  const REAL* iay, const REAL* iaz, const UINT nj, const REAL* jm, const REAL* jvx, const REAL* jvy,
                  const REAL* jvz, const REAL* jax, const REAL* jay, const REAL* jaz, const REAL dt, REAL* idvx,
                  REAL* idvy, REAL* idvz, REAL* ik)
{
    const REAL* hi[] = {im, ivx, ivy, ivz, iax, iay, iaz};
    const REAL* hj[] = {jm, jvx, jvy, jvz, jax, jay, jaz};
    REAL* ho[] = {idvx, idvy, idvz, ik};
    const double scal[] = {(double)dt};
    poison(vtable(K_NREG_V)->run_host((long long)ni, hi, (long long)nj, hj, scal, ho), (long long)ni, ho, (int)(sizeof(ho) / sizeof(ho[0])));
}

void sakura_kernel(const UINT ni, const REAL* im, const REAL* irx, const REAL* iry, const REAL* irz,
                   const REAL* ie2, const REAL* ivx, const REAL* ivy, const REAL* ivz, const UINT nj, const REAL* jm,
                   const REAL* jrx, const REAL* jry, const REAL* jrz, const REAL* je2, const REAL* jvx,
                   const REAL* jvy, const REAL* jvz, const REAL dt, const INT flag, REAL* idrx, REAL* idry,
                   REAL* idrz, REAL* idvx, REAL* idvy, REAL* idvz)
{
    const REAL* hi[] = {im, irx, iry, irz, ie2, ivx, ivy, ivz};
    const REAL* hj[] = {jm, jrx, jry, jrz, je2, jvx, jvy, jvz};
    REAL* ho[] = {idrx, idry, idrz, idvx, idvy, idvz};
    const double scal[] = {(double)dt, (double)flag};
    poison(vtable(K_SAKURA)->run_host((long long)ni, hi, (long long)nj, hj, scal, ho), (long long)ni, ho, (int)(sizeof(ho) / sizeof(ho[0])));
}

void kepler_solver_kernel(const REAL* im, const REAL* irx, const REAL* iry, const REAL* irz, const REAL* ie2,
                          const REAL* ivx, const REAL* ivy, const REAL* ivz, const REAL dt, REAL* ir1x, REAL* ir1y,
                          REAL* ir1z, REAL* iv1x, REAL* iv1y, REAL* iv1z)
{
    const REAL* hin[] = {im, irx, iry, irz, ie2, ivx, ivy, ivz};
    REAL* hout[] = {ir1x, ir1y, ir1z, iv1x, iv1y, iv1z};
    poison(kepler_run_host(1, hin, (double)dt, hout), 2, hout, 6);
}

}  // extern "C"
