// k_newton.cu -- phi, acc, tstep on the pair engine (acc_jerk: k_accjerk.cu).
#include "ops.cuh"
#include "runtime.cuh"

namespace tupan {
static inline NoParams no_params(const double*) { return NoParams(); }
static inline TstepParams<real_t> tstep_params(const double* s)
{
    TstepParams<real_t> p;
    p.eta = (real_t)s[0];
    p.eta_k1 = p.eta * (real_t)0.5;
    p.eta_k2 = p.eta * (real_t)0.375;
    return p;
}
TUPAN_DEFINE_VTABLE(vt_phi, PhiOp<real_t>, "phi_kernel", 5, 1, 0, 14, no_params)
TUPAN_DEFINE_VTABLE(vt_acc, AccOp<real_t>, "acc_kernel", 5, 3, 0, 20, no_params)
TUPAN_DEFINE_VTABLE(vt_tstep, TstepOp<real_t>, "tstep_kernel", 8, 2, 1, 42, tstep_params)
}  // namespace tupan
