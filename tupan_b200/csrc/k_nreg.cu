// k_nreg.cu -- nreg_X, nreg_V on the pair engine.
#include "ops.cuh"
#include "runtime.cuh"

namespace tupan {
static inline DtParams<real_t> dt_params(const double* s)
{
    DtParams<real_t> p;
    p.dt = (real_t)s[0];
    return p;
}
TUPAN_DEFINE_VTABLE(vt_nreg_x, NregXOp<real_t>, "nreg_Xkernel", 8, 7, 1, 37, dt_params)
TUPAN_DEFINE_VTABLE(vt_nreg_v, NregVOp<real_t>, "nreg_Vkernel", 7, 4, 1, 25, dt_params)
}  // namespace tupan
