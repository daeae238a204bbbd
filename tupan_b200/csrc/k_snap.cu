// k_snap.cu -- snap_crackle on the pair engine.
#include "ops.cuh"
#include "runtime.cuh"

namespace tupan {
static inline NoParams no_params(const double*) { return NoParams(); }
TUPAN_DEFINE_VTABLE(vt_snap_crackle, SnapCrackleOp<real_t>, "snap_crackle_kernel", 14, 6, 0, 114, no_params)
}  // namespace tupan
