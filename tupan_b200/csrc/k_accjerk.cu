// k_accjerk.cu -- acc_jerk on the pair engine, in a translation unit of its own: it is compiled with
// `-Xptxas -regUsageLevel 10` (tupan_b200/build.py), which gives the grouped kernel one non-FP64
// instruction per pair less and +1.2 % (profiles/r02_kernel_lab_grouped.txt, "flags"), while the same
// flag costs the other Newtonian kernels 1 %.
#include "ops.cuh"
#include "runtime.cuh"

namespace tupan {
static inline NoParams no_params_aj(const double*) { return NoParams(); }
TUPAN_DEFINE_VTABLE(vt_acc_jerk, AccJerkOp<real_t>, "acc_jerk_kernel", 8, 6, 0, 42, no_params_aj)
}  // namespace tupan
