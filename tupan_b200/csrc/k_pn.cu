// k_pn.cu -- pnacc on the pair engine; one straight-line variant per PN level.
#include "pn_ops.cuh"
#include "runtime.cuh"

namespace tupan {
// scal = order, inv1 .. inv7  (libtupan.h:121-150)
static inline PNParams<real_t> pn_params(const double* s)
{
    PNParams<real_t> p;
    p.c2 = (real_t)s[2]; p.c4 = (real_t)s[4]; p.c5 = (real_t)s[5]; p.c6 = (real_t)s[6]; p.c7 = (real_t)s[7];
    return p;
}
typedef PNAccOp<real_t, 0> PNLevel0;
typedef PNAccOp<real_t, 2> PNLevel2;
typedef PNAccOp<real_t, 4> PNLevel4;
typedef PNAccOp<real_t, 5> PNLevel5;
typedef PNAccOp<real_t, 6> PNLevel6;
typedef PNAccOp<real_t, 7> PNLevel7;
TUPAN_DEFINE_VTABLE(vt_pn0, PNLevel0, "pnacc_kernel", 8, 3, 8, 0, pn_params)
TUPAN_DEFINE_VTABLE(vt_pn2, PNLevel2, "pnacc_kernel", 8, 3, 8, 121, pn_params)
TUPAN_DEFINE_VTABLE(vt_pn4, PNLevel4, "pnacc_kernel", 8, 3, 8, 193, pn_params)
TUPAN_DEFINE_VTABLE(vt_pn5, PNLevel5, "pnacc_kernel", 8, 3, 8, 209, pn_params)
TUPAN_DEFINE_VTABLE(vt_pn6, PNLevel6, "pnacc_kernel", 8, 3, 8, 461, pn_params)
TUPAN_DEFINE_VTABLE(vt_pn7, PNLevel7, "pnacc_kernel", 8, 3, 8, 632, pn_params)

// order -> level: the reference's nested gates (pn_terms.h:509-546): >1, >3, >4, >5, >6
const KernelVTable* pn_vtable_for_order(double order)
{
    const long long o = (long long)order;
    if (o > 6) return &vt_pn7;
    if (o > 5) return &vt_pn6;
    if (o > 4) return &vt_pn5;
    if (o > 3) return &vt_pn4;
    if (o > 1) return &vt_pn2;
    return &vt_pn0;
}

namespace {
int pn_rw(const double* s) { return vt_pn7.row_width(s); }
int pn_na(const double* s) { return vt_pn7.n_acc(s); }
int pn_host(long long ni, const real_t* const* hi, long long nj, const real_t* const* hj, const double* s,
            real_t* const* ho)
{ return pn_vtable_for_order(s[0])->run_host(ni, hi, nj, hj, s, ho); }
int pn_dev(long long ni, const real_t* const* di, long long nj, const real_t* const* dj, const double* s,
           real_t* const* dout, cudaStream_t st)
{ return pn_vtable_for_order(s[0])->run_dev(ni, di, nj, dj, s, dout, st); }
int pn_pack(long long nj, const real_t* const* dj, const double* s, real_t* packed, cudaStream_t st)
{ return vt_pn7.pack(nj, dj, s, packed, st); }
int pn_slots(long long ni, long long rows, const double* s) { return pn_vtable_for_order(s[0])->sweep_slots(ni, rows, s); }
int pn_sweep(long long ni, const real_t* const* di, const real_t* packed, long long j0, long long j1,
             const double* s, real_t* partial, int slot0, cudaStream_t st)
{ return pn_vtable_for_order(s[0])->sweep(ni, di, packed, j0, j1, s, partial, slot0, st); }
int pn_fin(long long ni, const real_t* const* di, const real_t* partial, int nslots, const double* s,
           real_t* const* dout, cudaStream_t st)
{ return pn_vtable_for_order(s[0])->finalize(ni, di, partial, nslots, s, dout, st); }
int pn_multi(long long ni, const real_t* const* di, int nseg, const real_t* const* sp, const long long* sr,
             const double* s, real_t* partial, int slot0, cudaStream_t st)
{ return pn_vtable_for_order(s[0])->sweep_multi(ni, di, nseg, sp, sr, s, partial, slot0, st); }
}  // namespace
extern const KernelVTable vt_pnacc = {"pnacc_kernel", 8, 3, 8, 632, pn_rw, pn_na, pn_host, pn_dev,
                                      pn_pack, pn_slots, pn_sweep, pn_fin, pn_multi};
}  // namespace tupan
