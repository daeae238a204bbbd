/*
 * libtupan_cuda.h -- C ABI of the B200-native tupan kernel libraries.
 *
 * Two shared objects with IDENTICAL symbol names, one per precision, exactly like the
 * reference's cffi build (tupan/lib/cffi_backend.py:35-36,84-87):
 *
 *     libtupan_cuda_fp64.so   REAL = double, UINT = unsigned long, INT = long   (TUPAN_FP64)
 *     libtupan_cuda_fp32.so   REAL = float,  UINT = unsigned int,  INT = int
 *
 * Part 1 is the drop-in: the ten entry points of tupan/lib/src/libtupan.h:2-246 with the
 * same names, argument order and meaning -- `ni, <i arrays>, nj, <j arrays>, <scalars>,
 * <output arrays>`, caller-owned HOST buffers of length >= ni / nj, outputs valid on
 * return.  A maintainer can point tupan's cffi loader at these libraries unchanged (see
 * INTEGRATION.md).  The i and j arrays may be the same arrays; kepler's outputs may alias
 * its inputs (tupan/lib/extensions.py:642-646).
 *
 * Part 2 is what the reference has no equivalent of: device-resident entry points (device
 * pointers + a CUDA stream, asynchronous), the building blocks of the multi-GPU path, and
 * status / timing queries.  The reference's functions are `void` with no error channel
 * (SURVEY.md 8b); Part-1 functions keep that signature, report failures on stderr and
 * record them for tupan_cuda_last_error().
 *
 * There is no CPU fallback anywhere behind this header.
 */
#ifndef LIBTUPAN_CUDA_H
#define LIBTUPAN_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#ifdef TUPAN_FP64
typedef double REAL;
typedef unsigned long UINT;
typedef long INT;
#else
typedef float REAL;
typedef unsigned int UINT;
typedef int INT;
#endif

/* ------------------------------------------------------------------------------------ */
/* Part 1 -- drop-in for tupan/lib/src/libtupan.h                                        */
/* ------------------------------------------------------------------------------------ */

/* replaces libtupan.h:2-15 (phi_kernel.c:5-35).  phi_i = -sum_j m_j / sqrt(r^2 + e2) */
void phi_kernel(const UINT ni,
                const REAL *im, const REAL *irx, const REAL *iry, const REAL *irz, const REAL *ie2,
                const UINT nj,
                const REAL *jm, const REAL *jrx, const REAL *jry, const REAL *jrz, const REAL *je2,
                REAL *iphi);

/* replaces libtupan.h:17-32 (acc_kernel.c:5-41) */
void acc_kernel(const UINT ni,
                const REAL *im, const REAL *irx, const REAL *iry, const REAL *irz, const REAL *ie2,
                const UINT nj,
                const REAL *jm, const REAL *jrx, const REAL *jry, const REAL *jrz, const REAL *je2,
                REAL *iax, REAL *iay, REAL *iaz);

/* replaces libtupan.h:34-58 (acc_jerk_kernel.c:5-61) -- the headline kernel */
void acc_jerk_kernel(const UINT ni,
                     const REAL *im, const REAL *irx, const REAL *iry, const REAL *irz, const REAL *ie2,
                     const REAL *ivx, const REAL *ivy, const REAL *ivz,
                     const UINT nj,
                     const REAL *jm, const REAL *jrx, const REAL *jry, const REAL *jrz, const REAL *je2,
                     const REAL *jvx, const REAL *jvy, const REAL *jvz,
                     REAL *iax, REAL *iay, REAL *iaz, REAL *ijx, REAL *ijy, REAL *ijz);

/* replaces libtupan.h:60-96 (snap_crackle_kernel.c:5-83) */
void snap_crackle_kernel(const UINT ni,
                         const REAL *im, const REAL *irx, const REAL *iry, const REAL *irz, const REAL *ie2,
                         const REAL *ivx, const REAL *ivy, const REAL *ivz,
                         const REAL *iax, const REAL *iay, const REAL *iaz,
                         const REAL *ijx, const REAL *ijy, const REAL *ijz,
                         const UINT nj,
                         const REAL *jm, const REAL *jrx, const REAL *jry, const REAL *jrz, const REAL *je2,
                         const REAL *jvx, const REAL *jvy, const REAL *jvz,
                         const REAL *jax, const REAL *jay, const REAL *jaz,
                         const REAL *jjx, const REAL *jjy, const REAL *jjz,
                         REAL *isx, REAL *isy, REAL *isz, REAL *icx, REAL *icy, REAL *icz);

/* replaces libtupan.h:98-119 (tstep_kernel.c:5-50).  idt_a = eta/sqrt(1+sum w2),
 * idt_b = eta/sqrt(1+max w2) */
void tstep_kernel(const UINT ni,
                  const REAL *im, const REAL *irx, const REAL *iry, const REAL *irz, const REAL *ie2,
                  const REAL *ivx, const REAL *ivy, const REAL *ivz,
                  const UINT nj,
                  const REAL *jm, const REAL *jrx, const REAL *jry, const REAL *jrz, const REAL *je2,
                  const REAL *jvx, const REAL *jvy, const REAL *jvz,
                  const REAL eta,
                  REAL *idt_a, REAL *idt_b);

/* replaces libtupan.h:121-150 (pnacc_kernel.c:5-61); inv_k = clight^-k */
void pnacc_kernel(const UINT ni,
                  const REAL *im, const REAL *irx, const REAL *iry, const REAL *irz, const REAL *ie2,
                  const REAL *ivx, const REAL *ivy, const REAL *ivz,
                  const UINT nj,
                  const REAL *jm, const REAL *jrx, const REAL *jry, const REAL *jrz, const REAL *je2,
                  const REAL *jvx, const REAL *jvy, const REAL *jvz,
                  UINT order,
                  const REAL inv1, const REAL inv2, const REAL inv3, const REAL inv4,
                  const REAL inv5, const REAL inv6, const REAL inv7,
                  REAL *ipnax, REAL *ipnay, REAL *ipnaz);

/* replaces libtupan.h:152-178 (nreg_kernels.c:5-65) */
void nreg_Xkernel(const UINT ni,
                  const REAL *im, const REAL *irx, const REAL *iry, const REAL *irz, const REAL *ie2,
                  const REAL *ivx, const REAL *ivy, const REAL *ivz,
                  const UINT nj,
                  const REAL *jm, const REAL *jrx, const REAL *jry, const REAL *jrz, const REAL *je2,
                  const REAL *jvx, const REAL *jvy, const REAL *jvz,
                  const REAL dt,
                  REAL *idrx, REAL *idry, REAL *idrz, REAL *iax, REAL *iay, REAL *iaz, REAL *iu);

/* replaces libtupan.h:180-201 (nreg_kernels.c:68-116) */
void nreg_Vkernel(const UINT ni,
                  const REAL *im, const REAL *ivx, const REAL *ivy, const REAL *ivz,
                  const REAL *iax, const REAL *iay, const REAL *iaz,
                  const UINT nj,
                  const REAL *jm, const REAL *jvx, const REAL *jvy, const REAL *jvz,
                  const REAL *jax, const REAL *jay, const REAL *jaz,
                  const REAL dt,
                  REAL *idvx, REAL *idvy, REAL *idvz, REAL *ik);

/* replaces libtupan.h:203-229 (sakura_kernel.c:5-64); flag in {-2,-1,1,2}, else no-op */
void sakura_kernel(const UINT ni,
                   const REAL *im, const REAL *irx, const REAL *iry, const REAL *irz, const REAL *ie2,
                   const REAL *ivx, const REAL *ivy, const REAL *ivz,
                   const UINT nj,
                   const REAL *jm, const REAL *jrx, const REAL *jry, const REAL *jrz, const REAL *je2,
                   const REAL *jvx, const REAL *jvy, const REAL *jvz,
                   const REAL dt, const INT flag,
                   REAL *idrx, REAL *idry, REAL *idrz, REAL *idvx, REAL *idvy, REAL *idvz);

/* replaces libtupan.h:231-246 (kepler_solver_kernel.c:5-50); exactly two bodies */
void kepler_solver_kernel(const REAL *im, const REAL *irx, const REAL *iry, const REAL *irz, const REAL *ie2,
                          const REAL *ivx, const REAL *ivy, const REAL *ivz,
                          const REAL dt,
                          REAL *ir1x, REAL *ir1y, REAL *ir1z, REAL *iv1x, REAL *iv1y, REAL *iv1z);

/* ------------------------------------------------------------------------------------ */
/* Part 2 -- device-resident API, multi-GPU building blocks, status                       */
/* ------------------------------------------------------------------------------------ */

/* kernel ids for the generic entry points */
enum tupan_kernel {
    TUPAN_PHI = 0, TUPAN_ACC = 1, TUPAN_ACC_JERK = 2, TUPAN_SNAP_CRACKLE = 3, TUPAN_TSTEP = 4,
    TUPAN_PNACC = 5, TUPAN_NREG_X = 6, TUPAN_NREG_V = 7, TUPAN_SAKURA = 8, TUPAN_KEPLER = 9
};

/* All functions below return 0 on success, non-zero on failure (see tupan_cuda_last_error).
 * `iarr`/`jarr`/`out` are arrays of DEVICE pointers in the libtupan.h order of the kernel
 * (e.g. acc_jerk: m rx ry rz e2 vx vy vz | ax ay az jx jy jz); `scal` holds the kernel's
 * scalar arguments as doubles in libtupan.h order (tstep: eta; pnacc: order, inv1..inv7;
 * nreg_X/V: dt; sakura: dt, flag; kepler: dt); `stream` is a cudaStream_t (NULL = legacy
 * default stream).  Calls are asynchronous with respect to the host. */

int tupan_cuda_info(int kernel, int *n_in, int *n_out, int *n_scal, int *flops_per_pair);

/* out[i] = kernel(i-system, j-system): pack + sweep (+ finalize) on `stream` */
int tupan_cuda_run_dev(int kernel, long long ni, const void *const *iarr, long long nj,
                       const void *const *jarr, const double *scal, void *const *out, void *stream);

/* kepler for `pairs` independent binaries: arrays hold 2*pairs bodies, binary b = (2b, 2b+1) */
int tupan_cuda_kepler_dev(long long pairs, const void *const *arr, double dt, void *const *out, void *stream);

/* Sub-step doubling of the Kepler propagator (universal_kepler_solver.h:481-606: the reference
 * doubles without bound, 2^14 .. 2^27 sequential sub-steps for softened tight binaries).  Inside a
 * sakura sweep a pair gets 2^12 sub-steps; a pair that needs more is handed to a clean-up launch
 * (one thread per pair, up to 2^27 sub-steps, like the two-body entry point) whose result is added
 * to the owner's outputs before the call's outputs count as written -- same answer as the
 * reference, the other pairs do not wait.  Only a pair that exceeds 2^27 is counted here: number of
 * such pairs since the last query (resets; synchronises the device); < 0 on error.  The
 * synchronous Part-1 entry points check it themselves and fail loudly. */
long long tupan_cuda_kepler_limit_hits(void);
/* pairs handed to the clean-up launch since the library was loaded (synchronises the device) */
long long tupan_cuda_kepler_cleanup_pairs(void);

/* Building blocks.  A packed j buffer has tupan_cuda_row_width(kernel) REALs per particle
 * (row-major, 16-byte aligned rows) and can be all-gathered across GPUs as one tensor. */
int tupan_cuda_row_width(int kernel, const double *scal);
int tupan_cuda_n_acc(int kernel, const double *scal);
int tupan_cuda_pack_dev(int kernel, long long nj, const void *const *jarr, const double *scal,
                        void *packed, void *stream);
/* number of workspace slots a sweep of `rows` rows for `ni` particles will write */
int tupan_cuda_sweep_slots(int kernel, long long ni, long long rows, const double *scal);
/* sweep packed rows [j0, j1) into raw accumulators partial[slot0 ...][n_acc][ni] */
int tupan_cuda_sweep_dev(int kernel, long long ni, const void *const *iarr, const void *packed,
                         long long j0, long long j1, const double *scal, void *partial, int slot0,
                         void *stream);
/* combine `nslots` accumulator sets, apply the kernel's epilogue, write outputs */
int tupan_cuda_finalize_dev(int kernel, long long ni, const void *const *iarr, const void *partial,
                            int nslots, const double *scal, void *const *out, void *stream);

/* ---- Part 2b: packed rows in peer-visible memory (one process per GPU, NVLink/NVSwitch) ----
 * The multi-GPU path needs every rank's packed rows on every GPU.  With these calls the rows are
 * not copied at all: each rank packs into a buffer its peers have mapped (CUDA IPC), the ranks
 * meet at a device-side barrier on their streams, and tupan_cuda_sweep_dev is pointed at the
 * peer's buffer -- the pair kernel's TMA bulk copies then pull the tiles over NVLink while the
 * previous tiles are being computed.  New work; the reference has no multi-device path. */
/* allocate `bytes` of zeroed device memory other processes can map; handle64: 64 bytes out */
int tupan_cuda_peer_alloc(long long bytes, void **dptr, void *handle64);
/* map a buffer another process of this node allocated with tupan_cuda_peer_alloc */
int tupan_cuda_peer_open(const void *handle64, void **dptr);
int tupan_cuda_peer_close(void *dptr);
int tupan_cuda_peer_free(void *dptr);
/* barrier between the `world` ranks on `stream` (asynchronous for the host).  flags[r] is rank
 * r's flag block (>= (world + 1) * 4 bytes from tupan_cuda_peer_alloc, mapped by everybody);
 * the epoch counter lives in that block, so the call can be captured in a CUDA graph.  A peer
 * that does not arrive within timeout_s is reported by tupan_cuda_peer_timeouts(). */
int tupan_cuda_peer_barrier_dev(void *const *flags, int rank, int world, double timeout_s, void *stream);
/* barriers that gave up since the last call (synchronises the device) */
long long tupan_cuda_peer_timeouts(void);
/* ONE sweep over the packed rows of `nseg` (<= 8) owners, each in its own (peer-mapped) buffer:
 * seg_ptr[k] holds seg_rows[k] rows.  The kernel walks logical 128-row tiles -- no tile straddles
 * two owners -- so transfer over NVLink and arithmetic overlap tile by tile inside one launch.
 * Writes tupan_cuda_sweep_multi_slots(...) accumulator sets starting at slot0. */
int tupan_cuda_sweep_multi_slots(int kernel, long long ni, int nseg, const long long *seg_rows,
                                 const double *scal);
int tupan_cuda_sweep_multi_dev(int kernel, long long ni, const void *const *iarr, int nseg,
                               const void *const *seg_ptr, const long long *seg_rows, const double *scal,
                               void *partial, int slot0, void *stream);

/* min over i of |tstep[i]| on the device (the host-side reduction of
 * tupan/particles/body.py:364-368 fused behind tstep); result is written to *d_min. */
int tupan_cuda_abs_min_dev(long long n, const void *d_values, void *d_min, void *stream);

/* status, tuning, measurement */
int tupan_cuda_init(void);                         /* create the context on the current device */
int tupan_cuda_last_error(char *msg, int msg_len); /* code of the last failure (0 = none) */
void tupan_cuda_clear_error(void);
/* force a launch shape (tests/tuning): lane_split < 0 restores the heuristic.
 * lane_split: 0 = several particles per thread, lanes independent (for the fp64 kernels with a grouped
 * form: its first group shape); 1 = 2^js_log2 lanes of a warp share a particle; 2 = as 0 with the
 * kernel's second, smaller group shape (acc_jerk, acc, tstep, nreg_X in fp64; other kernels run 0).
 * jg = number of j chunks (> 1: raw accumulators go through a workspace and a finalize launch). */
void tupan_cuda_force_plan(int lane_split, int js_log2, int jg);
void tupan_cuda_last_plan(int *lane_split, int *js_log2, int *jg);
/* the launch shape the cost model picks for `kernel` on ni x nj pairs (no device needed) */
int tupan_cuda_plan_query(int kernel, long long ni, long long nj, const double *scal, int *lane_split,
                          int *js_log2, int *jg);
void tupan_cuda_set_timing(int enable);
/* stage times of the last call, milliseconds (CUDA events on the stream the call ran on;
 * a device-resident call has no h2d/d2h stage and reports 0 for them) */
void tupan_cuda_last_times(float *h2d, float *pack, float *pair, float *finalize, float *d2h);
/* stage times summed over every call since tupan_cuda_set_timing(1) (each call keeps its own CUDA
 * events, up to 256 calls; bare tupan_cuda_sweep_dev / _multi_dev calls count as calls with a pair
 * stage only); sums5 = {h2d, pack, pair, finalize, d2h} in ms.  Returns 1 if more than 256 calls
 * were made (the oldest are then missing from the sums), else 0.  Synchronises the device. */
int tupan_cuda_sum_times(float *sums5, long long *calls);
long long tupan_cuda_launch_count(void);           /* kernels launched since load */
/* kernels of this library replayed from a CUDA graph the caller captured (launches inside a
 * capture are counted once, at capture time; the caller adds them per replay) */
void tupan_cuda_count_launches(long long n);
int tupan_cuda_sm_count(void);
/* FMA-pipe micro-benchmark in the library's precision: sustained TFLOP/s over `ms` ms */
int tupan_cuda_fma_peak(double ms, double *tflops, double *sm_mhz_effective);
/* The other instruction shapes of the same pipe: kind 1 = x + y, 2 = x * y, 3 = fma(x, y, 1)
 * (two register operands + immediate); result in 1e12 instructions x lanes per second (an FMA
 * counts 2 flop).  On a B200 these run at the nominal 64 lanes/clk/SM in fp64, while the 3-register
 * chains of tupan_cuda_fma_peak need a third clock per instruction to collect their operands. */
int tupan_cuda_pipe_probe(int kind, double ms, double *tera_ops);
int tupan_cuda_real_bytes(void);                   /* sizeof(REAL): 8 or 4 */

/* ------------------------------------------------------------------------------------ */
/* Part 3 -- O(N) integrator updates on device-resident state      */
/* ------------------------------------------------------------------------------------ */

/* The reference advances the state on the host with numpy between two kernel calls and
 * reads min|tstep| back every step.  These entry points do the same arithmetic (same
 * operation order, one rounding per operation, no FMA contraction) on device arrays, with
 * the step size held in a device control block of TUPAN_CTL_SIZE doubles, so that a whole
 * step is enqueued on one stream with no host round trip.  All are asynchronous on `stream`. */
enum tupan_ctl {
    TUPAN_CTL_T_CURR = 0,   /* current time (the reference's class attribute t_curr)          */
    TUPAN_CTL_TAU = 1,      /* step chosen by tupan_cuda_step_begin_dev (0 once done)          */
    TUPAN_CTL_T_END = 2,    /* set by the caller                                               */
    TUPAN_CTL_ETA = 3,      /* set by the caller                                               */
    TUPAN_CTL_NSTEPS = 4,   /* steps completed                                                 */
    TUPAN_CTL_DONE = 5,     /* 1 when |t_curr| >= |t_end|: further steps are no-ops            */
    TUPAN_CTL_TAU_BASE = 6, /* Base.get_base_tstep of the step                                 */
    TUPAN_CTL_MIN_TS = 7,   /* the minimum time-step the block step was derived from           */
    TUPAN_CTL_SIZE = 8
};

/* replaces Base.get_base_tstep + get_min_block_tstep (integrator/__init__.py:48-78):
 * tau = copysign(min(|t_end|-|t_curr|, |eta|), eta); with d_min_tstep != NULL (a device
 * double holding min_i |tstep_i|) the power-of-two block step commensurate with t_curr. */
int tupan_cuda_step_begin_dev(void *d_ctl, const void *d_min_tstep, void *stream);

/* replaces H2/H4/H6/H8.epredict (integrator/hermite.py:25-43,75-91,127-158,202-242) after the
 * force evaluation: rv = {rx ry rz vx vy vz} is advanced in place, its old value is saved in
 * rv0; d0 = {ax ay az [jx jy jz [sx sy sz [cx cy cz]]]} (order/2 triples) of the old state. */
int tupan_cuda_hermite_predict_dev(int order, long long n, void *const *rv, void *const *rv0,
                                   const void *const *d0, const void *d_ctl, void *stream);
/* replaces H*.ecorrect (hermite.py:45-57,93-121,160-196,244-283) after the force evaluation
 * of the predicted state (derivatives d1) */
int tupan_cuda_hermite_correct_dev(int order, long long n, void *const *rv, void *const *rv0,
                                   const void *const *d0, const void *const *d1, const void *d_ctl,
                                   void *stream);
/* Individual block time-steps (the reference has none: its adaptive Hermite
 * moves every particle with the shared minimum block step, integrator/hermite.py:343-401).
 * order 4 or 6.  block_predict: Taylor prediction of ALL n particles from their own time[i] to
 * t_next; state = {r v a j [s]} as 3 arrays each (12 or 15), pred = {r v [a j]} (6 or 12: order 6
 * also needs a, j of the j-particles for snap_crackle).  block_correct: the Hermite corrector of
 * hermite.py:93-121,160-196 for the n ACTIVE particles (compact arrays), each with its own step
 * tau[i]; rv0 / d0 = state and derivatives at the start of the particle's step, d1 = derivatives
 * at the block time, rv = corrected {r v} out. */
int tupan_cuda_block_predict_dev(int order, long long n, const void *const *state, const void *time,
                                 double t_next, void *const *pred, void *stream);
int tupan_cuda_block_correct_dev(int order, long long n, const void *tau, const void *const *rv0,
                                 const void *const *d0, const void *const *d1, void *const *rv,
                                 void *stream);
/* block_select: the next block time t = min_i (time[i] + dt[i]) and the ascending list of the
 * particles that reach it, one launch: d_out2[0] = t, d_out2[1] = their number (two doubles),
 * d_idx = their indices (n int64 slots). */
int tupan_cuda_block_select_dev(long long n, const void *time, const void *dt, void *d_out2, void *d_idx,
                                void *stream);
/* block_quantize: new step of the n particles that arrived at t_next -- the largest power of two
 * <= ts[i] (tupan's pairwise criterion, tstep_kernel), <= dt_max, <= 2 tau[i] and commensurate
 * with t_next -- and their new time stamp. */
int tupan_cuda_block_quantize_dev(long long n, const void *ts, const void *tau, double t_next,
                                  double dt_max, void *dt_new, void *time_new, void *stream);
/* y[k] += x[k] * REAL(c_inner * (c_outer * tau)), k < narr <= 6; tau = ctl[TAU] (1 if d_ctl
 * is NULL).  Replaces drift_n / kick_n (integrator/sia.py:64-84), the half drifts and the
 * += (dr, dv) of sakura_step (integrator/sakura.py:25-48). */
int tupan_cuda_axpy_dev(int narr, long long n, void *const *y, const void *const *x, double c_outer,
                        double c_inner, const void *d_ctl, void *stream);
/* y[k] = x[k] / REAL(denom): r = mr / mtot, v = mv / mtot of nreg_x / nreg_v (nreg.py:28-30,51-53) */
int tupan_cuda_scale_dev(int narr, long long n, void *const *y, const void *const *x, double denom,
                         void *stream);
/* The post-Newtonian kick of the SIA integrators (kick_pn, integrator/sia.py:136-157) on either
 * side of the pnacc evaluation, with the bookkeeping of PNbodyMethods (particles/body.py:471-527).
 * arr = 23 device pointers: v[3] a[3] w[3] pna[3] mass r[3] pn_ke pn_mv[3] pn_am[3];
 * step = REAL(c_inner * (c_outer * tau)).
 *   phase 0:  v += (a step + w)/2
 *   phase 1:  pn_ke -= (v . m pna) step; pn_mv -= m pna step; pn_am -= (r x m pna) step;
 *             w = 2 pna step - w;  v += (a step + w)/2 */
int tupan_cuda_pn_kick_dev(int phase, long long n, void *const *arr, double c_outer, double c_inner,
                           const void *d_ctl, void *stream);
/* t_curr += tau; tstep[:] = tau; time += tau; nstep += 1 (hermite.py:398-401; sia.py:1105-1113;
 * sakura.py:136-139).  Array pointers may be NULL. */
int tupan_cuda_step_end_dev(long long n, void *d_time, void *d_nstep, void *d_tstep, void *d_ctl,
                            void *stream);

/* the same bookkeeping with the step given by the host (sub-systems of the hierarchical SIA
 * recursion, sia.py:1108-1113): tstep[:] = tau; time += tau; nstep += 1 */
int tupan_cuda_stamp_dev(long long n, void *d_time, void *d_nstep, void *d_tstep, double tau, void *stream);

/* deterministic reductions over particles; the result is ONE double written to d_out */
enum tupan_reduction {
    TUPAN_RED_SUM = 0,       /* sum x0                                (u, mk: nreg.py:31,54)     */
    TUPAN_RED_KINETIC = 1,   /* sum 0.5 m (vx^2+vy^2+vz^2)            (body.py:262-275)          */
    TUPAN_RED_HALF_DOT = 2,  /* 0.5 sum x0 x1  (m, phi)               (body.py:294-306)          */
    TUPAN_RED_SAKURA_DT = 3, /* eta / sqrt(1 + max((eta/x0)^2-(eta/x1)^2)), param = eta
                                (tstep, tstepij)                      (sakura.py:100-110)        */
    TUPAN_RED_ABS_MIN = 4,   /* min |x0|                              (body.py:364-368)          */
    TUPAN_RED_ABS_MAX = 5,   /* max |x0|                              (body.py:370-374)          */
    TUPAN_RED_DOT = 6,       /* sum x0 x1  (m r, m v: centre of mass, linear momentum; body.py:88-126) */
    TUPAN_RED_MOMENT = 7     /* sum x0 (x1 x4 - x2 x3)  (m, ra, rb, va, vb: one component of the
                                angular momentum, body.py:175-186)                               */
};
int tupan_cuda_reduce_dev(int what, long long n, const void *const *arrays, double param, void *d_out,
                          void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LIBTUPAN_CUDA_H */
