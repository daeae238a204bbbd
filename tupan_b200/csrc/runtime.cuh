// runtime.cuh -- host-side runtime shared by every kernel family: the per-library context
// (stream, cached device buffers, timing events, error state) and the generic runners that
// turn an Op into the entry points of a KernelVTable.
#pragma once
#include <mutex>
#include <stdio.h>
#include <string.h>
#include <vector>

#include "pair_engine.cuh"

namespace tupan {

enum KernelId {
    K_PHI = 0, K_ACC, K_ACC_JERK, K_SNAP_CRACKLE, K_TSTEP, K_PNACC, K_NREG_X, K_NREG_V, K_SAKURA,
    K_KEPLER, K_COUNT
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    std::vector<void*> retired;   // outgrown blocks, kept alive (see ensure)
    void* ensure(size_t bytes);
    void release();
};

// Page-locked host staging block (small host-pointer calls, see Runner::run_host).
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    void* ensure(size_t bytes);
    void release();
};

struct StageTimes { float h2d_ms, pack_ms, pair_ms, finalize_ms, d2h_ms; };

struct Context {
    std::mutex mu;
    bool ready = false;
    int device = -1;
    DeviceInfo info = {148};
    cudaStream_t stream = nullptr;
    DevBuf in_i[MAX_IN], in_j[MAX_IN], outb[MAX_OUT], jpack, partial;
    // small calls: every array of the call in ONE device block, mirrored by ONE pinned host block
    DevBuf stage_dev;
    HostBuf stage_host;
    // forced launch plan (tests, tuning); lane_split < 0 means "choose"
    Plan forced = {-1, 0, 1};
    // timing (only when enabled): every call records its stage boundaries into its own set of
    // events, a ring of TIME_RING sets, so that the stage times of all calls since
    // tupan_cuda_set_timing(1) can be summed afterwards without a host sync in between
    enum { TIME_RING = 256 };
    bool timing = false;
    cudaEvent_t (*evr)[6] = nullptr;
    unsigned* evmark = nullptr;   // which of a set's events were recorded
    long long nsets = 0;          // sets started since timing was enabled
    StageTimes last = {0, 0, 0, 0, 0};
    Plan last_plan = {0, 0, 1};
    long long launches = 0;   // kernels launched by this library since load
    int last_error = 0;       // 0 = ok, else cudaError_t (or -1 for argument errors)
    char last_msg[256] = {0};

    // The packed-row and accumulator scratch (jpack, partial) is shared by every call of the process,
    // whatever stream the caller passes: a call on another stream than the previous one first waits
    // for the event the previous call recorded behind its last use of the scratch.  (Skipped while a
    // stream is being captured into a CUDA graph: the graph's own launches are ordered by the capture.)
    cudaEvent_t scratch_done = nullptr;
    cudaStream_t scratch_stream = nullptr;
    bool scratch_used = false;
    static bool capturing(cudaStream_t s)
    {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        return cudaStreamIsCapturing(s, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone;
    }
    void scratch_acquire(cudaStream_t s)
    {
        if (capturing(s)) return;
        if (scratch_used && scratch_stream != s && scratch_done) cudaStreamWaitEvent(s, scratch_done, 0);
    }
    void scratch_release(cudaStream_t s)
    {
        if (capturing(s)) return;
        if (!scratch_done && cudaEventCreateWithFlags(&scratch_done, cudaEventDisableTiming) != cudaSuccess) {
            scratch_done = nullptr;
            return;
        }
        cudaEventRecord(scratch_done, s);
        scratch_stream = s;
        scratch_used = true;
    }

    int init();
    int fail(cudaError_t e, const char* where);
    int cur_set() const { return (int)((nsets - 1) % TIME_RING); }
    void begin_set()
    {
        nsets++;
        evmark[cur_set()] = 0;
    }
    void mark(int k, cudaStream_t s)
    {
        if (!timing || !evr) return;
        // a new call starts at event 0 (host-pointer entry), at event 1 (device-resident entry) or
        // at event 2 (a bare sweep of the multi-GPU path)
        if (nsets == 0 || k == 0 || (k == 1 && evmark[cur_set()] != 1u) || (k == 2 && !(evmark[cur_set()] & 2u)))
            begin_set();
        cudaEventRecord(evr[cur_set()][k], s);
        evmark[cur_set()] |= 1u << k;
    }
    // stage times of set q (stage k lies between events k and k+1); false if nothing was recorded
    bool set_times(int q, float (&t)[5]) const
    {
        bool any = false;
        for (int k = 0; k < 5; ++k) {
            t[k] = 0;
            if ((evmark[q] >> k & 1u) && (evmark[q] >> (k + 1) & 1u) && cudaEventSynchronize(evr[q][k + 1]) == cudaSuccess) {
                cudaEventElapsedTime(&t[k], evr[q][k], evr[q][k + 1]);
                any = true;
            }
        }
        return any;
    }
};

Context& ctx();

#define TUPAN_CHECK(call, where)                               \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return ctx().fail(e__, where); \
    } while (0)

// What every kernel family exports to the ABI layer.  `scal` carries the kernel's scalar
// arguments as doubles in libtupan.h order (eta | order,inv1..inv7 | dt | dt,flag).
struct KernelVTable {
    const char* name;
    int n_in;        // caller arrays per side
    int n_out;       // output arrays
    int n_scal;      // scalars in `scal`
    int flops;       // reference flops/pair convention (0 = data dependent)
    int (*row_width)(const double* scal);  // reals per packed row
    int (*n_acc)(const double* scal);      // raw accumulators per particle
    // host pointers in, host pointers out, synchronous (the libtupan.h contract)
    int (*run_host)(long long ni, const real_t* const* hi, long long nj, const real_t* const* hj,
                    const double* scal, real_t* const* hout);
    // device pointers, asynchronous on `stream`
    int (*run_dev)(long long ni, const real_t* const* di, long long nj, const real_t* const* dj,
                   const double* scal, real_t* const* dout, cudaStream_t stream);
    // building blocks (device-resident / multi-GPU path)
    int (*pack)(long long nj, const real_t* const* dj, const double* scal, real_t* packed, cudaStream_t stream);
    int (*sweep_slots)(long long ni, long long rows, const double* scal);
    int (*sweep)(long long ni, const real_t* const* di, const real_t* packed, long long j0, long long j1,
                 const double* scal, real_t* partial, int slot0, cudaStream_t stream);
    int (*finalize)(long long ni, const real_t* const* di, const real_t* partial, int nslots,
                    const double* scal, real_t* const* dout, cudaStream_t stream);
    // one sweep over the rows of several owners (peer-mapped buffers); slots as sweep_slots(ni, logical rows)
    int (*sweep_multi)(long long ni, const real_t* const* di, int nseg, const real_t* const* seg_ptr,
                       const long long* seg_rows, const double* scal, real_t* partial, int slot0,
                       cudaStream_t stream);
};

// logical rows of a multi-owner sweep: every owner's rows rounded up to whole 128-row tiles
inline long long multi_logical_rows(int nseg, const long long* seg_rows)
{
    long long tiles = 0;
    for (int k = 0; k < nseg; ++k)
        if (seg_rows[k] > 0) tiles += (seg_rows[k] + 127) / 128;
    return tiles * 128;
}

const KernelVTable* vtable(int kernel);

// ---------------------------------------------------------------------------------------
// Generic runners
// ---------------------------------------------------------------------------------------
template <class Op> struct Runner {
    typedef typename Op::real T;
    enum { NJP = Packed<Op>::NJP };

    static InRefs<T> in_refs(const T* const* a, int n)
    {
        InRefs<T> r;
        for (int k = 0; k < MAX_IN; ++k) r.p[k] = k < n ? a[k] : nullptr;
        return r;
    }
    static OutRefs<T> out_refs(T* const* a, int n)
    {
        OutRefs<T> r;
        for (int k = 0; k < MAX_OUT; ++k) r.p[k] = k < n ? a[k] : nullptr;
        return r;
    }
    static Plan plan_for(long long ni, long long rows)
    {
        Context& c = ctx();
        Plan p = c.forced.lane_split >= 0 ? c.forced : choose_plan<Op>(c.info, ni, rows);
        if (p.jg < 1) p.jg = 1;
        if (p.js_log2 < 0) p.js_log2 = 0;
        if (p.js_log2 > 5) p.js_log2 = 5;
        if (p.lane_split != 1) p.js_log2 = 0;
        return p;
    }

    static int row_width(const double*) { return NJP; }
    static int n_acc(const double*) { return Op::NA; }

    static int pack(int n_in, long long nj, const T* const* dj, T* packed, cudaStream_t s)
    {
        TUPAN_CHECK(launch_pack<Op>(in_refs(dj, n_in), nj, packed, s), "pack_j");
        if (nj > 0) ctx().launches++;
        return 0;
    }
    static int sweep_slots(long long ni, long long rows)
    {
        const Plan p = plan_for(ni, rows);
        ctx().last_plan = p;          // readable through tupan_cuda_last_plan (plan queries, tests)
        return p.jg;
    }
    static int sweep(int n_in, long long ni, const T* const* di, const T* packed, long long j0, long long j1,
                     const typename Op::Params& prm, T* partial, int slot0, cudaStream_t s)
    {
        Plan p = plan_for(ni, j1 - j0);
        OutRefs<T> none = out_refs(nullptr, 0);
        ctx().mark(2, s);
        TUPAN_CHECK(launch_pairs<Op>(p, in_refs(di, n_in), ni, packed, j0, j1, prm, partial, slot0, none, s),
                    "pair sweep");
        ctx().mark(3, s);
        if (ni > 0) ctx().launches++;
        ctx().last_plan = p;
        return 0;
    }
    static int sweep_multi(int n_in, long long ni, const T* const* di, int nseg, const T* const* seg_ptr,
                           const long long* seg_rows, const typename Op::Params& prm, T* partial, int slot0,
                           cudaStream_t s)
    {
        typedef Tune<Op> U;
        Context& c = ctx();
        if (nseg < 1 || nseg > MAX_SEG) return c.fail(cudaErrorInvalidValue, "sweep_multi: 1..8 owners");
        Segments<T> seg;
        seg.nseg = 0;
        seg.tile0[0] = 0;
        for (int k = 0; k < nseg; ++k) {
            if (seg_rows[k] <= 0) continue;                       // an owner without particles
            if (seg_rows[k] > 0x7fffffffLL) return c.fail(cudaErrorInvalidValue, "sweep_multi: rows per owner");
            seg.rows[seg.nseg] = (int)seg_rows[k];
            seg.ptr[seg.nseg] = seg_ptr[k];
            seg.tile0[seg.nseg + 1] = seg.tile0[seg.nseg] + (int)((seg_rows[k] + U::TJ - 1) / U::TJ);
            seg.nseg++;
        }
        if (seg.nseg == 0 || ni <= 0) return 0;
        const long long rows = (long long)seg.tile0[seg.nseg] * U::TJ;
        Plan p = plan_for(ni, rows);
        OutRefs<T> none = out_refs(nullptr, 0);
        c.mark(2, s);
        TUPAN_CHECK(launch_pairs<Op>(p, in_refs(di, n_in), ni, nullptr, 0, rows, prm, partial, slot0, none, s, &seg),
                    "pair sweep (multi-owner)");
        c.mark(3, s);
        c.launches++;
        c.last_plan = p;
        return 0;
    }
    static int finalize(int n_in, int n_out, long long ni, const T* const* di, const T* partial, int nslots,
                        const typename Op::Params& prm, T* const* dout, cudaStream_t s)
    {
        TUPAN_CHECK(launch_finalize<Op>(in_refs(di, n_in), ni, partial, nslots, out_refs(dout, n_out), prm, s),
                    "finalize");
        if (ni > 0) ctx().launches++;
        return after_outputs(n_out, ni, prm, dout, s);
    }
    // Ops with a deferred slow case (sakura) may have handed pairs to a clean-up launch
    static int after_outputs(int n_out, long long ni, const typename Op::Params& prm, T* const* dout, cudaStream_t s)
    {
        if constexpr (Defers<Op>::value != 0) {
            TUPAN_CHECK(Op::after_sweeps(ni, prm, out_refs(dout, n_out), s, &ctx().launches), "clean-up of deferred pairs");
        }
        return 0;
    }

    // device pointers in/out; uses the context's packed/partial buffers
    static int run_dev(int n_in, int n_out, long long ni, const T* const* di, long long nj, const T* const* dj,
                       const typename Op::Params& prm, T* const* dout, cudaStream_t s)
    {
        Context& c = ctx();
        if (ni <= 0) return 0;
        T* packed = static_cast<T*>(c.jpack.ensure((size_t)(nj > 0 ? nj : 1) * NJP * sizeof(T)));
        if (!packed) return c.fail(cudaErrorMemoryAllocation, "packed j buffer");
        c.scratch_acquire(s);
        c.mark(1, s);
        int rc = pack(n_in, nj, dj, packed, s);
        if (rc) return rc;
        c.mark(2, s);
        Plan p = plan_for(ni, nj);
        c.last_plan = p;
        if (p.jg == 1) {
            TUPAN_CHECK(launch_pairs<Op>(p, in_refs(di, n_in), ni, packed, 0, nj, prm, nullptr, 0,
                                         out_refs(dout, n_out), s),
                        "pair kernel");
            c.launches++;
            c.mark(3, s);
            rc = after_outputs(n_out, ni, prm, dout, s);
            if (rc) return rc;
        } else {
            T* part = static_cast<T*>(c.partial.ensure((size_t)p.jg * Op::NA * ni * sizeof(T)));
            if (!part) return c.fail(cudaErrorMemoryAllocation, "partial workspace");
            OutRefs<T> none = out_refs(nullptr, 0);
            TUPAN_CHECK(launch_pairs<Op>(p, in_refs(di, n_in), ni, packed, 0, nj, prm, part, 0, none, s),
                        "pair kernel");
            c.launches++;
            c.mark(3, s);
            rc = finalize(n_in, n_out, ni, di, part, p.jg, prm, dout, s);
            if (rc) return rc;
        }
        c.mark(4, s);
        c.scratch_release(s);
        return 0;
    }

    // host pointers in/out (libtupan.h contract): H2D, run, D2H, synchronous return.
    // A j array that is the same host array as an i array (ips is jps, or a prefix slice of
    // it) is not copied twice.
    //
    // Small calls (the reference's integrators at N ~ 1e3 make thousands of them) are dominated
    // by the per-copy cost of 14-28 separate transfers to and from pageable numpy memory
    // (measured on B200, N = 1024 acc_jerk: 20 us H2D + 73 us D2H around 30 us of kernels).
    // Up to STAGE_MAX bytes the arrays are therefore gathered into one page-locked block on the
    // host, moved with ONE copy each way, and live in one device block.
    enum { STAGE_MAX = 1 << 20 };
    static long long stage_stride(long long n) { return (n + 3) / 4 * 4; }   // keeps 16-byte alignment

    static int run_host(int n_in, int n_out, long long ni, const T* const* hi, long long nj, const T* const* hj,
                        const typename Op::Params& prm, T* const* hout)
    {
        Context& c = ctx();
        std::lock_guard<std::mutex> lock(c.mu);
        int rc = c.init();
        if (rc) return rc;
        if (ni <= 0) return 0;
        cudaStream_t s = c.stream;
        const T* di[MAX_IN];
        const T* dj[MAX_IN];
        T* dout[MAX_OUT];
        // which j arrays are the caller's i arrays again
        int alias[MAX_IN];
        int n_jown = 0;
        for (int k = 0; k < n_in; ++k) {
            alias[k] = -1;
            if (nj <= ni) {
                for (int q = 0; q < n_in; ++q)
                    if (hj[k] == hi[q]) { alias[k] = q; break; }
            }
            if (alias[k] < 0 && nj > 0) n_jown++;
        }
        const long long si = stage_stride(ni), sj = stage_stride(nj > 0 ? nj : 0);
        const size_t in_elems = (size_t)n_in * si + (size_t)n_jown * sj;
        const size_t all_elems = in_elems + (size_t)n_out * si;
        c.mark(0, s);
        if (all_elems * sizeof(T) <= (size_t)STAGE_MAX) {
            T* hs = static_cast<T*>(c.stage_host.ensure(all_elems * sizeof(T)));
            T* ds = static_cast<T*>(c.stage_dev.ensure(all_elems * sizeof(T)));
            if (!hs || !ds) return c.fail(cudaErrorMemoryAllocation, "staging block");
            size_t off = 0;
            for (int k = 0; k < n_in; ++k) {
                memcpy(hs + off, hi[k], (size_t)ni * sizeof(T));
                di[k] = ds + off;
                off += si;
            }
            for (int k = 0; k < n_in; ++k) {
                if (alias[k] >= 0) { dj[k] = di[alias[k]]; continue; }
                dj[k] = nullptr;
                if (nj <= 0) continue;
                memcpy(hs + off, hj[k], (size_t)nj * sizeof(T));
                dj[k] = ds + off;
                off += sj;
            }
            for (int k = 0; k < n_out; ++k) dout[k] = ds + in_elems + (size_t)k * si;
            TUPAN_CHECK(cudaMemcpyAsync(ds, hs, in_elems * sizeof(T), cudaMemcpyHostToDevice, s), "H2D block");
            rc = run_dev(n_in, n_out, ni, di, nj, dj, prm, dout, s);
            if (rc) return rc;
            TUPAN_CHECK(cudaMemcpyAsync(hs + in_elems, ds + in_elems, (size_t)n_out * si * sizeof(T),
                                        cudaMemcpyDeviceToHost, s), "D2H block");
            c.mark(5, s);
            TUPAN_CHECK(cudaStreamSynchronize(s), "synchronize");
            for (int k = 0; k < n_out; ++k)
                memcpy(hout[k], hs + in_elems + (size_t)k * si, (size_t)ni * sizeof(T));
            return 0;
        }
        for (int k = 0; k < n_in; ++k) {
            T* d = static_cast<T*>(c.in_i[k].ensure((size_t)ni * sizeof(T)));
            if (!d) return c.fail(cudaErrorMemoryAllocation, "i buffer");
            TUPAN_CHECK(cudaMemcpyAsync(d, hi[k], (size_t)ni * sizeof(T), cudaMemcpyHostToDevice, s), "H2D i");
            di[k] = d;
        }
        for (int k = 0; k < n_in; ++k) {
            dj[k] = alias[k] >= 0 ? di[alias[k]] : nullptr;
            if (!dj[k] && nj > 0) {
                T* d = static_cast<T*>(c.in_j[k].ensure((size_t)nj * sizeof(T)));
                if (!d) return c.fail(cudaErrorMemoryAllocation, "j buffer");
                TUPAN_CHECK(cudaMemcpyAsync(d, hj[k], (size_t)nj * sizeof(T), cudaMemcpyHostToDevice, s), "H2D j");
                dj[k] = d;
            }
        }
        for (int k = 0; k < n_out; ++k) {
            dout[k] = static_cast<T*>(c.outb[k].ensure((size_t)ni * sizeof(T)));
            if (!dout[k]) return c.fail(cudaErrorMemoryAllocation, "out buffer");
        }
        rc = run_dev(n_in, n_out, ni, di, nj, dj, prm, dout, s);
        if (rc) return rc;
        for (int k = 0; k < n_out; ++k)
            TUPAN_CHECK(cudaMemcpyAsync(hout[k], dout[k], (size_t)ni * sizeof(T), cudaMemcpyDeviceToHost, s), "D2H");
        c.mark(5, s);
        TUPAN_CHECK(cudaStreamSynchronize(s), "synchronize");
        return 0;
    }
};

// Glue: build the vtable entry points of an Op whose Params come from `scal` via MakePrm.
#define TUPAN_DEFINE_VTABLE(VT, OP, NAME, N_IN, N_OUT, N_SCAL, FLOPS, MAKEPRM)                                   \
    namespace {                                                                                                   \
    typedef Runner<OP> R_##VT;                                                                                    \
    int VT##_rw(const double* s) { return R_##VT::row_width(s); }                                                 \
    int VT##_na(const double* s) { return R_##VT::n_acc(s); }                                                     \
    int VT##_host(long long ni, const real_t* const* hi, long long nj, const real_t* const* hj,                  \
                  const double* s, real_t* const* ho)                                                             \
    { return R_##VT::run_host(N_IN, N_OUT, ni, hi, nj, hj, MAKEPRM(s), ho); }                                     \
    int VT##_dev(long long ni, const real_t* const* di, long long nj, const real_t* const* dj, const double* s,   \
                 real_t* const* dout, cudaStream_t st)                                                            \
    {                                                                                                             \
        Context& c = ctx();                                                                                       \
        std::lock_guard<std::mutex> lock(c.mu);                                                                   \
        int rc = c.init();                                                                                        \
        if (rc) return rc;                                                                                        \
        return R_##VT::run_dev(N_IN, N_OUT, ni, di, nj, dj, MAKEPRM(s), dout, st);                                \
    }                                                                                                             \
    int VT##_pack(long long nj, const real_t* const* dj, const double*, real_t* packed, cudaStream_t st)         \
    { return R_##VT::pack(N_IN, nj, dj, packed, st); }                                                            \
    int VT##_slots(long long ni, long long rows, const double*) { return R_##VT::sweep_slots(ni, rows); }        \
    int VT##_sweep(long long ni, const real_t* const* di, const real_t* packed, long long j0, long long j1,      \
                   const double* s, real_t* partial, int slot0, cudaStream_t st)                                  \
    { return R_##VT::sweep(N_IN, ni, di, packed, j0, j1, MAKEPRM(s), partial, slot0, st); }                       \
    int VT##_fin(long long ni, const real_t* const* di, const real_t* partial, int nslots, const double* s,      \
                 real_t* const* dout, cudaStream_t st)                                                            \
    { return R_##VT::finalize(N_IN, N_OUT, ni, di, partial, nslots, MAKEPRM(s), dout, st); }                      \
    int VT##_multi(long long ni, const real_t* const* di, int nseg, const real_t* const* sp,                     \
                   const long long* sr, const double* s, real_t* partial, int slot0, cudaStream_t st)            \
    { return R_##VT::sweep_multi(N_IN, ni, di, nseg, sp, sr, MAKEPRM(s), partial, slot0, st); }                   \
    }                                                                                                             \
    extern const KernelVTable VT = {NAME,      N_IN,      N_OUT,     N_SCAL,     FLOPS,      VT##_rw,  VT##_na,   \
                                    VT##_host, VT##_dev,  VT##_pack, VT##_slots, VT##_sweep, VT##_fin, VT##_multi};

}  // namespace tupan
