// pair_engine.cuh -- the one O(ni*nj) engine every tupan pairwise kernel runs on.
//
// Shape of the problem (every kernel of tupan/lib/src is of this form, e.g. the double
// loop acc_jerk_kernel.c:31-60):      out[i] = finish( reduce_j  pair(i, j) )
//
// B200 mapping
//   * j-state: packed once per call by pack_j_kernel into rows of NJP reals (AoS, row =
//     16-byte multiples) so that a tile of TJ rows is ONE contiguous block.  Tiles are
//     streamed into shared memory by the TMA engine with 1-D bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) through a STAGES-deep ring of
//     mbarrier-guarded buffers, issued by one elected thread; no thread spends registers
//     or issue slots on staging.
//   * i-state: WPT particles per thread, inputs and accumulators register-resident for the
//     whole j sweep.  All lanes of a warp read the same packed row with 16-byte LDS
//     (a broadcast, conflict-free).
//   * small ni: JS = 2^k lanes of a warp share one i-particle and stride over the rows of a
//     tile (lane split), partial accumulators are combined with warp shuffles; if that still
//     leaves SMs idle, the j range is also split over blockIdx.y, raw accumulators go to a
//     partial[slot][acc][i] workspace and finalize_kernel combines them and applies the
//     kernel's epilogue (deterministic, no atomics).  The same workspace mechanism combines
//     the local-shard and remote-shard sweeps of the multi-GPU path.
//
// An "Op" supplies the physics:
//   typedef real;  enum { NI, NJ, NA, NO, WPT, UNROLL };  struct Params;
//   load_i(const real* const* iarr, long long i, real (&s)[NI])
//   pack_j(const real* const* jarr, long long j, real (&row)[NJP])
//   zero(real (&a)[NA]);  pair(s, row, a, prm);  combine(a, b)
//   finish(iarr, i, a, prm, real* const* out)
//
// Ops whose pairs are cheap most of the time and very expensive now and then (sakura: a
// leapfrog for wide pairs, a universal-variable Kepler solve for close ones) set
// Defers<Op>::value and supply, instead of pair():
//   enum { ND };                                    reals that describe one deferred pair
//   bool pair_fast(s, row, a, prm, real (&d)[ND])   cheap case: accumulate, return false;
//                                                   expensive case: fill d, return true
//   bool pair_slow(d, prm, real (&c)[NA])           full evaluation of one deferred pair; false if
//                                                   even that gave up (bounded sub-stepping): c = 0
//   defer_to_cleanup(d, i, prm)                     ... and the pair is handed to a later launch
// pair_kernel_defer parks the expensive pairs of a warp in a shared-memory ring and runs them
// 32 at a time, one per lane, so the slow path executes with full warps instead of dragging 31
// idle lanes along each time one lane hits it.  Results return to the owner lane in ring order,
// which for any given particle is its own j order: the sums stay deterministic.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace tupan {

enum { MAX_IN = 14, MAX_OUT = 7 };

template <typename T> struct InRefs  { const T* p[MAX_IN]; };
template <typename T> struct OutRefs { T* p[MAX_OUT]; };

template <class Op> struct Packed {
    typedef typename Op::real T;
    enum { NJP = round_up(Op::NJ, Vec16<T>::N), ROW_BYTES = NJP * sizeof(T) };
};

// ---------------------------------------------------------------------------------------
// pack_j_kernel: SoA j arrays -> packed rows [j][NJP].  O(nj), coalesced reads.
// ---------------------------------------------------------------------------------------
template <class Op>
__global__ void pack_j_kernel(InRefs<typename Op::real> j, long long nj, typename Op::real* __restrict__ packed)
{
    typedef typename Op::real T;
    typedef typename Vec16<T>::type V;
    constexpr int NJP = Packed<Op>::NJP;
    constexpr int VN = Vec16<T>::N;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nj;
         r += (long long)gridDim.x * blockDim.x) {
        T row[NJP];
#pragma unroll
        for (int k = 0; k < NJP; ++k) row[k] = T(0);
        Op::pack_j(j.p, r, row);
        V* dst = reinterpret_cast<V*>(packed + r * NJP);
#pragma unroll
        for (int k = 0; k < NJP / VN; ++k) dst[k] = *reinterpret_cast<V*>(&row[k * VN]);
    }
}

// ---------------------------------------------------------------------------------------
// pair_kernel
// ---------------------------------------------------------------------------------------
// Rows that live in several buffers (multi-GPU: every owner's packed rows where they were packed,
// reached through peer mappings).  The sweep then runs over LOGICAL tiles: segment s contributes
// ceil(rows[s] / TJ) tiles, tile0[] are the prefix sums, and a tile never straddles two segments.
enum { MAX_SEG = 8 };
template <typename T> struct Segments {
    int nseg;                     // 0: one buffer (jpack, j0, j1)
    int rows[MAX_SEG];
    int tile0[MAX_SEG + 1];
    const T* ptr[MAX_SEG];
};

template <class Op> struct PairArgs {
    typedef typename Op::real T;
    InRefs<T> i;            // caller's i arrays, libtupan.h order
    long long ni;
    const T* jpack;         // packed j rows
    Segments<T> seg;        // ... or the rows of several owners (then j0 = 0, j1 = logical tiles * TJ)
    long long j0, j1;       // rows [j0, j1) are swept by this launch ...
    long long jchunk;       // ... blockIdx.y takes rows [j0 + y*jchunk, +jchunk)
    int js_log2;            // log2(lanes per i-particle), lane-split variant only
    int slot0;              // first workspace slot of this launch
    T* partial;             // nullptr: apply epilogue and write outputs directly
    OutRefs<T> out;
    typename Op::Params prm;
    int one;                // always 1; a trip count the compiler cannot see (pair_kernel_grouped)
};

template <class Op, int TJ, int STAGES> struct PairSmem {
    enum {
        TILE_BYTES = TJ * Packed<Op>::ROW_BYTES,
        BYTES = STAGES * TILE_BYTES + STAGES * 8
    };
};

template <class Op> struct Defers { enum { value = 0 }; };

template <class Op, int NT, int TJ, int STAGES, bool DEFER = (Defers<Op>::value != 0)> struct PairSmemDefer {
    enum { QCAP = 64, QUEUE_OFF = 0, WARP_BYTES = 0, BYTES = PairSmem<Op, TJ, STAGES>::BYTES };
};
template <class Op, int NT, int TJ, int STAGES> struct PairSmemDefer<Op, NT, TJ, STAGES, true> {
    enum {
        QCAP = 64,   // ring slots per warp: < 32 waiting + at most 32 pushed by one j step
        QUEUE_OFF = (PairSmem<Op, TJ, STAGES>::BYTES + 15) / 16 * 16,
        WARP_BYTES = QCAP * (Op::ND * (int)sizeof(typename Op::real) + (int)sizeof(int)),
        BYTES = QUEUE_OFF + (NT / 32) * WARP_BYTES
    };
};

template <class Op>
TUPAN_DEV void load_row(const typename Op::real* p, typename Op::real (&row)[Packed<Op>::NJP])
{
    typedef typename Op::real T;
    typedef typename Vec16<T>::type V;
    constexpr int VN = Vec16<T>::N;
#pragma unroll
    for (int k = 0; k < Packed<Op>::NJP / VN; ++k) {
        V v = reinterpret_cast<const V*>(p)[k];
        const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int c = 0; c < VN; ++c) row[k * VN + c] = e[c];
    }
}

// Where the tile that starts at (logical) row r0 lives and how many rows it holds.
template <class Op, int TJ, bool MULTI>
TUPAN_DEV const typename Op::real* locate_tile(const PairArgs<Op>& a, long long r0, long long jhi, int& cnt)
{
    constexpr int NJP = Packed<Op>::NJP;
    if (!MULTI) {
        const long long rem = jhi - r0;
        cnt = rem > TJ ? TJ : (int)rem;
        return a.jpack + r0 * NJP;
    }
    const int L = (int)(r0 / TJ);
    int s = 0;
    while (s + 1 < a.seg.nseg && L >= a.seg.tile0[s + 1]) ++s;
    const int off = (L - a.seg.tile0[s]) * TJ;
    const int rem = a.seg.rows[s] - off;
    cnt = rem > TJ ? TJ : rem;
    return a.seg.ptr[s] + (long long)off * NJP;
}

// MULTI = rows in several buffers (Segments); a separate instantiation so that the single-buffer
// kernel keeps exactly the code (and the 122 registers) it was tuned with -- folding the two
// cost the headline 3 %.
template <class Op, int NT, int WPT, int TJ, int STAGES, bool LANE_SPLIT, bool MULTI = false>
__global__ void __launch_bounds__(NT) pair_kernel(const __grid_constant__ PairArgs<Op> a)
{
    typedef typename Op::real T;
    constexpr int NJP = Packed<Op>::NJP;
    constexpr int TILE_ELEMS = TJ * NJP;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tiles = reinterpret_cast<T*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * TILE_ELEMS * sizeof(T));

    const int tid = threadIdx.x;
    const int jsl = LANE_SPLIT ? a.js_log2 : 0;
    const int js = 1 << jsl;
    // Lane split: the lanes of a warp are laid out particle-minor -- lane = jsub * ipw + (particle
    // within the warp), ipw = 32 / js -- so that neighbouring lanes read the SAME packed row (a
    // shared-memory broadcast).  Row-minor (lane = particle * js + jsub) made the 8 lanes of a
    // quarter warp read 8 rows 64 bytes apart: a 4-way bank conflict on every LDS.128, which is
    // what bounded the split kernels (N = 4096, 8 lanes per particle: 150 us).
    const int ipw = 32 >> jsl;                        // particles per warp
    const int jsub = (tid & 31) >> (5 - jsl);
    const int islot = (tid >> 5) * ipw + (tid & (ipw - 1));
    const int slots = NT >> jsl;
    const long long ibase = (long long)blockIdx.x * (slots * WPT);

    T is[WPT][Op::NI];
    T acc[WPT][Op::NA];
#pragma unroll
    for (int w = 0; w < WPT; ++w) {
        long long i = ibase + (long long)w * slots + islot;
        if (i > a.ni - 1) i = a.ni - 1;  // clamp: computes a duplicate, never stored
        Op::load_i(a.i.p, i, is[w]);
        Op::zero(acc[w]);
    }

    const long long jlo = a.j0 + (long long)blockIdx.y * a.jchunk;
    long long jhi = jlo + a.jchunk;
    if (jhi > a.j1) jhi = a.j1;
    const int ntiles = (jhi > jlo) ? (int)((jhi - jlo + TJ - 1) / TJ) : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t) {  // elected thread: start the bulk copy of tile t
        const int s = t % STAGES;
        int cnt;
        const T* src = locate_tile<Op, TJ, MULTI>(a, jlo + (long long)t * TJ, jhi, cnt);
        const unsigned bytes = (unsigned)cnt * Packed<Op>::ROW_BYTES;
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(tiles + s * TILE_ELEMS, src, bytes, &full[s]);
    };
    if (tid == 0) {
        for (int t = 0; t < STAGES && t < ntiles; ++t) issue(t);
    }

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(&full[s], (unsigned)(t / STAGES) & 1u);
        const T* sj = tiles + s * TILE_ELEMS;
        int cnt;
        locate_tile<Op, TJ, MULTI>(a, jlo + (long long)t * TJ, jhi, cnt);

        if (!LANE_SPLIT) {
            if (cnt == TJ) {
#pragma unroll Op::UNROLL
                for (int j = 0; j < TJ; ++j) {
                    T row[NJP];
                    load_row<Op>(sj + j * NJP, row);
#pragma unroll
                    for (int w = 0; w < WPT; ++w) Op::pair(is[w], row, acc[w], a.prm);
                }
            } else {
                for (int j = 0; j < cnt; ++j) {
                    T row[NJP];
                    load_row<Op>(sj + j * NJP, row);
#pragma unroll
                    for (int w = 0; w < WPT; ++w) Op::pair(is[w], row, acc[w], a.prm);
                }
            }
        } else {
            for (int j = jsub; j < cnt; j += js) {
                T row[NJP];
                load_row<Op>(sj + j * NJP, row);
#pragma unroll
                for (int w = 0; w < WPT; ++w) Op::pair(is[w], row, acc[w], a.prm);
            }
        }
        __syncthreads();  // every warp is done with stage s -> it may be refilled
        if (tid == 0 && t + STAGES < ntiles) issue(t + STAGES);
    }

    if (LANE_SPLIT) {
        for (int off = 16; off >= ipw; off >>= 1) {
#pragma unroll
            for (int w = 0; w < WPT; ++w) {
                T other[Op::NA];
#pragma unroll
                for (int k = 0; k < Op::NA; ++k) other[k] = __shfl_xor_sync(0xffffffffu, acc[w][k], off);
                Op::combine(acc[w], other);
            }
        }
    }

    if (jsub == 0) {
#pragma unroll
        for (int w = 0; w < WPT; ++w) {
            const long long i = ibase + (long long)w * slots + islot;
            if (i < a.ni) {
                if (a.partial != nullptr) {
                    T* dst = a.partial + ((long long)(a.slot0 + blockIdx.y) * Op::NA) * a.ni + i;
#pragma unroll
                    for (int k = 0; k < Op::NA; ++k) dst[(long long)k * a.ni] = acc[w][k];
                } else {
                    Op::finish(a.i.p, i, acc[w], a.prm, a.out.p);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// pair_kernel_grouped: the throughput shape for Ops that supply their pair body in GROUPED form.
//
// What it is for (measured on B200: tools/microbench2.cu, tools/sass_rf.py, tools/kernel_lab2.cu,
// profiles/r02_microbench_fp64_operands.txt, profiles/r02_kernel_lab_grouped.txt):
// an FP64 instruction holds the pipe for 2 clocks, but a DFMA whose three sources are three
// DISTINCT registers needs a THIRD clock to collect its operands, unless the operand-reuse cache
// serves one of them (the previous instruction of the warp read the same register in the same
// operand slot).  acc_jerk has 11 such DFMAs among its 32 FP64 instructions per pair (r.v chain,
// v - alpha r, the six accumulations).  ptxas, scheduling the unrolled loop of pair_kernel as one
// basic block, interleaves them with everything else and leaves 9 of them uncached: 64 + 9 clocks
// per pair modelled, 74.4 measured.  Here the G = W x U pairs of a group of U rows are written
// operation by operation (dependent operations are G instructions apart), and the operand-sharing
// DFMAs live in basic blocks of their own: a branch on a kernel argument that is always 1
// (PairArgs::one; `if (one != 0)`, GMODE bit 3, or a loop with that trip count) keeps ptxas from
// merging the blocks, and inside such a block it groups the DFMAs by shared operand and flags the
// reuse itself: 2.0 (3 x 2 pairs) to 3.2 (2 x 4) uncached three-register DFMAs per pair instead of 9.
//
// An Op opts in with
//   enum { GROUPED = 1, GW, GU, GNT, GMODE };  particles per thread, rows per group, threads per CTA,
//                                          mode bits handed to the phases (bit 3 is read here too)
//   enum { GALT, GW2, GU2, GMODE2, GCOST2_PERMILLE };   optional second shape (GroupedAlt below)
//   struct PV;                             what phase 1 hands to phase 2 for one pair
//   group_phase1<W, U, MODE>(is, rows, pv, prm, one)   everything up to the accumulation, W x U pairs
//   group_phase2<W, U, MODE>(pv, acc, prm)  the accumulation DFMAs of the group (fenced by the kernel)
// Partial tiles (the end of a j range) take Op::pair row by row.
// ---------------------------------------------------------------------------------------
template <class Op, typename = void> struct Grouped { enum { value = 0, W = 1, U = 1, NT = 256, MODE = 0 }; };
template <class Op> struct Grouped<Op, typename std::enable_if<(Op::GROUPED != 0)>::type> {
    enum { value = 1, W = Op::GW, U = Op::GU, NT = Op::GNT, MODE = Op::GMODE };
};
// A second group shape (Op::GALT: GW2 x GU2, GMODE2) with fewer particles per CTA: the launch plan takes
// it where the first shape's i-blocks would leave SMs idle or half-filled (small and medium ni).
template <class Op, typename = void> struct GroupedAlt { enum { value = 0, W = 1, U = 1, MODE = 0, COST_PERMILLE = 1000 }; };
template <class Op> struct GroupedAlt<Op, typename std::enable_if<(Op::GROUPED != 0 && Op::GALT != 0)>::type> {
    enum { value = 1, W = Op::GW2, U = Op::GU2, MODE = Op::GMODE2, COST_PERMILLE = Op::GCOST2_PERMILLE };
};

template <class Op, int NT, int W, int U, int MODE, int TJ, int STAGES, bool MULTI = false>
__global__ void __launch_bounds__(NT) pair_kernel_grouped(const __grid_constant__ PairArgs<Op> a)
{
    typedef typename Op::real T;
    constexpr int NJP = Packed<Op>::NJP;
    constexpr int TILE_ELEMS = TJ * NJP;
    constexpr int G = W * U;
    static_assert(TJ % U == 0, "a tile holds whole row groups");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tiles = reinterpret_cast<T*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * TILE_ELEMS * sizeof(T));

    const int tid = threadIdx.x;
    const long long ibase = (long long)blockIdx.x * (NT * W);

    T is[W][Op::NI];
    T acc[W][Op::NA];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        long long i = ibase + (long long)w * NT + tid;
        if (i > a.ni - 1) i = a.ni - 1;  // clamp: computes a duplicate, never stored
        Op::load_i(a.i.p, i, is[w]);
        Op::zero(acc[w]);
    }

    const long long jlo = a.j0 + (long long)blockIdx.y * a.jchunk;
    long long jhi = jlo + a.jchunk;
    if (jhi > a.j1) jhi = a.j1;
    const int ntiles = (jhi > jlo) ? (int)((jhi - jlo + TJ - 1) / TJ) : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t) {  // elected thread: start the bulk copy of tile t
        const int s = t % STAGES;
        int cnt;
        const T* src = locate_tile<Op, TJ, MULTI>(a, jlo + (long long)t * TJ, jhi, cnt);
        const unsigned bytes = (unsigned)cnt * Packed<Op>::ROW_BYTES;
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(tiles + s * TILE_ELEMS, src, bytes, &full[s]);
    };
    if (tid == 0) {
        for (int t = 0; t < STAGES && t < ntiles; ++t) issue(t);
    }
    const int one = a.one;

    auto grouped_tile = [&](const T* sj) {
#pragma unroll 1
        for (int j = 0; j < TJ; j += U) {
            T rows[U][NJP];
            typename Op::PV pv[G];
#pragma unroll
            for (int u = 0; u < U; ++u) load_row<Op>(sj + (j + u) * NJP, rows[u]);
            Op::template group_phase1<W, U, MODE>(is, rows, pv, a.prm, one);
            if (MODE & 8) {             // block 2 behind a never-skipped branch (cheaper than a one-trip loop)
                if (one != 0) Op::template group_phase2<W, U, MODE>(pv, acc, a.prm);
            } else {
#pragma unroll 1
                for (int z = 0; z < one; ++z) Op::template group_phase2<W, U, MODE>(pv, acc, a.prm);
            }
        }
    };
    auto plain_tile = [&](const T* sj, int cnt) {
        for (int j = 0; j < cnt; ++j) {
            T row[NJP];
            load_row<Op>(sj + j * NJP, row);
#pragma unroll
            for (int w = 0; w < W; ++w) Op::pair(is[w], row, acc[w], a.prm);
        }
    };

    if (!MULTI) {
        // one buffer: every tile but possibly the last is full -- the hot loop has no case split
        const long long span = jhi > jlo ? jhi - jlo : 0;       // a chunk beyond the end of the rows is empty
        const int nfull = (int)(span / TJ);
        for (int t = 0; t < nfull; ++t) {
            const int s = t % STAGES;
            mbar_wait(&full[s], (unsigned)(t / STAGES) & 1u);
            grouped_tile(tiles + s * TILE_ELEMS);
            __syncthreads();  // every warp is done with stage s -> it may be refilled
            if (tid == 0 && t + STAGES < ntiles) issue(t + STAGES);
        }
        if (nfull < ntiles) {
            const int s = nfull % STAGES;
            mbar_wait(&full[s], (unsigned)(nfull / STAGES) & 1u);
            plain_tile(tiles + s * TILE_ELEMS, (int)(span - (long long)nfull * TJ));
        }
    } else {
        for (int t = 0; t < ntiles; ++t) {
            const int s = t % STAGES;
            mbar_wait(&full[s], (unsigned)(t / STAGES) & 1u);
            const T* sj = tiles + s * TILE_ELEMS;
            int cnt;
            locate_tile<Op, TJ, MULTI>(a, jlo + (long long)t * TJ, jhi, cnt);
            if (cnt == TJ) grouped_tile(sj);
            else plain_tile(sj, cnt);
            __syncthreads();
            if (tid == 0 && t + STAGES < ntiles) issue(t + STAGES);
        }
    }

#pragma unroll
    for (int w = 0; w < W; ++w) {
        const long long i = ibase + (long long)w * NT + tid;
        if (i < a.ni) {
            if (a.partial != nullptr) {
                T* dst = a.partial + ((long long)(a.slot0 + blockIdx.y) * Op::NA) * a.ni + i;
#pragma unroll
                for (int k = 0; k < Op::NA; ++k) dst[(long long)k * a.ni] = acc[w][k];
            } else {
                Op::finish(a.i.p, i, acc[w], a.prm, a.out.p);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// pair_kernel_defer: the same sweep for Ops with a rare expensive case (see the top of the
// file).  One i-particle per thread.  The j loop is warp-uniform (lane split: a predicate
// instead of a per-lane trip count) because pushing into the warp's ring is a warp collective.
//
// The kernel is glue around two functions that are deliberately NOT inlined:
//   sweep_rows     the hot loop over the rows of a tile; call-free, so its registers are
//                  allocated on their own terms.  (Inlined next to the solver call, ptxas homed
//                  every loop-carried value -- i-state, accumulators -- in local memory for
//                  the whole loop: 29 Gpair/s in fp64.)  It returns when the tile is finished or
//                  the ring holds a full warp of deferred pairs;
//   the solver     (Op::pair_slow's own __noinline__ function, e.g. kepler_propagate).
// The state that crosses those calls lives in a small struct in local memory, read once per
// call (a tile of 128 rows) and written back once.
// ---------------------------------------------------------------------------------------
template <class Op> struct SweepState {
    typename Op::real is[Op::NI];
    typename Op::real acc[Op::NA];
    int qhead, qcount;   // the warp's ring of deferred pairs: first slot, pairs waiting
};

// Rows [k, steps) of one tile (lane split: row k*js + jsub) until the ring is full.  Returns
// the next k.
template <class Op, bool LANE_SPLIT>
__device__ __noinline__ int sweep_rows(SweepState<Op>* st, const typename Op::real* sj, int cnt, int k, int steps,
                                       int jsl, int jsub, int lane, typename Op::real* qdat, int* qown,
                                       typename Op::Params prm)
{
    typedef typename Op::real T;
    constexpr int NJP = Packed<Op>::NJP, ND = Op::ND, NA = Op::NA, QCAP = 64;
    T is[Op::NI], acc[NA];
#pragma unroll
    for (int q = 0; q < Op::NI; ++q) is[q] = st->is[q];
#pragma unroll
    for (int q = 0; q < NA; ++q) acc[q] = st->acc[q];
    const int qhead = st->qhead;
    int qcount = st->qcount;
    for (; k < steps && qcount < 32; ++k) {
        const int j = LANE_SPLIT ? (k << jsl) + jsub : k;
        bool need = false;
        T d[ND];
        if (!LANE_SPLIT || j < cnt) {
            T row[NJP];
            load_row<Op>(sj + j * NJP, row);
            need = Op::pair_fast(is, row, acc, prm, d);
        }
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (m != 0u) {
            if (need) {
                const int slot = (qhead + qcount + __popc(m & ((1u << lane) - 1u))) & (QCAP - 1);
#pragma unroll
                for (int q = 0; q < ND; ++q) qdat[slot * ND + q] = d[q];
                qown[slot] = lane;
            }
            qcount += __popc(m);
        }
    }
#pragma unroll
    for (int q = 0; q < NA; ++q) st->acc[q] = acc[q];
    st->qcount = qcount;
    return k;
}

// One deferred pair: description in slot[0..ND), result out in slot[0..NA).  Inlined into the
// glue kernel on purpose: as a __noinline__ function that itself calls the __noinline__ solver,
// nvcc 12.9 produced wrong results for softened pairs (bisected on a B200 against the golden
// vectors); the hot loop is unaffected either way, it lives in sweep_rows.
template <class Op>
__device__ __forceinline__ void deferred_pair(typename Op::real* slot, typename Op::Params prm, long long owner,
                                              long long ni)
{
    typedef typename Op::real T;
    T d[Op::ND], c[Op::NA];
#pragma unroll
    for (int k = 0; k < Op::ND; ++k) d[k] = slot[k];
    // a pair the bounded solver gave up on contributes 0 here and goes to the clean-up launch
    // (not for the clamped duplicates beyond ni: they are never stored)
    if (!Op::pair_slow(d, prm, c) && owner < ni) Op::defer_to_cleanup(d, owner, prm);
#pragma unroll
    for (int k = 0; k < Op::NA; ++k) slot[k] = c[k];
}

// __launch_bounds__(NT, 2): with (NT) alone ptxas settled for 80 registers and spilled inside
// sweep_rows to reach three CTAs per SM -- 147 -> 84 Gpair/s for sakura flag 1; two CTAs with 128
// registers and no spill in the hot function is the better trade.
template <class Op, int NT, int TJ, int STAGES, bool LANE_SPLIT, bool MULTI = false>
__global__ void __launch_bounds__(NT, 2) pair_kernel_defer(const __grid_constant__ PairArgs<Op> a)
{
    typedef typename Op::real T;
    typedef PairSmemDefer<Op, NT, TJ, STAGES, true> SM;
    constexpr int NJP = Packed<Op>::NJP;
    constexpr int TILE_ELEMS = TJ * NJP;
    constexpr int ND = Op::ND, NA = Op::NA, QCAP = SM::QCAP;
    static_assert(NA <= ND, "results are handed back through the ring slots");
    static_assert(QCAP == 64, "sweep_rows assumes a 64-slot ring");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tiles = reinterpret_cast<T*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * TILE_ELEMS * sizeof(T));

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    T* qdat = reinterpret_cast<T*>(smem_raw + SM::QUEUE_OFF + (tid >> 5) * SM::WARP_BYTES);  // [QCAP][ND]
    int* qown = reinterpret_cast<int*>(qdat + QCAP * ND);                                     // [QCAP]

    const int jsl = LANE_SPLIT ? a.js_log2 : 0;
    const int js = 1 << jsl;
    const int ipw = 32 >> jsl;                        // lane layout: see pair_kernel
    const int jsub = (tid & 31) >> (5 - jsl);
    const int islot = (tid >> 5) * ipw + (tid & (ipw - 1));
    const int slots = NT >> jsl;
    const long long ibase = (long long)blockIdx.x * slots;

    SweepState<Op> st;
    {
        long long i = ibase + islot;
        if (i > a.ni - 1) i = a.ni - 1;  // clamp: computes a duplicate, never stored
        Op::load_i(a.i.p, i, st.is);
        Op::zero(st.acc);
        st.qhead = 0;
        st.qcount = 0;
    }

    const long long jlo = a.j0 + (long long)blockIdx.y * a.jchunk;
    long long jhi = jlo + a.jchunk;
    if (jhi > a.j1) jhi = a.j1;
    const int ntiles = (jhi > jlo) ? (int)((jhi - jlo + TJ - 1) / TJ) : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t) {  // elected thread: start the bulk copy of tile t
        const int s = t % STAGES;
        int cnt;
        const T* src = locate_tile<Op, TJ, MULTI>(a, jlo + (long long)t * TJ, jhi, cnt);
        const unsigned bytes = (unsigned)cnt * Packed<Op>::ROW_BYTES;
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(tiles + s * TILE_ELEMS, src, bytes, &full[s]);
    };
    if (tid == 0) {
        for (int t = 0; t < STAGES && t < ntiles; ++t) issue(t);
    }

    // run the n (<= 32) oldest deferred pairs of this warp, one per lane, and hand the results
    // to their owner lanes in ring order (for any one particle: its own j order)
    auto drain = [&](int n) {
        __syncwarp();
        const int qhead = st.qhead;
        if (lane < n) {
            const int slot = (qhead + lane) & (QCAP - 1);
            const int ol = qown[slot];                       // the lane that owns this pair
            const long long owner = ibase + (tid >> 5) * ipw + (ol & (ipw - 1));
            deferred_pair<Op>(qdat + slot * ND, a.prm, owner, a.ni);
        }
        __syncwarp();
        for (int e = 0; e < n; ++e) {
            const int slot = (qhead + e) & (QCAP - 1);
            if (qown[slot] == lane) {
                T c[NA];
#pragma unroll
                for (int k = 0; k < NA; ++k) c[k] = qdat[slot * ND + k];
                Op::combine(st.acc, c);
            }
        }
        __syncwarp();
        st.qhead = (qhead + n) & (QCAP - 1);
        st.qcount -= n;
    };

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(&full[s], (unsigned)(t / STAGES) & 1u);
        const T* sj = tiles + s * TILE_ELEMS;
        int cnt;
        locate_tile<Op, TJ, MULTI>(a, jlo + (long long)t * TJ, jhi, cnt);
        const int steps = (cnt + js - 1) >> jsl;
        int k = 0;
        while (true) {
            k = sweep_rows<Op, LANE_SPLIT>(&st, sj, cnt, k, steps, jsl, jsub, lane, qdat, qown, a.prm);
            if (st.qcount >= 32) drain(32);
            else break;                      // ring not full: the tile is finished
        }
        __syncthreads();  // every warp is done with stage s -> it may be refilled
        if (tid == 0 && t + STAGES < ntiles) issue(t + STAGES);
    }
    while (st.qcount > 0) drain(st.qcount < 32 ? st.qcount : 32);

    T acc[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) acc[k] = st.acc[k];
    if (LANE_SPLIT) {
        for (int off = 16; off >= ipw; off >>= 1) {
            T other[NA];
#pragma unroll
            for (int k = 0; k < NA; ++k) other[k] = __shfl_xor_sync(0xffffffffu, acc[k], off);
            Op::combine(acc, other);
        }
    }

    if (jsub == 0) {
        const long long i = ibase + islot;
        if (i < a.ni) {
            if (a.partial != nullptr) {
                T* dst = a.partial + ((long long)(a.slot0 + blockIdx.y) * NA) * a.ni + i;
#pragma unroll
                for (int k = 0; k < NA; ++k) dst[(long long)k * a.ni] = acc[k];
            } else {
                Op::finish(a.i.p, i, acc, a.prm, a.out.p);
            }
        }
    }
}

// The kernel an Op runs on: <throughput shape> and <lane split>.
template <class Op, bool LANE_SPLIT> struct KernelOf {
    typedef void (*Fn)(const PairArgs<Op>);
    template <int NT, int TJ, int STAGES, bool MULTI = false, bool ALT = false> static Fn get()
    {
        if constexpr (Defers<Op>::value != 0) return pair_kernel_defer<Op, NT, TJ, STAGES, LANE_SPLIT, MULTI>;
        else if constexpr (!LANE_SPLIT && Grouped<Op>::value != 0 && ALT && GroupedAlt<Op>::value != 0)
            return pair_kernel_grouped<Op, NT, GroupedAlt<Op>::W, GroupedAlt<Op>::U, GroupedAlt<Op>::MODE, TJ, STAGES, MULTI>;
        else if constexpr (!LANE_SPLIT && Grouped<Op>::value != 0)
            return pair_kernel_grouped<Op, NT, Grouped<Op>::W, Grouped<Op>::U, Grouped<Op>::MODE, TJ, STAGES, MULTI>;
        else return pair_kernel<Op, NT, LANE_SPLIT ? 1 : Op::WPT, TJ, STAGES, LANE_SPLIT, MULTI>;
    }
};

// ---------------------------------------------------------------------------------------
// finalize_kernel: combine nslots raw accumulator sets and apply the epilogue.
// ---------------------------------------------------------------------------------------
template <class Op>
__global__ void finalize_kernel(InRefs<typename Op::real> iarr, long long ni,
                                const typename Op::real* __restrict__ partial, int nslots,
                                OutRefs<typename Op::real> out, typename Op::Params prm)
{
    typedef typename Op::real T;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= ni) return;
    T acc[Op::NA];
#pragma unroll
    for (int k = 0; k < Op::NA; ++k) acc[k] = partial[(long long)k * ni + i];
    for (int s = 1; s < nslots; ++s) {
        T other[Op::NA];
#pragma unroll
        for (int k = 0; k < Op::NA; ++k) other[k] = partial[((long long)s * Op::NA + k) * ni + i];
        Op::combine(acc, other);
    }
    Op::finish(iarr.p, i, acc, prm, out.p);
}

// ---------------------------------------------------------------------------------------
// Host side: launch plan and launcher.
// ---------------------------------------------------------------------------------------
struct Plan {
    int lane_split;  // 0: WPT particles per thread, lanes independent; 1: JS lanes per particle;
                     // 2: as 0 with the Op's second group shape (GroupedAlt; ops without one run 0)
    int js_log2;     // lane-split only
    int jg;          // number of j chunks over blockIdx.y (>1 -> workspace + finalize)
};

struct DeviceInfo {
    int sm_count;
};

template <class Op> struct Tune {
    enum {
        TJ = 128, STAGES = 4, NT_SPLIT = 128,
        NT = Grouped<Op>::value ? (int)Grouped<Op>::NT : 256,                 // throughput shape
        WPT = Grouped<Op>::value ? (int)Grouped<Op>::W : (Defers<Op>::value ? 1 : (int)Op::WPT),
        ALT = Grouped<Op>::value && GroupedAlt<Op>::value,                    // second throughput shape
        WPT2 = ALT ? (int)GroupedAlt<Op>::W : WPT,
        SMEM = (int)PairSmemDefer<Op, NT, TJ, STAGES>::BYTES
    };
};

// Issue cost of one pair on the kernel's main pipe, in warp instructions (FP64 or FP32); the
// launch-plan model below only needs it to within ~20 %.  Specialised next to each Op.
template <class Op> struct OpCost { enum { value = 32 }; };

// Launch shape: the candidate with the smallest predicted time.
//
// The model was fitted to forced-plan sweeps on a B200 (tools/plan_probe.py,
// profiles/r01_plan_probe_accjerk.txt).  What it encodes:
//   * throughput shape (WPT particles per thread, unrolled j loop, j range cut into g chunks
//     over blockIdx.y).  A CTA's 8 warps keep an SM's pipe busy on their own, and CTAs are
//     handed out greedily, so the SM -- not the CTA slot -- is the unit of load balance:
//     a launch lasts as long as the busiest SM, which gets ceil(CTAs / SMs) chunks (measured to
//     within 2 % at N = 4096 ... 16384).  N = 4096 (8 i-blocks): 16 chunks of 2 tiles on 128 SMs
//     take 57 us where the old wave-count rule chose a 32-lane split at 190 us;
//   * lane split (2^js lanes of a warp share one particle, 128-thread CTAs).  One warp per
//     SM sub-partition per CTA: bounded by the dependent-instruction latency of the rows a
//     lane owns, by the pipe when several CTAs share an SM, and by a per-tile hand-shake;
//   * g > 1 costs a finalize launch that reads g accumulator sets.
template <class Op>
inline Plan choose_plan(const DeviceInfo& dev, long long ni, long long nj)
{
    typedef Tune<Op> U;
    typedef typename Op::real T;
    Plan best_plan = {0, 0, 1};
    const int TJ = U::TJ;
    const double sms = dev.sm_count > 0 ? dev.sm_count : 148;
    const double clk_per_us = 1965.0;
    const bool dp = sizeof(T) == 8;
    const double cpw = OpCost<Op>::value * (dp ? 2.33 : 1.3);   // pipe clocks per warp per row
    const double lat = OpCost<Op>::value * (dp ? 6.25 : 5.0);   // clocks per row for a lone warp
    const long long tiles = nj > 0 ? (nj + TJ - 1) / TJ : 1;
    const double fin_bytes_per_us = 2.0e6;
    auto finalize_us = [&](long long g) {
        return g > 1 ? 9.0 + (double)g * (double)ni * Op::NA * sizeof(T) / fin_bytes_per_us : 0.0;
    };
    double best = 1e300;

    for (int shape = 0; shape <= (U::ALT ? 1 : 0); ++shape) {   // throughput shape(s)
        const int wpt = shape ? (int)U::WPT2 : (int)U::WPT;
        const double cost = shape ? GroupedAlt<Op>::COST_PERMILLE * 1e-3 : 1.0;
        const long long IB = (long long)U::NT * wpt;
        const long long iblocks = (ni + IB - 1) / IB;
        const double tile_us = (double)TJ * wpt * cpw * cost * ((double)U::NT / 32 / 4) / clk_per_us;
        const long long maxg = tiles < 64 ? tiles : 64;
        for (long long g = 1; g <= maxg; ++g) {
            const long long tpc = (tiles + g - 1) / g;            // tiles per chunk
            const long long chunks = (tiles + tpc - 1) / tpc;     // non-empty chunks
            if (chunks != g) continue;                            // same split as a smaller g
            const double ctas = (double)iblocks * chunks;
            const double per_sm = (double)((long long)((ctas + sms - 1) / sms));
            // equal CTAs handed out greedily: the busiest SM runs ceil(CTAs / SMs) of them
            const double busiest = per_sm * (tpc + 0.1);       // + prologue/epilogue of each CTA
            const double t = 8.0 + busiest * tile_us + finalize_us(chunks);
            if (t < best) { best = t; best_plan.lane_split = shape ? 2 : 0; best_plan.js_log2 = 0; best_plan.jg = (int)g; }
        }
    }
    for (int js = 0; js <= 5 && (TJ >> js) >= 4; ++js) {          // lane split
        const long long per_cta = U::NT_SPLIT >> js;
        const long long icta = (ni + per_cta - 1) / per_cta;
        for (long long g = 1; g <= 64 && g <= tiles; g *= 2) {
            const long long tpc = (tiles + g - 1) / g;
            const double rows_per_lane = (double)tpc * TJ / (1 << js);
            const double ctas = (double)icta * g;
            const double per_sm = (double)((long long)((ctas + sms - 1) / sms));
            const double pipe = per_sm * rows_per_lane * cpw * 1.15;
            const double chain = rows_per_lane * lat;
            // every CTA walks the same tiles at the same time: the more of them (and the fewer rows a
            // lane takes from each), the longer the hand-over of a tile
            const double per_tile = 0.3 + ctas / 1000.0 + 0.04 * js;
            const double t = 13.0 + 0.3 * js + (per_sm - 1.0) + (pipe > chain ? pipe : chain) / clk_per_us
                             + per_tile * tpc + finalize_us(g);
            if (t < best) { best = t; best_plan.lane_split = 1; best_plan.js_log2 = js; best_plan.jg = (int)g; }
        }
    }
    return best_plan;
}

// Sweep rows [j0, j1) of `jpack` for all ni particles.
//   partial == nullptr (requires plan.jg == 1): epilogue applied, outputs written.
//   partial != nullptr: raw accumulators go to slots [slot0, slot0 + plan.jg).
template <class Op>
inline cudaError_t launch_pairs(const Plan& plan, const InRefs<typename Op::real>& iarr, long long ni,
                                const typename Op::real* jpack, long long j0, long long j1,
                                const typename Op::Params& prm, typename Op::real* partial, int slot0,
                                const OutRefs<typename Op::real>& out, cudaStream_t stream,
                                const Segments<typename Op::real>* seg = nullptr)
{
    typedef Tune<Op> U;
    PairArgs<Op> a;
    a.i = iarr;
    a.ni = ni;
    a.jpack = jpack;
    a.seg.nseg = 0;
    if (seg) {                  // rows of several owners: sweep the logical tiles [0, tile0[nseg])
        a.seg = *seg;
        j0 = 0;
        j1 = (long long)seg->tile0[seg->nseg] * U::TJ;
    }
    a.j0 = j0;
    a.j1 = j1;
    const long long rows = j1 - j0;
    long long chunk = (rows + plan.jg - 1) / plan.jg;
    chunk = (chunk + U::TJ - 1) / U::TJ * U::TJ;
    if (chunk < U::TJ) chunk = U::TJ;
    a.jchunk = chunk;
    a.js_log2 = plan.js_log2;
    a.slot0 = slot0;
    a.partial = partial;
    a.out = out;
    a.prm = prm;
    a.one = 1;
    if (ni <= 0) return cudaSuccess;
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    const bool multi = seg != nullptr;
    // function attributes are per device and per instantiation
    static int attr_device[6] = {-1, -1, -1, -1, -1, -1};
    auto prepare = [&](typename KernelOf<Op, false>::Fn k, size_t smem, int which) {
        if (attr_device[which] != dev_id) {
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_device[which] = dev_id;
        }
    };
    if (plan.lane_split != 1) {
        const bool alt = plan.lane_split == 2 && U::ALT;
        auto k = alt ? (multi ? KernelOf<Op, false>::template get<U::NT, U::TJ, U::STAGES, true, true>()
                              : KernelOf<Op, false>::template get<U::NT, U::TJ, U::STAGES, false, true>())
                     : (multi ? KernelOf<Op, false>::template get<U::NT, U::TJ, U::STAGES, true>()
                              : KernelOf<Op, false>::template get<U::NT, U::TJ, U::STAGES, false>());
        const size_t smem = U::SMEM;
        prepare(k, smem, (alt ? 4 : 0) + (multi ? 1 : 0));
        const long long per_cta = (long long)U::NT * (alt ? (int)U::WPT2 : (int)U::WPT);
        dim3 grid((unsigned)((ni + per_cta - 1) / per_cta), (unsigned)plan.jg);
        k<<<grid, U::NT, smem, stream>>>(a);
    } else {
        auto k = multi ? KernelOf<Op, true>::template get<U::NT_SPLIT, U::TJ, U::STAGES, true>()
                       : KernelOf<Op, true>::template get<U::NT_SPLIT, U::TJ, U::STAGES, false>();
        const size_t smem = PairSmemDefer<Op, U::NT_SPLIT, U::TJ, U::STAGES>::BYTES;
        prepare(k, smem, multi ? 3 : 2);
        const long long per_cta = (long long)(U::NT_SPLIT >> plan.js_log2);
        dim3 grid((unsigned)((ni + per_cta - 1) / per_cta), (unsigned)plan.jg);
        k<<<grid, U::NT_SPLIT, smem, stream>>>(a);
    }
    return cudaGetLastError();
}

template <class Op>
inline cudaError_t launch_pack(const InRefs<typename Op::real>& jarr, long long nj,
                               typename Op::real* packed, cudaStream_t stream)
{
    if (nj <= 0) return cudaSuccess;
    long long blocks = (nj + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_j_kernel<Op><<<(unsigned)blocks, 256, 0, stream>>>(jarr, nj, packed);
    return cudaGetLastError();
}

template <class Op>
inline cudaError_t launch_finalize(const InRefs<typename Op::real>& iarr, long long ni,
                                   const typename Op::real* partial, int nslots,
                                   const OutRefs<typename Op::real>& out, const typename Op::Params& prm,
                                   cudaStream_t stream)
{
    if (ni <= 0) return cudaSuccess;
    finalize_kernel<Op><<<(unsigned)((ni + 255) / 256), 256, 0, stream>>>(iarr, ni, partial, nslots, out, prm);
    return cudaGetLastError();
}

}  // namespace tupan
