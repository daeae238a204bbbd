"""Device-resident integrators with the interface of the reference's ``tupan.integrator``.

``Integrator(eta, time, ps, method=...)`` mirrors ``tupan/integrator/__init__.py:102-157``:
``initialize / evolve_step(t_end) / finalize``, ``.time``, ``.particle_system``; the same
method names (``hermite2..8``, ``ahermite2..8``, ``siaXYs|a.dkd|kdk``, ``nreg``, ``anreg``,
``sakura``, ``asakura``).  What differs is where the state lives: the reference keeps numpy
arrays on the host, copies the particle system every step (``ps.copy()``, hermite.py:29) and
reads ``abs(ps.tstep).min()`` back to choose the step; here the SoA state is uploaded once
and stays in HBM, every O(N) update is one of the kernels of ``csrc/k_update.cu`` (Part 3 of
``include/libtupan_cuda.h``), the force evaluations are the pair kernels (device entry
points, or the i-sharded multi-GPU path when a process group is given), and the step size is
derived on the device -- a whole Hermite / SIA step is enqueued on one stream without a host
round trip.  ``.time`` / ``.particle_system`` synchronise and download.

The O(N) updates follow the reference's operation order exactly, so the only numerical
difference from the reference is the summation order inside the pair kernels.

The hierarchical ``sia..h`` methods split the system by time-step at every level
(data-dependent sub-system sizes): their recursion is driven from the host like the
reference's, with the sub-systems gathered on the device (torch indexing as plumbing) and
rectangular ``acc`` calls between the slow and fast sets.  Post-Newtonian kicks raise.
There is no CPU path: the CUDA library must be loadable and a GPU present.
"""
import ctypes
import gc

import numpy as np
import torch
import torch.distributed as dist

from . import backend
from .device import KERNEL_INPUTS, KERNEL_OUTPUTS, run as run_kernel
from .sharded import ShardedKernel, shard_bounds

CTL_T_CURR, CTL_TAU, CTL_T_END, CTL_ETA, CTL_NSTEPS, CTL_DONE, CTL_TAU_BASE, CTL_MIN_TS = range(8)
RED_SUM, RED_KINETIC, RED_HALF_DOT, RED_SAKURA_DT, RED_ABS_MIN, RED_ABS_MAX, RED_DOT, RED_MOMENT = range(8)

R3, V3, A3, J3, S3, C3 = (("rx", "ry", "rz"), ("vx", "vy", "vz"), ("ax", "ay", "az"),
                          ("jx", "jy", "jz"), ("sx", "sy", "sz"), ("cx", "cy", "cz"))
DERIVS = A3 + J3 + S3 + C3

# (A, B) of the reference's SIAxy.coefs (integrator/sia.py:302-303, 360-362, 425-428, 497-501,
# 576-581, 662-668, 755-762, 855-864; Yoshida 1990, Omelyan et al. 2003, Blanes & Moan 2002,
# Kahan & Li 1997)
SIA_COEFS = {
    "sia21": ([1.0], [0.5]),
    "sia22": ([0.5], [0.1931833275037836, 0.6136333449924328]),
    "sia43": ([1.3512071919596575, -1.7024143839193150], [0.6756035959798288, -0.17560359597982877]),
    "sia44": ([0.7123418310626056, -0.21234183106260562],
              [0.1786178958448091, -0.06626458266981843, 0.7752933736500186]),
    "sia45": ([-0.0844296195070715, 0.354900057157426, 0.459059124699291],
              [0.2750081212332419, -0.1347950099106792, 0.35978688867743724]),
    "sia46": ([0.209515106613362, -0.143851773179818, 0.434336666566456],
              [0.0792036964311957, 0.353172906049774, -0.0420650803577195, 0.21937695575349958]),
    "sia67": ([0.7845136104775573, 0.23557321335935813, -1.177679984178871, 1.3151863206839112],
              [0.39225680523877865, 0.5100434119184577, -0.47105338540975644, 0.06875316825252015]),
    "sia69": ([0.39103020330868477, 0.334037289611136, -0.7062272811875614, 0.08187754964805945,
               0.7985644772393624],
              [0.19551510165434238, 0.3625337464599104, -0.1860949957882127, -0.31217486576975095,
               0.44022101344371095]),
}


def operator_sequence(outer, inner):
    """The palindromic composition every SIAxy.dkd / .kdk / .bridge_sf spells out (e.g.
    sia.py:441-453): the two weight lists interleaved starting with `outer`, mirrored about
    the last one.  -> [(is_outer, weight)]"""
    seq = []
    for i in range(max(len(outer), len(inner))):
        if i < len(outer):
            seq.append((True, outer[i]))
        if i < len(inner):
            seq.append((False, inner[i]))
    return seq + seq[-2::-1]


class DeviceState(object):
    """SoA particle arrays in HBM (attribute names of particles/body.py:26-39)."""

    BASE = ("mass", "eps2", "rx", "ry", "rz", "vx", "vy", "vz", "time", "tstep")

    def __init__(self, ps, device, lo=0, hi=None):
        self.device = torch.device(device)
        self.lo, self.hi = lo, ps.n if hi is None else hi
        self.n = self.hi - self.lo
        self.np_dtype = np.dtype(ps.mass.dtype)
        self.dtype = torch.float64 if self.np_dtype == np.float64 else torch.float32
        self.t = {}
        for k in self.BASE:
            self.t[k] = self._up(np.asarray(getattr(ps, k), self.np_dtype))
        itype = np.int64 if self.np_dtype == np.float64 else np.int32      # same bits as the UINT array
        self.t["nstep"] = self._up(np.ascontiguousarray(getattr(ps, "nstep")).view(itype))

    def _up(self, a):
        return torch.from_numpy(np.ascontiguousarray(a[self.lo:self.hi])).to(self.device).contiguous()

    def need(self, *names):
        for k in names:
            if k not in self.t:
                self.t[k] = torch.zeros(self.n, dtype=self.dtype, device=self.device)
        return [self.t[k] for k in names]

    def __getitem__(self, k):
        return self.t[k]

    def download(self, ps, names=None):
        """Write the device arrays back into the host container's arrays (in place)."""
        for k, t in self.t.items():
            if names is not None and k not in names:
                continue
            if k.endswith("0") or k.startswith("_"):
                continue
            h = t.cpu().numpy()
            if k not in ps.__dict__:
                if not hasattr(ps, "register_auxiliary_attribute"):
                    continue
                ps.register_auxiliary_attribute(k, "real")
            dst = getattr(ps, k)
            if k in ("nstep", "id"):
                dst = dst.view(h.dtype)
            dst[self.lo:self.hi] = h


class _Lib(object):
    """Thin caller of Part 3 of include/libtupan_cuda.h on torch tensors."""

    def __init__(self, prec):
        self.lib = backend.require_gpu(prec)
        self._ptrs = {}

    PTR_CACHE_MAX = 4096

    def ptrs(self, tensors):
        key = tuple(t.data_ptr() for t in tensors)
        p = self._ptrs.get(key)
        if p is None:
            for t in tensors:
                if not (t.is_cuda and t.is_contiguous()):
                    raise TypeError("contiguous CUDA tensors required")
            if len(self._ptrs) >= self.PTR_CACHE_MAX:
                # the hierarchical SIA recursion makes fresh sub-system tensors every step: the cache
                # only pays for the long-lived state arrays, so it is simply dropped when it fills up
                self._ptrs.clear()
            p = self._ptrs[key] = (ctypes.c_void_p * len(key))(*key)
        return p

    def check_async(self):
        """Failures of asynchronous device work that only leave a counter behind: a peer barrier of
        the p2p transport that gave up waiting, a Kepler pair beyond 2^27 sub-steps.  Called where
        the integrators synchronise anyway (clock read-back, state download)."""
        hits = self.lib.tupan_cuda_peer_timeouts()
        if hits:
            raise backend.TupanCudaError("peer barrier: %d wait(s) timed out -- forces of this run are not "
                                         "trustworthy" % hits)
        hits = self.lib.tupan_cuda_kepler_limit_hits()
        if hits:
            raise backend.TupanCudaError("%d Kepler pair(s) needed more than 2^27 sub-steps" % hits)

    @staticmethod
    def stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def ok(self, rc, what):
        if rc != 0:
            backend.check(self.lib, what)
            raise backend.TupanCudaError("%s failed with code %d" % (what, rc))

    def step_begin(self, ctl, dmin=None):
        self.ok(self.lib.tupan_cuda_step_begin_dev(ctl.data_ptr(), dmin.data_ptr() if dmin is not None else None,
                                                   self.stream()), "step_begin")

    def predict(self, order, n, rv, rv0, d0, ctl):
        self.ok(self.lib.tupan_cuda_hermite_predict_dev(order, n, self.ptrs(rv), self.ptrs(rv0), self.ptrs(d0),
                                                        ctl.data_ptr(), self.stream()), "hermite_predict")

    def correct(self, order, n, rv, rv0, d0, d1, ctl):
        self.ok(self.lib.tupan_cuda_hermite_correct_dev(order, n, self.ptrs(rv), self.ptrs(rv0), self.ptrs(d0),
                                                        self.ptrs(d1), ctl.data_ptr(), self.stream()),
                "hermite_correct")

    def pn_kick(self, phase, n, arrays, c_outer, c_inner, ctl):
        self.ok(self.lib.tupan_cuda_pn_kick_dev(phase, n, self.ptrs(arrays), c_outer, c_inner, ctl.data_ptr(),
                                                self.stream()), "pn_kick")

    def axpy(self, n, y, x, c_outer=1.0, c_inner=1.0, ctl=None):
        self.ok(self.lib.tupan_cuda_axpy_dev(len(y), n, self.ptrs(y), self.ptrs(x), c_outer, c_inner,
                                             ctl.data_ptr() if ctl is not None else None, self.stream()), "axpy")

    def scale(self, n, y, x, denom):
        self.ok(self.lib.tupan_cuda_scale_dev(len(y), n, self.ptrs(y), self.ptrs(x), denom, self.stream()), "scale")

    def step_end(self, n, time, nstep, tstep, ctl):
        self.ok(self.lib.tupan_cuda_step_end_dev(n, time.data_ptr(), nstep.data_ptr(), tstep.data_ptr(),
                                                 ctl.data_ptr(), self.stream()), "step_end")

    def reduce(self, what, n, arrays, out, param=0.0):
        self.ok(self.lib.tupan_cuda_reduce_dev(what, n, self.ptrs(arrays), param, out.data_ptr(), self.stream()),
                "reduce")


class Integrator(object):
    PROVIDED_METHODS = (["hermite%d" % o for o in (2, 4, 6, 8)] + ["ahermite%d" % o for o in (2, 4, 6, 8)]
                        + ["%s%s.%s" % (s, k, o) for s in sorted(SIA_COEFS) for k in "sah" for o in ("dkd", "kdk")]
                        + ["nreg", "anreg", "sakura", "asakura"])

    def __init__(self, eta, time, ps, method=None, device=None, group=None, shard=None, graph=None, pn_order=0,
                 clight=None, **kwargs):
        if method not in self.PROVIDED_METHODS:
            raise ValueError("Unexpected integration method: %r. Provided methods: %s"
                             % (method, self.PROVIDED_METHODS))
        # Base.__init__, integrator/__init__.py:24-37
        self.pn = None
        if pn_order and pn_order > 0:
            if clight is None:
                raise TypeError("'clight' is not defined. Please set the speed of light argument 'clight' "
                                "when using 'pn_order' > 0.")
            if not (method and method.startswith("sia") and method[5] in "sa"):
                raise NotImplementedError("post-Newtonian corrections: shared-step SIA methods only "
                                          "(the reference's Hermite/NREG/Sakura have none either)")
            inv1 = 1.0 / float(clight)
            self.pn = (int(pn_order),) + tuple(inv1 ** k for k in range(1, 8))     # extensions.py:31-60
        self.reporter = kwargs.pop("reporter", None)
        for k in ("viewer", "dumpper", "dump_freq", "gl_freq"):
            kwargs.pop(k, None)
        if kwargs:
            raise TypeError("Integrator.__init__ received unexpected keyword arguments: %s." % ", ".join(kwargs))
        self.eta, self.method, self.ps = float(eta), method, ps
        self.group = group
        # shard=None: i-shard over the ranks of `group` whenever torch.distributed is initialised
        if shard is None:
            shard = dist.is_initialized()
        self.world = dist.get_world_size(group) if shard else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        bounds = shard_bounds(ps.n, self.world)
        self.st = DeviceState(ps, self.device, bounds[self.rank], bounds[self.rank + 1])
        self.n_total = ps.n
        self.prec = "float64" if self.st.dtype == torch.float64 else "float32"
        self.L = _Lib(self.prec)
        self.ctl = torch.zeros(8, dtype=torch.float64, device=self.device)
        self.ctl[CTL_T_CURR] = float(time)
        self.ctl[CTL_ETA] = self.eta
        self._t_end = None
        self._scalar = torch.zeros(20, dtype=torch.float64, device=self.device)
        self._sharded = {}
        self.is_initialized = False
        # A Hermite / SIA step is a fixed sequence of launches whose only step-dependent input
        # (tau) lives in device memory: after two eager steps it is captured once into a CUDA
        # graph and replayed -- small-N steps are launch-bound otherwise (19 launches per
        # ahermite4 step).  Not for the sharded path (collectives on a second stream) nor for
        # sakura / nreg (they read a scalar back every step).
        capturable = self.world == 1 and ("hermite" in method or (method.startswith("sia") and method[5] != "h"))
        self.use_graph = capturable if graph is None else (bool(graph) and capturable)
        self._graph, self._graph_kernels, self._eager = None, 0, 0
        self.adaptive = method.startswith("a") or (method.startswith("sia") and method[5] == "a")
        if "hermite" in method:
            self.order = int(method[-1])
            self._step_name = "_hermite_step"
        elif method.startswith("sia"):
            self._step_name = "_sia_step"
            A, B = SIA_COEFS[method[:5]]
            kdk = method.endswith("kdk")
            self._kdk = kdk
            if method[5] == "h":
                if self.world > 1:
                    raise NotImplementedError("hierarchical SIA is single-GPU")
                self._step_name = "_sia_h_step"
                self.st.t["id"] = self.st._up(np.ascontiguousarray(ps.id).view(
                    np.int64 if self.st.np_dtype == np.float64 else np.int32))
            # bridge_sf with an empty fast set (sia.py:341-352): only the sf_drifts act, each
            # one evolve(slow, B_i * tau); evolve = the dkd / kdk composition (sia.py:308-337)
            self._bridge = [w for outer, w in operator_sequence(B, A) if outer]
            self._evolve = [(outer != kdk, w) for outer, w in operator_sequence(B, A)]   # (is_drift, weight)
        elif "sakura" in method:
            self._step_name = "_sakura_step"
        else:
            self._step_name = "_nreg_step"
            self._nreg = None

    # ---- forces --------------------------------------------------------------------------
    def force(self, kernel, out_names, scalars=(), inputs=None):
        """out_names: tensors (by state name) receiving KERNEL_OUTPUTS[kernel], in order."""
        st = self.st
        outs = st.need(*out_names)
        out = dict(zip(KERNEL_OUTPUTS[kernel], outs))
        src = {a: st[(inputs or {}).get(a, a)] for a in KERNEL_INPUTS[kernel]}
        if self.world == 1:
            run_kernel(kernel, src, src, scalars, out)
        else:
            sk = self._sharded.get(kernel)
            if sk is None:
                sk = self._sharded[kernel] = ShardedKernel(kernel, self.n_total, st.dtype, self.device,
                                                           group=self.group)
            sk.evaluate(src, scalars, out)

    def reduce(self, what, names, slot, param=0.0):
        """Reduction over ALL particles into self._scalar[slot] (device); all-reduced over ranks."""
        out = self._scalar[slot:slot + 1]
        self.L.reduce(what, self.st.n, [self.st[k] for k in names], out, param)
        if self.world > 1:
            if what in (RED_ABS_MIN,):
                op = dist.ReduceOp.MIN
            elif what in (RED_ABS_MAX,):
                op = dist.ReduceOp.MAX
            elif what == RED_SAKURA_DT:
                op = dist.ReduceOp.MIN            # dt = eta/sqrt(1 + max w2) is monotone in the max
            else:
                op = dist.ReduceOp.SUM
            dist.all_reduce(out, op=op, group=self.group)
        return out

    # ---- reference API ------------------------------------------------------------------------
    def initialize(self, t_end):
        if self.reporter:
            self.reporter.diagnostic_report(self.particle_system)
        self.is_initialized = True

    def finalize(self, t_end):
        torch.cuda.synchronize(self.device)
        hits = self.L.lib.tupan_cuda_kepler_limit_hits()
        if hits != 0:
            raise backend.TupanCudaError("%d pair(s) hit the Kepler sub-step bound during the run" % hits)

    def _set_t_end(self, t_end):
        if self._t_end != t_end:
            self._t_end = t_end
            self.ctl[CTL_T_END] = float(t_end)

    def evolve_step(self, t_end):
        """One step toward t_end, enqueued on the current stream (Base.evolve_step,
        integrator/__init__.py:80-99).  A step requested at or past t_end is a no-op."""
        if not self.is_initialized:
            self.initialize(t_end)
        self._set_t_end(t_end)
        if self._graph is not None:
            self._graph.replay()
            self.L.lib.tupan_cuda_count_launches(self._graph_kernels)
        else:
            self._step()
            self._eager += 1
            if self.use_graph and self._eager == 2:
                self._capture()
        if self.reporter:
            self.reporter.diagnostic_report(self.particle_system)

    def _step(self):
        # looked up by name: a bound method stored on self would be a reference cycle, and the
        # cyclic GC would then destroy this object's CUDA graph at an arbitrary moment
        getattr(self, self._step_name)()

    def _capture(self):
        lib = self.L.lib
        torch.cuda.synchronize(self.device)
        # destroying another CUDA graph (or freeing device memory) while this stream captures
        # invalidates the capture: collect garbage now, and keep the collector off meanwhile
        gc.collect()
        was_enabled = gc.isenabled()
        gc.disable()
        try:
            before = lib.tupan_cuda_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step()                   # recorded, not executed
            self._graph_kernels = lib.tupan_cuda_launch_count() - before
            lib.tupan_cuda_count_launches(-self._graph_kernels)
            self._graph = g
        finally:
            if was_enabled:
                gc.enable()

    def evolve(self, t_end, check_every=8, max_steps=None):
        """The driver loop `while abs(time) < t_end: evolve_step(t_end)` (simulation.py:187-201)
        with the clock read back only every `check_every` steps.  -> steps taken."""
        while True:
            for _ in range(check_every):
                self.evolve_step(t_end)
            c = self.ctl.cpu()
            self.L.check_async()               # a sync point anyway: surface asynchronous failures here
            if not (abs(float(c[CTL_T_CURR])) < abs(t_end)):
                break
            if max_steps is not None and int(c[CTL_NSTEPS]) >= max_steps:
                break
        return int(c[CTL_NSTEPS])

    @property
    def time(self):
        return float(self.ctl[CTL_T_CURR].item())

    @property
    def nsteps(self):
        return int(self.ctl[CTL_NSTEPS].item())

    @property
    def particle_system(self):
        self.st.download(self.ps)
        self.L.check_async()
        return self.ps

    def write_psdf(self, fname, fmode="a", with_potential=False):
        """Append a snapshot of the device-resident state to a PSDF stream (tupan/io/psdfio.py:25-35;
        tupan_b200/psdf.py): the columns go to the host with one copy each, straight from HBM."""
        from .psdf import PSDFWriter
        if self.world > 1:
            raise NotImplementedError("snapshots of a sharded state: use particle_system (gathers) and PSDFWriter")
        if with_potential:
            self.force("phi_kernel", ("phi",))
        cols = {k: v for k, v in self.st.t.items() if not (k.endswith("0") or k.startswith("_"))}
        if not with_potential:
            cols.pop("phi", None)
        return PSDFWriter(fname).dump(cols, fmode=fmode, t=self.time)

    def energies(self):
        """(kinetic, potential) as particles/body.py:262-306 defines them, reduced on the device."""
        self.force("phi_kernel", ("phi",))
        self.reduce(RED_KINETIC, ("mass",) + V3, 2)
        self.reduce(RED_HALF_DOT, ("mass", "phi"), 3)
        if self.pn:                                        # ke += pn_ke, body.py:283-287
            self.st.need("pn_ke")
            self.reduce(RED_SUM, ("pn_ke",), 4)
            ke, pe, pn_ke = self._scalar[2:5].cpu().tolist()
            return ke + pn_ke, pe
        ke, pe = self._scalar[2:4].cpu().tolist()
        return ke, pe

    def diagnostics(self):
        """The quantities of the reference's Diagnostic report (simulation.py:75-129): energies,
        virial, centre of mass, linear and angular momentum -- reduced on the device, one
        read-back of 17 doubles.  (Newtonian; the reference adds PN terms when enabled.)"""
        self.force("phi_kernel", ("phi",))
        self.reduce(RED_KINETIC, ("mass",) + V3, 0)
        self.reduce(RED_HALF_DOT, ("mass", "phi"), 1)
        self.reduce(RED_SUM, ("mass",), 2)
        for c in range(3):
            self.reduce(RED_DOT, ("mass", R3[c]), 3 + c)
            self.reduce(RED_DOT, ("mass", V3[c]), 6 + c)
            a, b = (c + 1) % 3, (c + 2) % 3
            self.reduce(RED_MOMENT, ("mass", R3[a], R3[b], V3[a], V3[b]), 9 + c)
        v = self._scalar[:12].cpu().tolist()
        ke, pe, mtot = v[0], v[1], v[2]
        return {"time": self.time, "ke": ke, "pe": pe, "te": ke + pe, "virial": 2 * ke + pe, "mtot": mtot,
                "com_r": [x / mtot for x in v[3:6]], "com_v": [x / mtot for x in v[6:9]],
                "lmom": v[6:9], "amom": v[9:12]}

    # ---- Hermite (integrator/hermite.py:390-410) ----------------------------------------------
    def _derivs(self, suffix):
        o = self.order
        kernel_out = (A3 if o == 2 else A3 + J3)
        names = tuple(k + suffix for k in kernel_out)
        self.force("acc_kernel" if o == 2 else "acc_jerk_kernel", names)
        if o >= 6:
            # snap_crackle consumes the a, j just computed for the same state (hermite.py:132-133)
            alias = {k: k + suffix for k in A3 + J3}
            self.force("snap_crackle_kernel", tuple(k + suffix for k in S3 + C3), inputs=alias)

    def _begin(self, min_names=None):
        if min_names is not None:
            dmin = self.reduce(RED_ABS_MIN, min_names, 0)
            self.L.step_begin(self.ctl, dmin)
        else:
            self.L.step_begin(self.ctl)

    # The raw per-particle criterion goes to a scratch array ("_tstep"): ps.tstep itself holds the
    # step actually taken (hermite.py:399), also after no-op steps past t_end.
    def _end(self):
        st = self.st
        self.L.step_end(st.n, st["time"], st["nstep"], st["tstep"], self.ctl)

    def _hermite_step(self):
        st, L, o = self.st, self.L, self.order
        if self.adaptive:                                  # get_hermite_tstep, hermite.py:343-349
            self.force("tstep_kernel", ("_tstep", "tstepij"), (self.eta,))
            self._begin(("_tstep",))
        else:
            self._begin()
        nd = DERIVS[:3 * (o // 2)]
        rv = st.need(*(R3 + V3))
        rv0 = st.need(*[k + "0" for k in R3 + V3])
        d0 = st.need(*[k + "0" for k in nd])
        d1 = st.need(*nd)
        self._derivs("0")                                  # epredict: forces of ps0
        L.predict(o, st.n, rv, rv0, d0, self.ctl)
        for _ in range(2):                                 # epec(2, ...), hermite.py:59-67
            self._derivs("")
            L.correct(o, st.n, rv, rv0, d0, d1, self.ctl)
        self._end()

    # ---- SIA, shared time-step (integrator/sia.py:1032-1123) ----------------------------------
    def _sia_step(self):
        st, L = self.st, self.L
        if self.n_total <= 2:
            raise NotImplementedError("n <= 2 goes to the Kepler solver (fewbody.py); use tupan_cuda_kepler_dev")
        if self.adaptive:
            self.force("tstep_kernel", ("_tstep", "tstepij"), (self.eta,))
            self._begin(("_tstep",))
        else:
            self._begin()
        r, v, a = st.need(*R3), st.need(*V3), st.need(*A3)
        if self.pn:
            PNA, PMR, PMV = ("pnax", "pnay", "pnaz"), ("pn_mrx", "pn_mry", "pn_mrz"), ("pn_mvx", "pn_mvy", "pn_mvz")
            pn_arr = st.need(*(V3 + A3 + ("wx", "wy", "wz") + PNA + ("mass",) + R3 + ("pn_ke",) + PMV
                               + ("pn_amx", "pn_amy", "pn_amz")))
            pmr, pmv = st.need(*PMR), st.need(*PMV)
        for wb in self._bridge:
            for is_drift, w in self._evolve:
                if is_drift:
                    L.axpy(st.n, r, v, wb, w, self.ctl)    # drift_n
                    if self.pn:                            # drift_pn: pn_drift_com_r
                        L.axpy(st.n, pmr, pmv, wb, w, self.ctl)
                elif self.pn:                              # kick = set_acc + kick_pn (sia.py:136-157)
                    self.force("acc_kernel", A3)
                    L.pn_kick(0, st.n, pn_arr, wb, w, self.ctl)
                    self.force("pnacc_kernel", PNA, self.pn)
                    L.pn_kick(1, st.n, pn_arr, wb, w, self.ctl)
                else:
                    self.force("acc_kernel", A3)           # kick = set_acc + kick_n
                    L.axpy(st.n, v, a, wb, w, self.ctl)
        self._end()

    # ---- SIA, hierarchical (individual block steps by recursive slow/fast splitting) ----------------
    # Mirrors SIA.recurse / split / join / sf_drift / sf_kick (sia.py:25-58, 192-294, 1089-1123).
    # A sub-system is a dict of device tensors; forces between sub-systems are rectangular
    # acc calls (ni != nj).  The recursion, like the reference's, is host logic: the size of
    # the slow and fast sets is data the host has to see.
    def _sub_force(self, kernel, ips, jps, out_names, scalars=()):
        for k in out_names:
            if k not in ips:
                ips[k] = torch.zeros_like(ips["mass"])
        src_i = {a: ips[a] for a in KERNEL_INPUTS[kernel]}
        src_j = {a: jps[a] for a in KERNEL_INPUTS[kernel]}
        run_kernel(kernel, src_i, src_j, scalars, dict(zip(KERNEL_OUTPUTS[kernel], [ips[k] for k in out_names])))

    @staticmethod
    def _sub_n(sub):
        return sub["mass"].numel()

    def _sub_axpy(self, sub, y, x, dt):
        self.L.axpy(self._sub_n(sub), [sub[k] for k in y], [sub[k] for k in x], dt, 1.0, None)

    def _sub_evolve(self, sub, tau):
        """SIAxy.dkd / .kdk on one sub-system (sia.py:308-337); n <= 2 -> FewBody.evolve."""
        n = self._sub_n(sub)
        if n == 0:
            return sub
        if n == 1:                                         # fewbody.py:18-27
            self._sub_axpy(sub, R3, V3, tau)
            return sub
        if n == 2:                                         # kepler_solver, in place (fewbody.py:34-44)
            ins = [sub[k] for k in ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz")]
            outs = [sub[k] for k in R3 + V3]
            self.L.ok(self.L.lib.tupan_cuda_kepler_dev(1, self.L.ptrs(ins), tau, self.L.ptrs(outs), self.L.stream()),
                      "kepler")
            return sub
        for is_drift, w in self._evolve:
            if is_drift:
                self._sub_axpy(sub, R3, V3, w * tau)
            else:
                self._sub_force("acc_kernel", sub, sub, A3)
                self._sub_axpy(sub, V3, A3, w * tau)
        return sub

    def _sub_recurse(self, sub, tau):
        n = self._sub_n(sub)
        if n == 0:
            return sub
        self._sub_force("tstep_kernel", sub, sub, ("tstep", "tstepij"), (self.eta,))
        if n <= 2:                                         # split(): stop the recursion
            slow, fast = sub, {k: v[:0] for k, v in sub.items()}
        else:
            cond = sub["tstep"].abs() > abs(tau)
            slow = {k: v[cond].contiguous() for k, v in sub.items()}
            ncond = ~cond
            fast = {k: v[ncond].contiguous() for k, v in sub.items()}
        A, B = SIA_COEFS[self.method[:5]]
        for outer, w in operator_sequence(B, A):           # bridge_sf, e.g. sia.py:341-352
            if outer:                                      # sf_drift
                slow = self._sub_evolve(slow, w * tau)
                fast = self._sub_recurse(fast, w * tau)
            elif self._sub_n(slow) and self._sub_n(fast):  # sf_kick
                self._sub_force("acc_kernel", slow, fast, A3)
                self._sub_force("acc_kernel", fast, slow, A3)
                self._sub_axpy(slow, V3, A3, w * tau)
                self._sub_axpy(fast, V3, A3, w * tau)
        if self._sub_n(fast) == 0:
            self._h_t += tau
        ns = self._sub_n(slow)
        if ns:
            self.L.ok(self.L.lib.tupan_cuda_stamp_dev(ns, slow["time"].data_ptr(), slow["nstep"].data_ptr(),
                                                      slow["tstep"].data_ptr(), tau, self.L.stream()), "stamp")
        if not self._sub_n(fast):                          # join()
            return slow
        if not ns:
            return fast
        for k in set(slow) | set(fast):
            if k not in slow:
                slow[k] = torch.zeros_like(slow["mass"])
            if k not in fast:
                fast[k] = torch.zeros_like(fast["mass"])
        return {k: torch.cat([slow[k], fast[k]]) for k in slow}

    def _sia_h_step(self):
        c = self.ctl.cpu()
        self._h_t = float(c[CTL_T_CURR])
        t_end = float(c[CTL_T_END])
        if not (abs(self._h_t) < abs(t_end)):
            return
        # Base.get_base_tstep, integrator/__init__.py:48-57
        dt = min(abs(t_end) - abs(self._h_t), abs(self.eta))
        dt = max(dt, abs(t_end) * (2 * np.finfo(np.float64).eps))
        tau = float(np.copysign(dt, self.eta))
        sub = {k: v for k, v in self.st.t.items() if not (k.endswith("0") or k.startswith("_"))}
        self.st.t = self._sub_recurse(sub, tau)
        self.ctl[CTL_T_CURR] = self._h_t
        self.ctl[CTL_NSTEPS] += 1

    # ---- Sakura (integrator/sakura.py:22-50, 100-147) -------------------------------------------
    def _tau_host(self):
        return float(self.ctl[CTL_TAU].item())

    def _sakura_step(self):
        st, L = self.st, self.L
        if self.adaptive:                                  # get_sakura_tstep
            self.force("tstep_kernel", ("_tstep", "tstepij"), (self.eta,))
            dmin = self.reduce(RED_SAKURA_DT, ("_tstep", "tstepij"), 0, self.eta)
            self.L.step_begin(self.ctl, dmin)
        else:
            self._begin()
        tau = self._tau_host()      # the pair kernel takes dt by value: one 8-byte read-back per step
        if tau == 0.0:
            return
        r, v = st.need(*R3), st.need(*V3)
        drv = st.need("drx", "dry", "drz", "dvx", "dvy", "dvz")
        L.axpy(st.n, r, v, 0.5, 1.0, self.ctl)
        for flag in (-1, 1):
            self.force("sakura_kernel", ("drx", "dry", "drz", "dvx", "dvy", "dvz"), (tau / 2, flag))
            L.axpy(st.n, r + v, drv)
        L.axpy(st.n, r, v, 0.5, 1.0, self.ctl)
        self._end()

    # ---- NREG (integrator/nreg.py) -------------------------------------------------------------------
    def _nreg_x(self, dt):
        st, L = self.st, self.L
        self.force("nreg_Xkernel", ("mrx", "mry", "mrz", "ax", "ay", "az", "u"), (dt,))
        L.scale(st.n, st.need(*R3), st.need("mrx", "mry", "mrz"), self._nreg["mtot"])
        u = self.reduce(RED_SUM, ("u",), 0)
        self._nreg["U"] = 0.5 * self._real(u.item())
        self._nreg["t"] = self._nreg["t"] + dt

    def _nreg_v(self, dt):
        st, L = self.st, self.L
        self.force("nreg_Vkernel", ("mvx", "mvy", "mvz", "mk"), (dt,))
        L.scale(st.n, st.need(*V3), st.need("mvx", "mvy", "mvz"), self._nreg["mtot"])
        mk = self.reduce(RED_SUM, ("mk",), 0)
        K = 0.25 * self._real(mk.item()) / self._nreg["mtot"]
        self._nreg["W"] = K - self._nreg["E0"]

    def _real(self, x):
        # numpy's REAL scalar arithmetic of the reference (u.sum() is a REAL scalar)
        return self.st.np_dtype.type(x)

    def _anreg_step(self, h):
        g = self._nreg
        self._nreg_x(0.5 * (h / g["W"]))
        self._nreg_v(h / g["U"])
        self._nreg_x(0.5 * (h / g["W"]))

    def _nreg_step(self):
        st = self.st
        if self._nreg is None:                             # NREG.initialize, nreg.py:110-131
            ke, pe = self.energies()
            self.reduce(RED_SUM, ("mass",), 0)
            self._nreg = {"E0": ke + pe, "W": -pe, "U": -pe, "S": -pe, "mtot": float(self._scalar[0].item()),
                          "t": self.time}
        g = self._nreg
        self._begin()
        c = self.ctl.cpu()
        if c[CTL_DONE] != 0:
            return
        tau = float(c[CTL_TAU])
        t0 = g["t"]
        if self.method == "anreg":
            self._anreg_step(tau / 2)
        else:                                              # nreg_step, nreg.py:87-94
            self._anreg_step(0.5 * (tau * g["S"]))
            g["S"] = 1 / (2 / g["W"] - 1 / g["S"])
            self._anreg_step(0.5 * (tau * g["S"]))
        dt = g["t"] - t0
        # ps.tstep[...] = dt; ps.time += tau; ps.nstep += 1; the clock advanced by dt (nreg.py:159-166)
        st["tstep"].fill_(float(dt))
        st["time"].add_(tau)
        st["nstep"].add_(1)
        self.ctl[CTL_T_CURR] = float(g["t"])
        self.ctl[CTL_NSTEPS] += 1
