"""CPU restatement of the reference's integrators -- TEST INFRASTRUCTURE ONLY.

numpy restatement of ``tupan/integrator`` (Hermite 2/4/6/8, SIA 21...69 shared / adaptive /
hierarchical, NREG, Sakura) driven by the CPU oracle kernels (``oracle/tupan_oracle.c`` or
``oracle/_ref``).  Every function cites the reference lines it follows and keeps their
operation order, so that with the bit-identical oracle kernels underneath the final states
reproduce the reference's bit for bit.

Parity status: pinned -- ``tests/test_oracle_integrators.py`` checks every case of
``tests/golden/integrators_fp{64,32}.npz`` (produced by the reference's own
``tupan.integrator.Integrator`` with ``tests/golden/make_golden_integrators.py``).

Only ``tests/`` may import this module; the product integrators are
``tupan_b200/integrator.py`` + ``tupan_b200/csrc/k_update.cu`` and have no CPU path.
"""
import math
import sys

import numpy as np

from . import call, call_threaded, load

S8 = ("mass", "rx", "ry", "rz", "eps2", "vx", "vy", "vz")
S5 = ("mass", "rx", "ry", "rz", "eps2")
S14 = S8 + ("ax", "ay", "az", "jx", "jy", "jz")
SV = ("mass", "vx", "vy", "vz", "ax", "ay", "az")


class Bodies(object):
    """SoA arrays + the class-level clock of the reference's ParticleSystem
    (particles/body.py:26-39; `type(ps).t_curr`, integrator/__init__.py:20)."""

    def __init__(self, arrays, prec, kind="oracle", threads=1):
        self.threads = threads                    # > 1: contiguous i-slices in a thread pool
        self.prec = np.dtype(prec).name
        self.dtype = np.dtype(prec)
        self.lib = load(kind, self.prec)
        self.kind = kind
        self.a = {k: np.ascontiguousarray(v).copy() for k, v in arrays.items()}
        n = len(self.a["mass"])
        utype = np.uint64 if self.dtype == np.float64 else np.uint32
        self.a.setdefault("id", np.arange(n).astype(utype))
        self.a.setdefault("time", np.zeros(n, self.dtype))
        self.a.setdefault("tstep", np.zeros(n, self.dtype))
        self.a.setdefault("nstep", np.zeros(n, utype))
        self.clock = [0.0]          # shared by every slice, like the class attribute t_curr
        self.pn = None              # (order, clight) when post-Newtonian corrections are on

    @property
    def n(self):
        return len(self.a["mass"])

    def __getattr__(self, k):
        try:
            return self.__dict__["a"][k]
        except KeyError:
            raise AttributeError(k)

    def need(self, *names):
        for k in names:
            if k not in self.a:
                self.a[k] = np.zeros(self.n, self.dtype)

    def copy(self):
        o = Bodies.__new__(Bodies)
        o.__dict__.update(self.__dict__)
        o.a = {k: v.copy() for k, v in self.a.items()}
        return o

    def select(self, mask):                       # ps[condition] (fancy index = copy)
        o = Bodies.__new__(Bodies)
        o.__dict__.update(self.__dict__)
        o.a = {k: np.ascontiguousarray(v[mask]) for k, v in self.a.items()}
        return o

    def append(self, other):                      # join(): slow.append(fast), sia.py:46-58
        for k in set(self.a) | set(other.a):
            if k not in self.a:
                self.a[k] = np.zeros(self.n, other.a[k].dtype)
            if k not in other.a:
                other.a[k] = np.zeros(other.n, self.a[k].dtype)
        self.a = {k: np.concatenate([self.a[k], other.a[k]]) for k in self.a}

    # ---- force setters (particles/body.py:324-361 -> lib/extensions.py) -------------------
    def _call(self, name, attrs, jps, scalars, outs):
        self.need(*outs)
        jps.need(*[a for a in attrs if a not in jps.a])
        args = ([self.n] + [self.a[k] for k in attrs] + [jps.n] + [jps.a[k] for k in attrs]
                + list(scalars) + [self.a[k] for k in outs])
        if self.threads > 1:
            call_threaded(self.lib, name, self.prec, self.threads, *args)
        else:
            call(self.lib, name, self.prec, *args)

    def set_phi(self, jps):
        self._call("phi_kernel", S5, jps, (), ("phi",))

    def set_acc(self, jps):
        self._call("acc_kernel", S5, jps, (), ("ax", "ay", "az"))

    def set_acc_jerk(self, jps):
        self._call("acc_jerk_kernel", S8, jps, (), ("ax", "ay", "az", "jx", "jy", "jz"))

    def set_snap_crackle(self, jps):
        self._call("snap_crackle_kernel", S14, jps, (), ("sx", "sy", "sz", "cx", "cy", "cz"))

    def set_pnacc(self, jps):                     # extensions.py:395-446, Clight :31-60
        order, c = self.pn
        inv = [(1.0 / c) ** k for k in range(1, 8)]
        self._call("pnacc_kernel", S8, jps, [order] + inv, ("pnax", "pnay", "pnaz"))

    # PNbodyMethods, particles/body.py:471-527
    def pn_kick_ke(self, tau):
        self.need("pn_ke")
        pnfx, pnfy, pnfz = self.mass * self.pnax, self.mass * self.pnay, self.mass * self.pnaz
        self.a["pn_ke"] -= (self.vx * pnfx + self.vy * pnfy + self.vz * pnfz) * tau

    def pn_drift_com_r(self, tau):
        self.need("pn_mrx", "pn_mry", "pn_mrz", "pn_mvx", "pn_mvy", "pn_mvz")
        for c in "xyz":
            self.a["pn_mr" + c] += self.a["pn_mv" + c] * tau

    def pn_kick_lmom(self, tau):
        self.need("pn_mvx", "pn_mvy", "pn_mvz")
        for c in "xyz":
            self.a["pn_mv" + c] -= (self.mass * self.a["pna" + c]) * tau

    def pn_kick_amom(self, tau):
        self.need("pn_amx", "pn_amy", "pn_amz")
        pnfx, pnfy, pnfz = self.mass * self.pnax, self.mass * self.pnay, self.mass * self.pnaz
        self.a["pn_amx"] -= (self.ry * pnfz - self.rz * pnfy) * tau
        self.a["pn_amy"] -= (self.rz * pnfx - self.rx * pnfz) * tau
        self.a["pn_amz"] -= (self.rx * pnfy - self.ry * pnfx) * tau

    def set_tstep(self, jps, eta):
        self._call("tstep_kernel", S8, jps, (eta,), ("tstep", "tstepij"))

    def sakura(self, dt, flag):
        self._call("sakura_kernel", S8, self, (dt, flag), ("drx", "dry", "drz", "dvx", "dvy", "dvz"))

    def nreg_x(self, dt):
        self._call("nreg_Xkernel", S8, self, (dt,), ("mrx", "mry", "mrz", "ax", "ay", "az", "u"))

    def nreg_v(self, dt):
        self._call("nreg_Vkernel", SV, self, (dt,), ("mvx", "mvy", "mvz", "mk"))

    def kepler(self, dt):                         # fewbody.py:34-44, in place (extensions.py:642-646)
        args = [self.a[k] for k in S8] + [dt] + [self.a[k] for k in ("rx", "ry", "rz", "vx", "vy", "vz")]
        call(self.lib, "kepler_solver_kernel", self.prec, *args)

    # ---- diagnostics (particles/body.py:262-306, 364-368) -------------------------------
    @property
    def kinetic_energy(self):                     # body.py:262-289 (+ pn_ke when PN is on)
        ke = 0.5 * self.mass * (self.vx ** 2 + self.vy ** 2 + self.vz ** 2)
        if self.pn:
            self.need("pn_ke")
            ke += self.a["pn_ke"]
        return float(ke.sum())

    @property
    def potential_energy(self):
        self.set_phi(self)
        return 0.5 * float((self.mass * self.phi).sum())

    @property
    def total_mass(self):
        return float(self.mass.sum())

    def min_tstep(self):
        return abs(self.tstep).min()


# ---- Base (integrator/__init__.py:48-78) ---------------------------------------------------
def get_base_tstep(t_curr, t_end, eta):
    dt = min(abs(t_end) - abs(t_curr), abs(eta))
    dt = max(dt, abs(t_end) * (2 * sys.float_info.epsilon))
    return math.copysign(dt, eta)


def get_min_block_tstep(min_ts, t_curr, tau):
    power = int(np.log2(min_ts) - 1)
    min_bts = 2.0 ** power
    t_next = t_curr + min_bts
    while t_next % min_bts != 0:
        min_bts /= 2
    min_bts = math.copysign(min_bts, tau)
    if abs(min_bts) > abs(tau):
        min_bts = tau
    return min_bts


# ---- Hermite (integrator/hermite.py) --------------------------------------------------------
def _forces(ps, order):
    if order == 2:
        ps.set_acc(ps)
        return
    ps.set_acc_jerk(ps)
    if order >= 6:
        ps.set_snap_crackle(ps)


def hermite_predict(ps, tau, order):
    """H2/H4/H6/H8.epredict, hermite.py:25-43, 75-91, 127-158, 202-242."""
    ps0 = ps.copy()
    _forces(ps0, order)
    ps1 = ps
    for c in "xyz":
        a = ps0.a
        v, ac = a["v" + c], a["a" + c]
        if order == 2:
            ps1.a["r" + c] += (ac * tau / 2 + v) * tau
            ps1.a["v" + c] += ac * tau
        elif order == 4:
            j = a["j" + c]
            ps1.a["r" + c] += ((j * tau / 3 + ac) * tau / 2 + v) * tau
            ps1.a["v" + c] += (j * tau / 2 + ac) * tau
        elif order == 6:
            j, s = a["j" + c], a["s" + c]
            ps1.a["r" + c] += (((s * tau / 4 + j) * tau / 3 + ac) * tau / 2 + v) * tau
            ps1.a["v" + c] += ((s * tau / 3 + j) * tau / 2 + ac) * tau
        else:
            j, s, k = a["j" + c], a["s" + c], a["c" + c]
            ps1.a["r" + c] += ((((k * tau / 5 + s) * tau / 4 + j) * tau / 3 + ac) * tau / 2 + v) * tau
            ps1.a["v" + c] += (((k * tau / 4 + s) * tau / 3 + j) * tau / 2 + ac) * tau
    return ps1, ps0


def hermite_correct(ps1, ps0, tau, order):
    """H2/H4/H6/H8.ecorrect, hermite.py:45-57, 93-121, 160-196, 244-283 (v first, then r with
    the NEW v)."""
    _forces(ps1, order)
    for c in "xyz":
        p0, p1 = ps0.a, ps1.a
        r0, v0, a0, a1 = p0["r" + c], p0["v" + c], p0["a" + c], p1["a" + c]
        if order == 2:
            p1["v" + c][...] = ((a0 + a1) * tau / 2 + v0)
            v1 = p1["v" + c]
            p1["r" + c][...] = ((v0 + v1) * tau / 2 + r0)
        elif order == 4:
            j0, j1 = p0["j" + c], p1["j" + c]
            p1["v" + c][...] = (((j0 - j1) * tau / 6 + (a0 + a1)) * tau / 2 + v0)
            v1 = p1["v" + c]
            p1["r" + c][...] = (((a0 - a1) * tau / 6 + (v0 + v1)) * tau / 2 + r0)
        elif order == 6:
            j0, j1, s0, s1 = p0["j" + c], p1["j" + c], p0["s" + c], p1["s" + c]
            p1["v" + c][...] = ((((s0 + s1) * tau / 12 + (j0 - j1)) * tau / 5 + (a0 + a1)) * tau / 2 + v0)
            v1 = p1["v" + c]
            p1["r" + c][...] = ((((j0 + j1) * tau / 12 + (a0 - a1)) * tau / 5 + (v0 + v1)) * tau / 2 + r0)
        else:
            j0, j1, s0, s1 = p0["j" + c], p1["j" + c], p0["s" + c], p1["s" + c]
            c0, c1 = p0["c" + c], p1["c" + c]
            p1["v" + c][...] = (((((c0 - c1) * tau / 20 + (s0 + s1)) * tau / 3 + 3 * (j0 - j1)) * tau / 14
                                 + (a0 + a1)) * tau / 2 + v0)
            v1 = p1["v" + c]
            p1["r" + c][...] = (((((s0 - s1) * tau / 20 + (j0 + j1)) * tau / 3 + 3 * (a0 - a1)) * tau / 14
                                 + (v0 + v1)) * tau / 2 + r0)
    return ps1


def hermite_step(ps, method, eta, tau):
    """Hermite.do_step, hermite.py:390-410."""
    order = int(method[-1])
    if "ahermite" in method:
        ps.set_tstep(ps, eta)
        tau = get_min_block_tstep(ps.min_tstep(), ps.clock[0], tau)
    ps1, ps0 = hermite_predict(ps, tau, order)
    for _ in range(2):
        ps1 = hermite_correct(ps1, ps0, tau, order)
    ps1.clock[0] += tau
    ps1.tstep[...] = tau
    ps1.a["time"] += tau
    ps1.a["nstep"] += 1
    return ps1


# ---- SIA (integrator/sia.py) ---------------------------------------------------------------------
SIA_COEFS = {     # sia.py:302-303, 360-362, 425-428, 497-501, 576-581, 662-668, 755-762, 855-864
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


def palindrome(first, second):
    """Operator sequence of every SIAxy.dkd/kdk/bridge_sf (e.g. sia.py:441-453): the two
    coefficient lists interleaved, then mirrored about the last element.
    -> [(which_list, coefficient)], which_list 0 = `first` (outer), 1 = `second`."""
    seq = []
    for i in range(max(len(first), len(second))):
        if i < len(first):
            seq.append((0, first[i]))
        if i < len(second):
            seq.append((1, second[i]))
    return seq + seq[-2::-1]


def sia_drift(ips, tau):                         # drift / drift_n / drift_pn, sia.py:64-71, 90-98, 165-173
    ips.a["rx"] += ips.vx * tau
    ips.a["ry"] += ips.vy * tau
    ips.a["rz"] += ips.vz * tau
    if ips.pn:
        ips.pn_drift_com_r(tau)
    return ips


def sia_kick(ips, tau):                          # kick + kick_n / kick_pn, sia.py:77-84, 104-159, 179-186
    ips.set_acc(ips)
    if ips.pn:                                   # the active branch of kick_pn, sia.py:136-157
        ips.need("wx", "wy", "wz")
        for c in "xyz":
            ips.a["v" + c] += (ips.a["a" + c] * tau + ips.a["w" + c]) / 2
        ips.set_pnacc(ips)
        ips.pn_kick_ke(tau)
        ips.pn_kick_lmom(tau)
        ips.pn_kick_amom(tau)
        for c in "xyz":
            ips.a["w" + c][...] = 2 * ips.a["pna" + c] * tau - ips.a["w" + c]
        for c in "xyz":
            ips.a["v" + c] += (ips.a["a" + c] * tau + ips.a["w" + c]) / 2
        return ips
    ips.a["vx"] += ips.ax * tau
    ips.a["vy"] += ips.ay * tau
    ips.a["vz"] += ips.az * tau
    return ips


def fewbody_evolve(ips, tau):                    # fewbody.py:46-57
    if ips.n == 0:
        return ips
    if ips.n == 1:
        return sia_drift(ips, tau)
    ips.kepler(tau)
    return ips


def sia_evolve(ips, tau, name, kdk):
    """SIAxy.dkd / .kdk (e.g. sia.py:308-337)."""
    if ips.n <= 2:
        return fewbody_evolve(ips, tau)
    A, B = SIA_COEFS[name]
    for which, c in palindrome(B, A):            # B outermost; dkd: B drifts, A kicks; kdk: swapped
        is_drift = (which == 0) != kdk
        ips = sia_drift(ips, c * tau) if is_drift else sia_kick(ips, c * tau)
    return ips


def sf_kick(slow, fast, tau):                    # sia.py:206-294 (Newtonian branch)
    if slow.n and fast.n:
        slow.set_acc(fast)
        fast.set_acc(slow)
        for p in (slow, fast):
            p.a["vx"] += p.ax * tau
            p.a["vy"] += p.ay * tau
            p.a["vz"] += p.az * tau
    return slow, fast


class SIA(object):
    def __init__(self, method, eta):
        self.method, self.eta = method, eta
        self.name = method[:5]
        self.kdk = "kdk" in method
        kind = method[5]
        self.update_tstep = kind in "ah"         # sia.py:1036-1046
        self.shared_tstep = kind in "sa"

    def bridge(self, slow, fast, tau):           # SIAxy.bridge_sf, e.g. sia.py:341-352
        A, B = SIA_COEFS[self.name]
        for which, c in palindrome(B, A):
            if which == 0:                       # sf_drift, sia.py:192-200
                slow = sia_evolve(slow, c * tau, self.name, self.kdk)
                fast = self.recurse(fast, c * tau)
            else:
                slow, fast = sf_kick(slow, fast, c * tau)
        return slow, fast

    def recurse(self, ps, tau):                  # sia.py:1089-1123
        if not ps.n:
            return ps
        flag = -1
        if self.update_tstep:
            flag = 1
            ps.set_tstep(ps, self.eta)
            if self.shared_tstep:
                tau = get_min_block_tstep(ps.min_tstep(), ps.clock[0], tau)
        cond = abs(ps.tstep) > flag * abs(tau)
        if ps.n <= 2:                            # split(), sia.py:25-40
            slow, fast = ps, ps.select(np.zeros(ps.n, bool))
        else:
            slow, fast = ps.select(cond), ps.select(~cond)
        slow, fast = self.bridge(slow, fast, tau)
        if fast.n == 0:
            ps.clock[0] += tau
        if slow.n:
            slow.tstep[...] = tau
            slow.a["time"] += tau
            slow.a["nstep"] += 1
        if not fast.n:                           # join(), sia.py:46-58
            return slow
        if not slow.n:
            return fast
        slow.append(fast)
        return slow


# ---- NREG (integrator/nreg.py) -------------------------------------------------------------------
class NREG(object):
    def __init__(self, ps, method):
        self.method = method                     # nreg.py:118-124
        self.E0 = ps.kinetic_energy + ps.potential_energy
        self.W = self.U = self.S = -ps.potential_energy

    def x(self, ps, dt):                         # nreg_x, nreg.py:22-33
        mtot = ps.total_mass
        ps.nreg_x(dt)
        ps.rx[...] = ps.mrx / mtot
        ps.ry[...] = ps.mry / mtot
        ps.rz[...] = ps.mrz / mtot
        self.U = 0.5 * ps.u.sum()
        ps.clock[0] += dt
        return ps

    def v(self, ps, dt):                         # nreg_v, nreg.py:45-57
        mtot = ps.total_mass
        ps.nreg_v(dt)
        ps.vx[...] = ps.mvx / mtot
        ps.vy[...] = ps.mvy / mtot
        ps.vz[...] = ps.mvz / mtot
        K = 0.25 * ps.mk.sum() / mtot
        self.W = K - self.E0
        return ps

    def anreg_step(self, ps, h):                 # nreg.py:78-84
        ps = self.x(ps, 0.5 * (h / self.W))
        ps = self.v(ps, (h / self.U))
        ps = self.x(ps, 0.5 * (h / self.W))
        return ps

    def step(self, ps, tau):                     # NREG.do_step, nreg.py:151-175
        t0 = ps.clock[0]
        if "anreg" in self.method:
            ps = self.anreg_step(ps, tau / 2)
        else:                                    # nreg_step, nreg.py:87-94
            ps = self.anreg_step(ps, 0.5 * (tau * self.S))
            self.S = 1 / (2 / self.W - 1 / self.S)
            ps = self.anreg_step(ps, 0.5 * (tau * self.S))
        dt = ps.clock[0] - t0
        ps.tstep[...] = dt
        ps.a["time"] += tau
        ps.a["nstep"] += 1
        return ps


# ---- Sakura (integrator/sakura.py) ---------------------------------------------------------------
def sakura_step(ps, tau):                        # sakura.py:22-50
    for c in "xyz":
        ps.a["r" + c] += ps.a["v" + c] * tau / 2
    for flag in (-1, 1):
        ps.sakura(tau / 2, flag)
        for c in "xyz":
            ps.a["r" + c] += ps.a["dr" + c]
        for c in "xyz":
            ps.a["v" + c] += ps.a["dv" + c]
    for c in "xyz":
        ps.a["r" + c] += ps.a["v" + c] * tau / 2
    return ps


def sakura_do_step(ps, method, eta, tau):        # Sakura.do_step + get_sakura_tstep, sakura.py:100-147
    if "asakura" in method:
        ps.set_tstep(ps, eta)
        iw2_a = (eta / ps.tstep) ** 2
        iw2_b = (eta / ps.tstepij) ** 2
        w2_sakura = (iw2_a - iw2_b).max()
        dt_sakura = eta / (1 + w2_sakura) ** 0.5
        ps.tstep[...] = dt_sakura
        tau = get_min_block_tstep(ps.min_tstep(), ps.clock[0], tau)
    ps = sakura_step(ps, tau)
    ps.clock[0] += tau
    ps.tstep[...] = tau
    ps.a["time"] += tau
    ps.a["nstep"] += 1
    return ps


# ---- driver (simulation.py:187-201 + Base.evolve_step, integrator/__init__.py:80-99) ---------------
def evolve(arrays, prec, method, eta, t_end, kind="oracle", t0=0.0, max_steps=None, threads=1, pn=None):
    """-> (Bodies after the run, number of steps).  pn = (order, clight) turns the post-Newtonian
    corrections on (Base.__init__, integrator/__init__.py:24-37; SIA only, as in the reference)."""
    ps = Bodies(arrays, prec, kind, threads)
    ps.pn = pn
    ps.clock[0] = t0
    sia = SIA(method, eta) if method.startswith("sia") else None
    nreg = None
    steps = 0
    while abs(ps.clock[0]) < t_end:
        if max_steps is not None and steps >= max_steps:
            break
        tau = get_base_tstep(ps.clock[0], t_end, eta)
        if "hermite" in method:
            ps = hermite_step(ps, method, eta, tau)
        elif sia is not None:
            ps = sia.recurse(ps, tau)
        elif "nreg" in method:
            if nreg is None:                     # NREG.initialize, nreg.py:110-131
                nreg = NREG(ps, method)
            ps = nreg.step(ps, tau)
        elif "sakura" in method:
            ps = sakura_do_step(ps, method, eta, tau)
        else:
            raise ValueError(method)
        steps += 1
    return ps, steps
