"""Reader for the committed golden vectors (tests/golden/*.npz, made by make_golden.py)."""
import os
import re

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAGS = {"float64": "fp64", "float32": "fp32"}


def load_set(name, prec):
    """-> (inputs dict, {(ni, nj): {case: [outputs]}}, (eta, dt, clight))"""
    z = np.load(os.path.join(GOLDEN, "%s_%s.npz" % (name, TAGS[prec])))
    inputs, shapes = {}, {}
    for key in z.files:
        if key.startswith("in/"):
            inputs[key[3:]] = z[key]
        elif key != "scalars":
            shape, case, q = key.split("/")
            m = re.match(r"i(\d+)_j(\d+)", shape)
            ninj = (int(m.group(1)), int(m.group(2)))
            shapes.setdefault(ninj, {}).setdefault(case, {})[int(q)] = z[key]
    for ninj in shapes:
        for case in shapes[ninj]:
            d = shapes[ninj][case]
            shapes[ninj][case] = [d[q] for q in range(len(d))]
    return inputs, shapes, tuple(z["scalars"])


def split_case(case):
    """'pnacc_kernel_o7' -> ('pnacc_kernel', ('o', 7)); 'sakura_kernel_f-1' -> (.., ('f', -1))"""
    m = re.match(r"(.*)_([of])(-?\d+)$", case)
    if m:
        return m.group(1), (m.group(2), int(m.group(3)))
    return case, None


def load_kepler(prec):
    z = np.load(os.path.join(GOLDEN, "kepler_%s.npz" % TAGS[prec]))
    n = 1 + max(int(k.split("/")[0]) for k in z.files)
    cases = []
    for c in range(n):
        ins = {k.split("/")[2]: z[k] for k in z.files if k.startswith("%d/in/" % c)}
        outs = {k.split("/")[2]: z[k] for k in z.files if k.startswith("%d/out/" % c)}
        cases.append((ins, float(z["%d/dt" % c]), outs))
    return cases


def load_integrator_cases(prec, name="integrators"):
    """-> {case: (inputs dict, outputs dict, meta)}; meta = eta, t_end, steps, t_final, ke0, pe0,
    ke1, pe1 (tests/golden/make_golden_integrators.py)."""
    z = np.load(os.path.join(GOLDEN, "%s_%s.npz" % (name, TAGS[prec])))
    cases = {}
    for key in z.files:
        case, rest = key.split("/", 1)
        ins, outs, meta = cases.setdefault(case, ({}, {}, [None]))
        if rest.startswith("in/"):
            ins[rest[3:]] = z[key]
        elif rest.startswith("out/"):
            outs[rest[4:]] = z[key]
        else:
            meta[0] = z[key]
    return {k: (i, o, m[0]) for k, (i, o, m) in cases.items()}
