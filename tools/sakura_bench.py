"""Device-resident throughput of sakura_kernel (GPU box): pairs/s on Plummer spheres and on the
binary-rich Plummer sphere of BASELINE.json configs[4], with the fraction of pairs that take
the Kepler branch (sakura_kernel_common.h:94-123: r2 <= (64 m / v2)^2) counted on the host.

    python tools/sakura_bench.py [float64,float32]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tupan_b200 import device, ics  # noqa: E402


def kepler_fraction(d, sample=512):
    """Fraction of (i, j) pairs with r2 <= (64 m / v2)^2, on a sample of i rows (flag = 1)."""
    n = d["mass"].numel()
    idx = torch.linspace(0, n - 1, min(sample, n), device="cuda").long()
    r2 = sum((d[k][idx, None] - d[k][None, :]) ** 2 for k in ("rx", "ry", "rz"))
    v2 = sum((d[k][idx, None] - d[k][None, :]) ** 2 for k in ("vx", "vy", "vz"))
    m = d["mass"][idx, None] + d["mass"][None, :]
    R = 64 * m / v2
    kep = (r2 <= R * R) & (r2 > 0)
    return kep.double().mean().item()


def main():
    precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["float64"]
    cases = [("plummer", 1024), ("plummer", 4096), ("plummer", 16384), ("binary-rich", 16384),
             ("plummer", 65536)]
    for prec in precs:
        for kind, n in cases:
            ps = ics.make_plummer(n, seed=1, dtype=prec) if kind == "plummer" else \
                ics.make_binary_rich(n, seed=1, dtype=prec)
            d = device.to_device(ps)
            frac = kepler_fraction(d)
            for dt in (1.0 / 64, 1.0 / 1024):
                for flag in (1, 2):
                    out = device.run("sakura_kernel", d, d, (dt, flag))
                    torch.cuda.synchronize()
                    reps = 3
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(reps):
                        device.run("sakura_kernel", d, d, (dt, flag), out)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / reps
                    chk = float(sum(out[k].double().abs().sum() for k in out))
                    print("%-8s sakura %-11s N=%-6d dt=%-10.3g flag=%d kepler-branch %.3f%%  %9.3f ms  %8.2f Gpair/s  chk %.12e"
                          % (prec, kind, n, dt, flag, 100 * frac, ms, float(n) * n / ms * 1e-6, chk), flush=True)


if __name__ == "__main__":
    main()
