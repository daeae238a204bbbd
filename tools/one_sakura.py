"""One sakura evaluation for profiling (ncu): binary-rich Plummer N = 16384 (BASELINE configs[4]),
dt = 1/1024, flag = 1, device-resident."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tupan_b200 import device, ics  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = device.to_device(ics.make_binary_rich(n, seed=1))
out = device.run("sakura_kernel", d, d, (1.0 / 1024, 1))
torch.cuda.synchronize()
device.run("sakura_kernel", d, d, (1.0 / 1024, 1), out)
torch.cuda.synchronize()
print("ok", float(out["drx"].abs().sum()))
