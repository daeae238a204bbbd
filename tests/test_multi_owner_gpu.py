"""The multi-owner sweep (tupan_cuda_sweep_multi_dev: ONE launch over packed rows that live in
several buffers, the kernel of the peer-memory transport) on a single GPU: the rows of one system
are cut into ragged pieces held in separate allocations; the result must equal the ordinary
evaluation.  Covers every launch shape through forced plans."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = (("acc_jerk_kernel", ()), ("tstep_kernel", (1.0 / 64,)), ("phi_kernel", ()), ("snap_crackle_kernel", ()),
         ("sakura_kernel", (1.0 / 256, 1)), ("pnacc_kernel", (7,) + tuple(128.0 ** -k for k in range(1, 8))))


@pytest.mark.parametrize("plan", ((-1, 0, 1), (0, 0, 1), (0, 0, 5), (2, 0, 5), (1, 0, 3), (1, 3, 2), (1, 5, 1)))
def test_multi_owner_sweep_equals_single_buffer(plan):
    import torch
    from tupan_b200 import backend, device, ics, sharded
    n = 3000
    ps = ics.make_plummer(n, seed=11)
    full = device.to_device(ps)
    g = torch.Generator(device="cpu").manual_seed(3)
    for k in ("ax", "ay", "az", "jx", "jy", "jz"):
        full[k] = torch.randn(n, dtype=torch.float64, generator=g).cuda()
    eng = sharded.CudaEngine("float64")
    lib = backend.require_gpu("float64")
    cuts = [0, 1000, 1037, 1037, 2900, n]           # ragged owners, one of them empty
    ni = 777                                       # rectangular: the first 777 particles as i-set
    try:
        for kernel, scal in CASES:
            ins = device.KERNEL_INPUTS[kernel]
            jt = [full[a] for a in ins]
            it = [full[a][:ni].contiguous() for a in ins]
            lib.tupan_cuda_force_plan(-1, 0, 1)
            ref = device.run(kernel, dict(zip(ins, it)), full, scal)
            width = eng.row_width(kernel, scal)
            packed = torch.zeros(n * width, dtype=torch.float64, device="cuda")
            eng.pack(kernel, jt, scal, packed)
            pieces = [packed[a * width:b * width].clone() for a, b in zip(cuts[:-1], cuts[1:])]
            rows = [b - a for a, b in zip(cuts[:-1], cuts[1:])]
            ptrs = [p.data_ptr() if p.numel() else 0 for p in pieces]
            lib.tupan_cuda_force_plan(*plan)
            nslots = eng.sweep_multi_slots(kernel, ni, rows, scal)
            assert nslots >= 1
            na = eng.n_acc(kernel, scal)
            partial = torch.empty(nslots * na * ni, dtype=torch.float64, device="cuda")
            eng.sweep_multi(kernel, it, ptrs, rows, scal, partial, 0)
            out = {a: torch.empty(ni, dtype=torch.float64, device="cuda") for a in device.KERNEL_OUTPUTS[kernel]}
            eng.finalize(kernel, it, partial, nslots, scal, [out[a] for a in device.KERNEL_OUTPUTS[kernel]])
            torch.cuda.synchronize()
            for name in device.KERNEL_OUTPUTS[kernel]:
                a, b = out[name].cpu().numpy(), ref[name].cpu().numpy()
                err = np.max(np.abs(a - b)) / np.max(np.abs(b))
                assert err < (1e-10 if kernel == "sakura_kernel" else 1e-12), (kernel, plan, name, err)
    finally:
        lib.tupan_cuda_force_plan(-1, 0, 1)
