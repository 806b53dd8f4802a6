"""``__graft_entry__.smoke()`` and the bench step's call against the emulated C ABI (see run_gpu_tests_on_cpu.py)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_gpu_tests_on_cpu as R  # noqa: E402

sys.path.insert(0, R.ROOT)


def main():
    R.E.install()
    R.nv.call = R.extra_call
    R.nv.launch_count = lambda: R.LAUNCHES[0]
    R.nv.reset_launch_count = lambda: R.LAUNCHES.__setitem__(0, 0)
    torch.cuda.synchronize = lambda *a, **k: None
    import __graft_entry__ as entry
    from HyGrid import functional as Fn
    with R.CpuDevices():
        entry.smoke()
        # bench.py's step: caller-owned output buffer, float32 fast math
        x = torch.rand(4, 3, 32, 48) * 255
        y = torch.empty(4, 3, 32, 48)
        assert Fn.rect_to_hex(x, (32, 48), "bilinear", out_dtype=torch.float32, math="fast", out=y) is y
        ref = np.stack([R.O.rect_to_hex_resample(x[i].numpy(), (32, 48), "bilinear") for i in range(4)])
        assert float(np.abs(y.numpy() - ref).max()) <= 1e-5 * 255 * 4
    print("ok smoke and bench step")


if __name__ == "__main__":
    main()
