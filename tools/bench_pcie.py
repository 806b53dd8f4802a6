#!/usr/bin/env python
"""Ceiling of the end-to-end leg: concurrent H2D + D2H of pinned buffers (the C2 step moves 3.2 GB each way; the float64
drop-in leg 3.2 GB in and 6.4 GB out).  Under torchrun every rank copies at the same time through its own GPU's link --
the N-rank floor of the `e2e` numbers in bench.py (all ranks share the host's memory system and PCIe root complexes).

    python tools/bench_pcie.py                                                    # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_pcie.py    # N GPUs of one box, concurrently
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 * 3 * 1024 * 1024
hx, hy = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(2 * n, dtype=torch.float32).pin_memory()
dx, dy = torch.empty(n, dtype=torch.float32, device="cuda"), torch.empty(2 * n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, out_elems=n, reps=3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                dx.copy_(hx, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hy[:out_elems].copy_(dy[:out_elems], non_blocking=True)
    torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t) / reps * 1e3], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


run(True, True, n, 1)
out = {"ranks_copying_concurrently": world, "bytes_h2d": n * 4, "h2d_only_ms": run(True, False), "d2h_only_ms": run(False, True),
       "both_ms_f32_out": run(True, True, n), "both_ms_f64_out": run(True, True, 2 * n)}
out["h2d_GBps"] = n * 4 / out["h2d_only_ms"] / 1e6
out["d2h_GBps"] = n * 4 / out["d2h_only_ms"] / 1e6
out["note"] = ("max over ranks; both_ms_f32_out is the floor of bench.py's e2e step (3.2 GB each way), both_ms_f64_out of the "
               "float64 drop-in leg (3.2 GB in, 6.4 GB out)")
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
