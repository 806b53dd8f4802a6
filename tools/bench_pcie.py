#!/usr/bin/env python
"""Ceiling of the end-to-end leg: concurrent H2D + D2H of pinned 3.2 GB buffers (the C2 step moves 3.2 GB each way)."""
import json
import time
import torch

n = 256 * 3 * 1024 * 1024
hx, hy = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
dx, dy = torch.empty(n, dtype=torch.float32, device="cuda"), torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=3):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                dx.copy_(hx, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hy.copy_(dy, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


run(True, True, 1)
out = {"bytes_each_way": n * 4, "h2d_only_ms": run(True, False), "d2h_only_ms": run(False, True), "both_ms": run(True, True)}
out["h2d_GBps"] = n * 4 / out["h2d_only_ms"] / 1e6
out["d2h_GBps"] = n * 4 / out["d2h_only_ms"] / 1e6
out["both_GBps_each_way"] = n * 4 / out["both_ms"] / 1e6
print(json.dumps(out))
