#!/usr/bin/env python
"""Where the host time of one eager C5 training step goes (the step is launch-bound: ~130 launches in ~2 ms).

    python tools/profile_c5_host.py [--steps 200]

Prints the host enqueue time per step (no synchronisation inside the loop, the GPU queue kept short by a sync every 10
steps), the GPU time per step (CUDA events) and a cProfile table of the step sorted by own time."""
import argparse
import cProfile
import io
import json
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from hexcnn import HexCNN  # noqa: E402
from HyGrid.distributed import FlatGradBucket  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--fused-sgd", type=int, default=1)
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    model = HexCNN().to(dev)
    x = torch.randn(64, 3, 128, 128, device=dev)
    t = torch.randint(0, 10, (64,), device=dev)
    bucket = FlatGradBucket(model.parameters(), groups=[3, 3, 5], overlap=True)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, **({"fused": True} if a.fused_sgd else {}))

    def step():
        bucket.zero_()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(model(x).float(), t)
        loss.backward()
        bucket.finish()
        opt.step()
        return loss

    for _ in range(10):
        step()
    torch.cuda.synchronize()
    # host enqueue time: the queue is drained every 10 steps so that the host never blocks on a full launch queue
    host = 0.0
    for i in range(a.steps):
        if i % 10 == 0:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        host += time.perf_counter() - t0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"host_enqueue_ms_per_step": round(host / a.steps * 1e3, 4), "back_to_back_ms_per_step": round(e0.elapsed_time(e1) / a.steps, 4),
                      "fused_sgd": bool(a.fused_sgd)}), flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        step()
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
    print(s.getvalue())


if __name__ == "__main__":
    main()
