#!/usr/bin/env python
"""BASELINE config 5: hex CNN training step, data-parallel over the GPUs of one box.

    python tools/hexcnn_ddp.py [--batch 64] [--steps 10]                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/hexcnn_ddp.py --batch 64 --steps 10                                 # N GPUs, 64 images each

Model (builder-defined): HexConvModule 3->32->64->128 (BN + ReLU), HexPool2d('max', 2, 2) between,
global average, Linear -> 10, on 128 x 128 hex lattices.  Every gradient lives in one flat fp32 bucket
(HyGrid.distributed.FlatGradBucket) cut into a few groups; each group is all-reduced over NCCL as soon as backward
has produced its last gradient (the collective overlaps the backward of the earlier layers); conv / pool kernels are
the library's and write their weight gradients straight into the bucket.  Prints one JSON line: ms per step (max over ranks, CUDA events), images/s, bucket size, and whether
all ranks hold identical parameters after the steps (the all-reduce did its job)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from HyGrid import _native as nv  # noqa: E402
from HyGrid.distributed import FlatGradBucket, shard_range  # noqa: E402
from hexcnn import HexCNN  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--hw", type=int, default=128)
    ap.add_argument("--autocast", action="store_true")
    ap.add_argument("--graph", action="store_true", help="capture the whole training step in one CUDA graph and replay it")
    ap.add_argument("--groups", type=int, default=3, help="bucket groups; every group is all-reduced as soon as its gradients landed")
    ap.add_argument("--blocking", action="store_true", help="round-1 behaviour: one blocking all-reduce after backward")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)                                     # identical initial weights on every rank
    model = HexCNN().to(dev)
    bucket = FlatGradBucket(model.parameters(), groups=1 if a.blocking else ([3, 3, 5] if a.groups == 3 else a.groups),
                            overlap=not a.blocking)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, fused=True)
    g = torch.Generator(device="cpu").manual_seed(1)
    data = torch.randn(a.batch * world, 3, a.hw, a.hw, generator=g)
    target = torch.randint(0, 10, (a.batch * world,), generator=g)
    lo, hi = shard_range(a.batch * world, rank, world)
    x, t = data[lo:hi].to(dev), target[lo:hi].to(dev)

    def step():
        bucket.zero_()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.autocast):
            loss = nn.functional.cross_entropy(model(x).float(), t)
        loss.backward()                                      # groups of the flat bucket are all-reduced from inside backward
        bucket.finish()                                      # join (blocking mode: the one collective happens here)
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if a.graph:
        # the step is launch-bound (about 130 kernels of a few microseconds): one captured graph replays it without the
        # host-side launch cost.  Inputs, parameters, the flat gradient bucket and the optimizer state are all static.
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = step()
        eager_step = step

        def step():                                           # noqa: F811
            graph.replay()
            return static_loss
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    nv.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    check = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
    same = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        lo_, hi_ = check.clone(), check.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        same = bool(torch.allclose(lo_, hi_, rtol=0, atol=1e-6 * float(hi_.abs()) + 1e-9))
    if rank == 0:
        print(json.dumps({"workload": f"hex CNN train step, {a.batch} img/GPU x {world} GPU, {a.hw}x{a.hw} hex lattice",
                          "ms_per_step": round(float(ms), 3), "images_per_s": round(a.batch * world / float(ms) * 1e3, 1),
                          "hex_mpix_per_s": round(a.batch * world * a.hw * a.hw / float(ms) / 1e3, 1),
                          "grad_bucket_bytes": bucket.nbytes, "ranks_in_sync": same, "loss": round(float(loss), 4),
                          "library_launches_per_step": nv.launch_count() // a.steps, "autocast": a.autocast,
                          "cuda_graph": a.graph, "overlap": not a.blocking, "group_bytes": list(bucket.group_bytes())}), flush=True)
    if world > 1:
        if a.graph:
            # tearing the process group down while a captured graph still holds NCCL kernels hung the run on 8 GPUs
            # (the result line above had been printed): drop the graph first and leave without the collective teardown
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
