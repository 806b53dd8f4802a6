"""Host-side logic of the N > 1 path on CPU: world_size-2 gloo processes exercise the batch sharding and the
flat gradient bucket (the one collective of the path: conv weight / bias gradients)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from HyGrid.distributed import FlatGradBucket, shard_batch, shard_range


def test_shard_range_tiles_the_batch():
    for n in (0, 1, 7, 64, 256, 257):
        for world in (1, 2, 3, 4, 8):
            pieces = [shard_range(n, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
            sizes = [b - a for a, b in pieces]
            assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
    x = torch.arange(10).view(10, 1)
    assert torch.equal(torch.cat([shard_batch(x, r, 3) for r in range(3)]), x)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        kern = torch.nn.Parameter(torch.zeros(4, 2, 1, 7))          # a HexConv2d-shaped kernel
        bucket = FlatGradBucket(list(lin.parameters()) + [kern])
        data = torch.arange(8 * 5, dtype=torch.float32).view(8, 5) / 10
        x = shard_batch(data)                                         # rank/world from the process group
        assert x.shape[0] == 4 and float(x[0, 0]) == (0.0 if rank == 0 else 2.0)
        loss = lin(x).square().sum() + (kern * (rank + 1)).sum()
        loss.backward()
        assert lin.weight.grad.data_ptr() == bucket.view(0).data_ptr()     # accumulated in place
        bucket.all_reduce(average=True)
        # reference: the same step on the whole batch in one process, averaged over 2 shards
        lin2 = torch.nn.Linear(5, 3)
        lin2.load_state_dict(lin.state_dict())
        lin2(data).square().sum().backward()
        ok = torch.allclose(lin.weight.grad, lin2.weight.grad / 2, atol=1e-5)
        ok &= torch.allclose(lin.bias.grad, lin2.bias.grad / 2, atol=1e-5)
        ok &= torch.allclose(kern.grad, torch.full_like(kern, 1.5))
        work = bucket.all_reduce(average=False, async_op=True)       # async variant: sum of identical buffers
        before = bucket.flat.clone()
        work.wait()
        ok &= torch.allclose(bucket.flat, before * 2) or torch.allclose(bucket.flat, before)   # wait() may have landed already
        bucket.zero_()
        ok &= float(lin.weight.grad.abs().sum()) == 0.0
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _overlap_worker(rank, world, port, ret):
    """Overlapped exchange: three Linear layers, one bucket group per layer.  Backward produces the gradients of the
    LAST layer first; its group must be all-reduced before the first layer's gradients even exist, and the reduced
    values must equal the single-process full-batch gradients.  A bfloat16 parameter in the float32 bucket (mixed
    dtype) must come back reduced in its own .grad, and the gradient-sink path (what the hex-conv weight-gradient kernel
    uses) must count as "landed" without an autograd hook."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
        half = torch.nn.Parameter(torch.ones(7, dtype=torch.bfloat16))
        sink = torch.nn.Parameter(torch.zeros(4, 2, 1, 7))              # stands for a HexConv2d kernel
        params = list(net.parameters()) + [half, sink]
        bucket = FlatGradBucket(params, groups=4, overlap=True)
        assert len(bucket.bounds) == 4 and sink._hg_grad_sink is not None and half._hg_grad_sink is None
        data = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 20
        x = shard_batch(data)
        ok = True
        for step in range(2):                                          # second step: the bucket state must have been reset
            bucket.zero_()
            loss = net(x).square().sum() + (half.float() * (rank + 1)).sum()
            loss.backward()
            # what the conv kernel does: accumulate into the sink's view, then report
            sink._hg_grad_sink.view.add_(float(rank + 1))
            sink._hg_grad_sink.landed()
            trace = list(bucket.trace)
            bucket.finish()
            # order: gradients arrive last layer first; every group is launched right after its own last gradient
            # and -- the overlap -- before the first layer's gradients even exist
            pos = {ev: i for i, ev in enumerate(trace)}
            for g, (lo, hi) in enumerate(bucket.bounds):
                ok &= ("launch", g) in pos and all(pos[("ready", i)] < pos[("launch", g)] for i in range(lo, hi))
            first_launch = min(pos[("launch", g)] for g in range(len(bucket.bounds)))
            ok &= first_launch < pos[("ready", 0)] and first_launch < pos[("ready", 1)]
            ok &= pos[("launch", bucket.group_of[4])] < pos[("ready", 0)]          # the last layer's group went out early
            ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
            ref.load_state_dict(net.state_dict())
            ref(data).square().sum().backward()
            for p, q in zip(net.parameters(), ref.parameters()):
                ok &= torch.allclose(p.grad, q.grad / 2, atol=1e-5)
            ok &= half.grad is not None and half.grad.dtype == torch.bfloat16 and torch.allclose(half.grad.float(), torch.full((7,), 1.5))
            ok &= torch.allclose(sink.grad, torch.full_like(sink, 1.5))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_overlapped_group_all_reduce_gloo_world2():
    port = 31500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_overlap_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_bucket_single_process_sink_and_mixed_dtype():
    """No process group: finish() must still hand every gradient back (mixed dtype written back, sinks counted)."""
    lin = torch.nn.Linear(3, 2)
    half = torch.nn.Parameter(torch.ones(5, dtype=torch.bfloat16))
    bucket = FlatGradBucket(list(lin.parameters()) + [half], groups=2, overlap=True)
    for _ in range(2):
        bucket.zero_()
        (lin(torch.ones(4, 3)).sum() + (half.float() * 3).sum()).backward()
        bucket.finish()
        assert torch.allclose(lin.weight.grad, torch.full((2, 3), 4.0)) and lin.weight.grad.data_ptr() == bucket.view(0).data_ptr()
        assert half.grad.dtype == torch.bfloat16 and torch.allclose(half.grad.float(), torch.full((5,), 3.0))
    bucket.detach()
    assert getattr(lin.weight, "_hg_grad_sink", None) is None


def test_flat_bucket_all_reduce_gloo_world2():
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
