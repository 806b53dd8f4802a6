"""Oracle parity at the REAL BASELINE shapes (configs 3, 4 and 5), not at scaled-down stand-ins.

* C3  HexConv2d(64, 64, radius 2, padding 1) under autocast(bfloat16) on a 256 x 256 lattice: forward, dx, dW, db against
      the torch-CPU oracle evaluated on the same bf16-rounded operands (1e-4 of the range; the only difference left is the
      fp32 summation order), on a batch of 8 and -- for the persistent (image, band, column tile) walk, the > 2^31-byte
      offsets and the weight-gradient atomics -- on the full batch of 128 (four random images against the oracle, the
      batch-wide dW / db against the sum of per-chunk launches).
* C4  one 3 x 2160 x 3840 image group taken from BOTH ends of a batch-64 tensor (byte offsets beyond 2^32): rect->hex
      bilinear, hex->rect linear (HG_MATH_FAST <= 1e-5 * max, HG_MATH_EXACT float64 bit-identical) and all five levels of the
      average-pool pyramid, each compared with the oracle over the whole plane.
* C5  one training step of the hex CNN (tools/hexcnn.py): loss and every gradient against the same network evaluated with
      the oracle's operators on the CPU.
* HexConv2dAdaptivePadding backward.

Every call goes through the C ABI (HyGrid modules -> ctypes -> libhygrid_b200.so)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import hexframes_oracle as HO
from oracle import hygrid_oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


# ----------------------------------------------------------------------------------------------------
# C3
# ----------------------------------------------------------------------------------------------------
def _c3_oracle(xq, wq, b, gyq):
    xr, wr, br = xq.clone().requires_grad_(), wq.clone().requires_grad_(), b.clone().requires_grad_()
    ref = HO.hexconv2d(xr, wr, br, 0, 2, 1, 1)
    (ref * gyq).sum().backward()
    return ref.detach(), xr.grad, wr.grad, br.grad


def test_c3_layer_batch8_vs_oracle():
    from HyGrid import HexFrames as hf
    torch.manual_seed(31)
    m = hf.HexConv2d(64, 64, 0, 2, stride=1, padding=1).cuda()
    with torch.no_grad():                                   # bf16-exact weights: the tensor-core kernel rounds them anyway
        m.kernel.copy_(m.kernel.bfloat16().float())
    xq = torch.randn(8, 64, 256, 256).bfloat16().float()
    gyq = torch.randn(8, 64, 256, 256).bfloat16().float()
    ref, dx, dw, db = _c3_oracle(xq, m.kernel.detach().cpu(), m.bias.detach().cpu(), gyq)
    x = xq.cuda().requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    assert m._autocast_tc and y.dtype == torch.float32 and y.shape == x.shape
    (y * gyq.cuda()).sum().backward()
    assert _rel(y.detach().cpu(), ref) <= 1e-4
    assert _rel(x.grad.cpu(), dx) <= 1e-4
    assert _rel(m.kernel.grad.cpu(), dw) <= 1e-4
    assert _rel(m.bias.grad.cpu(), db) <= 1e-4


def test_c3_layer_full_batch128():
    from HyGrid import HexFrames as hf
    torch.manual_seed(32)
    N = 128
    m = hf.HexConv2d(64, 64, 0, 2, stride=1, padding=1).cuda()
    with torch.no_grad():
        m.kernel.copy_(m.kernel.bfloat16().float())
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(N, 64, 256, 256, device="cuda", generator=g).bfloat16().float().requires_grad_()   # 2.1 GB
    gy = torch.randn(N, 64, 256, 256, device="cuda", generator=g).bfloat16().float()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    (y * gy).sum().backward()
    assert x.numel() * 4 >= 2 ** 31
    # four images (both ends of the batch, where the byte offsets are largest) against the oracle
    pick = [0, 37, 90, N - 1]
    ref, dx, _, _ = _c3_oracle(x.detach()[pick].cpu(), m.kernel.detach().cpu(), m.bias.detach().cpu(), gy[pick].cpu())
    assert _rel(y.detach()[pick].cpu(), ref) <= 1e-4
    assert _rel(x.grad[pick].cpu(), dx) <= 1e-4
    # dW / db of the whole batch = sum over chunks of 8 images, each chunk a launch of the size test_c3_layer_batch8
    # pins against the oracle (float64 accumulation of the partials); one chunk from the middle of the batch is compared
    # with the oracle here as well
    dw = torch.zeros_like(m.kernel, dtype=torch.float64)
    db = torch.zeros_like(m.bias, dtype=torch.float64)
    mc = hf.HexConv2d(64, 64, 0, 2, stride=1, padding=1).cuda()
    mc.load_state_dict(m.state_dict())
    for c in range(0, N, 8):
        mc.kernel.grad = mc.bias.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yc = mc(x.detach()[c:c + 8])
        (yc * gy[c:c + 8]).sum().backward()
        dw += mc.kernel.grad.double()
        db += mc.bias.grad.double()
        if c == 64:
            _, _, dwo, dbo = _c3_oracle(x.detach()[c:c + 8].cpu(), m.kernel.detach().cpu(), m.bias.detach().cpu(), gy[c:c + 8].cpu())
            assert _rel(mc.kernel.grad.cpu(), dwo) <= 1e-4 and _rel(mc.bias.grad.cpu(), dbo) <= 1e-4
    assert _rel(m.kernel.grad.double(), dw) <= 1e-4
    assert _rel(m.bias.grad.double(), db) <= 1e-4


def test_c3_stack_four_layers_autocast_step():
    """The config as BASELINE states it -- 4 x HexConv2d(64, 64) -- on a batch of 2: the stack's output and the gradient of
    every layer against the oracle chain (each oracle layer rounds its operands to bfloat16 like the tensor-core kernels:
    2e-2 of the range, SURVEY.md 8c's bf16 tolerance, since rounding points differ after the first layer)."""
    from HyGrid import HexFrames as hf
    torch.manual_seed(33)
    net = torch.nn.Sequential(*[hf.HexConv2d(64, 64, 0, 2, stride=1, padding=1) for _ in range(4)]).cuda()
    x = torch.randn(2, 64, 256, 256)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = net(x.cuda())
    y.float().sum().backward()
    ws = [l.kernel.detach().cpu().requires_grad_() for l in net]
    bs = [l.bias.detach().cpu().requires_grad_() for l in net]
    cur = x
    for w, b in zip(ws, bs):
        cur = HO.hexconv2d(cur.bfloat16().float(), w.bfloat16().float(), b, 0, 2, 1, 1)
    cur.sum().backward()
    assert _rel(y.detach().cpu(), cur.detach()) <= 2e-2
    for l, w, b in zip(net, ws, bs):
        assert _rel(l.kernel.grad.cpu(), w.grad) <= 2e-2
        assert _rel(l.bias.grad.cpu(), b.grad) <= 2e-2


def test_c3_bfloat16_activations_end_to_end():
    """The opt-in `out_dtype = torch.bfloat16` (bfloat16 tensors at the layer boundary instead of the reference's float32
    buffer): output and data gradient are the oracle's values rounded once to bfloat16 (2^-8 relative), the weight
    gradient accumulates in float32."""
    from HyGrid import HexFrames as hf
    torch.manual_seed(34)
    m = hf.HexConv2d(64, 64, 0, 2, stride=1, padding=1).cuda()
    m.out_dtype = torch.bfloat16
    with torch.no_grad():
        m.kernel.copy_(m.kernel.bfloat16().float())
    xq = torch.randn(2, 64, 128, 256).bfloat16().float()
    gyq = torch.randn(2, 64, 128, 256).bfloat16().float()
    ref, dx, dw, db = _c3_oracle(xq, m.kernel.detach().cpu(), m.bias.detach().cpu(), gyq)
    x = xq.cuda().bfloat16().requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    assert y.dtype == torch.bfloat16 and m._autocast_tc
    (y.float() * gyq.cuda()).sum().backward()
    assert x.grad.dtype == torch.bfloat16
    assert _rel(y.detach().float().cpu(), ref) <= 8e-3
    assert _rel(x.grad.float().cpu(), dx) <= 8e-3
    assert _rel(m.kernel.grad.cpu(), dw) <= 1e-3
    assert _rel(m.bias.grad.cpu(), db) <= 1e-3


# ----------------------------------------------------------------------------------------------------
# C4
# ----------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c4_batch():
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(64, 3, 2160, 3840, device="cuda", generator=g)          # 6.4 GB: offsets of the last images > 2^32 bytes
    yield x
    del x
    torch.cuda.empty_cache()


C4_PICK = (0, 63)


def test_c4_rect_to_hex_full_planes(c4_batch):
    from HyGrid import functional as Fn
    x = c4_batch
    fast = Fn.rect_to_hex(x, None, "bilinear", out_dtype=torch.float32, math="fast")
    exact32 = Fn.rect_to_hex(x, None, "bilinear", out_dtype=torch.float32, math="exact")
    for n in C4_PICK:
        ref = O.rect_to_hex_resample(x[n].cpu().numpy(), None, "bilinear")
        assert ref.dtype == np.float64
        assert float(np.abs(fast[n].cpu().numpy() - ref).max()) <= 1e-5 * float(np.abs(ref).max())
        assert np.array_equal(exact32[n].cpu().numpy(), ref.astype(np.float32))
        exact = Fn.rect_to_hex(x[n:n + 1], None, "bilinear")                   # float64, HG_MATH_EXACT: the drop-in call
        assert exact.dtype == torch.float64 and np.array_equal(exact[0].cpu().numpy(), ref)
    near = Fn.rect_to_hex(x, None, "nearest")
    for n in C4_PICK:
        assert np.array_equal(near[n].cpu().numpy(), O.rect_to_hex_resample(x[n].cpu().numpy(), None, "nearest"))


def test_c4_hex_to_rect_full_planes(c4_batch):
    from HyGrid import functional as Fn
    x = c4_batch
    fast = Fn.hex_to_rect(x, None, "linear", out_dtype=torch.float32, math="fast", twin="np")
    exact = Fn.hex_to_rect(x, None, "linear", out_dtype=torch.float32, math="exact", twin="np")
    for n in C4_PICK:
        ref = O.hex_to_rect_resample(x[n].cpu().numpy(), None, "linear", twin="np")
        assert float(np.abs(fast[n].cpu().numpy() - ref).max()) <= 1e-5 * float(np.abs(ref).max())
        assert np.array_equal(exact[n].cpu().numpy(), ref.astype(np.float32))
        e64 = Fn.hex_to_rect(x[n:n + 1], None, "linear", twin="np")           # float64 result: bit-identical
        assert e64.dtype == torch.float64 and np.array_equal(e64[0].cpu().numpy(), ref)
    near = Fn.hex_to_rect(x, None, "nearest", twin="torch")
    for n in C4_PICK:
        assert np.array_equal(near[n].cpu().numpy(), O.hex_to_rect_resample(x[n].cpu().numpy(), None, "nearest", twin="torch"))


def test_c4_pool_pyramid_full_planes(c4_batch):
    from HyGrid import HexFrames as hf
    pool = hf.HexPool2d("average", 2, 2)
    cur = c4_batch
    refs = {n: c4_batch[n:n + 1].cpu() for n in C4_PICK}
    shapes = []
    for _ in range(5):
        cur = pool(cur)
        shapes.append(tuple(cur.shape[-2:]))
        for n in C4_PICK:
            refs[n] = HO.hexpool2d(refs[n], "average", 2, 2)
            assert cur[n:n + 1].shape == refs[n].shape
            assert float((cur[n:n + 1].cpu() - refs[n]).abs().max()) <= 1e-6
    assert shapes == [(1080, 1919), (540, 959), (270, 479), (135, 239), (67, 119)]
    mx = hf.HexPool2d("max", 2, 2)(c4_batch)
    for n in C4_PICK:
        assert torch.equal(mx[n:n + 1].cpu(), HO.hexpool2d(c4_batch[n:n + 1].cpu(), "max", 2, 2))


# ----------------------------------------------------------------------------------------------------
# C5
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["fp32-direct", "fp32-tensor-cores", "autocast"])
def test_c5_train_step_vs_oracle_model(mode):
    from HyGrid import HexFrames as hf
    from hexcnn import HexCNN
    from hexcnn_oracle import oracle_forward
    autocast = mode == "autocast"
    torch.manual_seed(51)
    model = HexCNN().cuda().train()
    g = torch.Generator().manual_seed(52)
    x = torch.randn(8, 3, 128, 128, generator=g)
    t = torch.randint(0, 10, (8,), generator=g)
    hf.set_fp32_tensor_cores(mode != "fp32-direct")
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss = torch.nn.functional.cross_entropy(model(x.cuda()).float(), t.cuda())
        loss.backward()
    finally:
        hf.set_fp32_tensor_cores(True)
    params = {k: v.detach().cpu().clone().requires_grad_() for k, v in model.named_parameters()}
    ref = torch.nn.functional.cross_entropy(oracle_forward(params, x, autocast=autocast), t)
    ref.backward()
    # fp32-direct: the CUDA-core stencil and the library's batch norm against torch-CPU arithmetic (1e-4 / 1e-3).
    # fp32-tensor-cores (the default for float32 callers): c2 / c3 run as three bfloat16 passes over split operands.  Layer by
    #   layer that is 5e-6 of the range (test_fp32_tensor_core_route_keeps_fp32_accuracy; measured at exactly these layer shapes),
    #   but a perturbation of that size flips a few max-pool arg-maxima / ReLU signs that sit on near-ties, which re-routes their
    #   gradients: measured 7.6e-3 on c2's kernel gradient, 1.5e-3 on c1's -- asserted at 2e-2 (the direct path's 1e-6 flips fewer).
    # autocast: the oracle network rounds to bfloat16 exactly where the product does (conv operands, output gradients, the
    #   classifier); what is left is summation order and values that sit on a rounding boundary -- measured 2.1e-2 for the
    #   deepest kernel, < 1e-2 elsewhere, asserted at 4e-2 (bf16 2e-2 per layer).
    tol = 2e-2 if autocast else 1e-4
    gtol = {"fp32-direct": 1e-3, "fp32-tensor-cores": 2e-2, "autocast": 4e-2}[mode]
    assert abs(float(loss) - float(ref)) <= tol * max(1.0, abs(float(ref)))
    for k, v in model.named_parameters():
        assert v.grad is not None, k
        assert _rel(v.grad.cpu(), params[k].grad) <= gtol, k
    # running statistics of the three batch norms
    sd = model.state_dict()
    assert all(bool(torch.isfinite(sd[f"{b}.bn.running_var"]).all()) for b in ("c1", "c2", "c3"))


# ----------------------------------------------------------------------------------------------------
# HexConv2dAdaptivePadding backward; even-rows-only output
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [(2, 1, 1, 19, 23), (2, 2, 1, 20, 21), (3, 1, 1, 17, 18), (2, 1, 2, 16, 16), (3, 2, 1, 25, 14)])
def test_adaptive_padding_forward_backward_vs_oracle(cfg):
    from HyGrid import HexFrames as hf
    r, s, d, H, W = cfg
    torch.manual_seed(61)
    m = hf.HexConv2dAdaptivePadding(5, 7, 0, r, stride=s, dilation=d).cuda()
    x = torch.randn(2, 5, H, W)
    xr = x.clone().requires_grad_()
    wr, br = m.kernel.detach().cpu().requires_grad_(), m.bias.detach().cpu().requires_grad_()
    pl, pr, pt, pb = HO.adaptive_padding(H, W, r, s, d)
    ref = HO.hexconv2d(torch.nn.functional.pad(xr, (pl, pr, pt, pb)), wr, br, 0, r, s, 0, d, 1)
    gy = torch.randn_like(ref)
    (ref * gy).sum().backward()
    xg = x.cuda().requires_grad_()
    y = m(xg)
    assert y.shape == ref.shape and _rel(y.detach().cpu(), ref.detach()) <= 1e-4
    (y * gy.cuda()).sum().backward()
    assert _rel(xg.grad.cpu(), xr.grad) <= 1e-4
    assert _rel(m.kernel.grad.cpu(), wr.grad) <= 1e-3
    assert _rel(m.bias.grad.cpu(), br.grad) <= 1e-3


@pytest.mark.parametrize("cfg", [(3, 9, 2, 1, 1, 0, 0), (3, 8, 2, 1, 1, 0, 1), (5, 12, 3, 1, 1, 0, 0), (1, 7, 2, 1, 1, 1, 0),
                                 (4, 11, 2, 2, 1, 0, 1), (5, 13, 2, 1, 2, 0, 0)])
def test_even_rows_only_output(cfg):
    """A padded height in [k_h, k_h + s): the reference skips the odd conv and returns the even-row result alone
    (HexFrames.py:163-164; pinned by tests/golden/live_check.py)."""
    from HyGrid import HexFrames as hf
    H, W, r, s, d, pad, off = cfg
    torch.manual_seed(62)
    m = hf.HexConv2d(2, 3, off, r, stride=s, padding=pad, dilation=d).cuda()
    x = torch.randn(2, 2, H, W)
    xr = x.clone().requires_grad_()
    ref = HO.hexconv2d(xr, m.kernel.detach().cpu(), m.bias.detach().cpu(), off, r, s, pad, d, 1)
    assert ref.shape[2] == 1
    xg = x.cuda().requires_grad_()
    y = m(xg)
    assert y.shape == ref.shape and _rel(y.detach().cpu(), ref.detach()) <= 1e-4
    ref.sum().backward(); y.sum().backward()
    assert _rel(xg.grad.cpu(), xr.grad) <= 1e-4


# ----------------------------------------------------------------------------------------------------
# reflect / replicate / circular frames resolved inside the conv loaders (no padded copy of x)
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["reflect", "replicate", "circular"])
@pytest.mark.parametrize("cfg", [
    # Cin, Cout, H, W, radius, stride, pad, offset, tensor cores
    (4, 6, 20, 22, 2, 1, 2, 1, False),
    (5, 7, 17, 19, 3, 2, 3, 0, False),
    (64, 64, 40, 136, 2, 1, 1, 0, True),
    (32, 48, 33, 130, 2, 1, 2, 1, True),
])
def test_conv_padding_modes_in_kernel(mode, cfg):
    from HyGrid import HexFrames as hf
    from HyGrid import _native as nv
    Cin, Cout, H, W, r, s, pad, off, tc = cfg
    torch.manual_seed(71)
    m = hf.HexConv2d(Cin, Cout, off, r, stride=s, padding=pad, padding_mode=mode).cuda()
    x = torch.randn(2, Cin, H, W)
    if tc:                                                   # bf16-exact operands: what is left is the summation order
        x = x.bfloat16().float()
        with torch.no_grad():
            m.kernel.copy_(m.kernel.bfloat16().float())
        m.algo = 2
    xr = x.clone().requires_grad_()
    wr, br = m.kernel.detach().cpu().requires_grad_(), m.bias.detach().cpu().requires_grad_()
    ref = HO.hexconv2d(xr, wr, br, off, r, s, pad, 1, 1, padding_mode=mode)
    gy = torch.randn_like(ref)
    if tc:
        gy = gy.bfloat16().float()
    (ref * gy).sum().backward()
    xg = x.cuda().requires_grad_()
    nv.reset_launch_count()
    y = m(xg)
    assert nv.launch_count() == 1, "the frame must be resolved by the conv kernel itself, not by a pad launch"
    assert y.shape == ref.shape and _rel(y.detach().cpu(), ref.detach()) <= 1e-4
    (y * gy.cuda()).sum().backward()
    assert _rel(xg.grad.cpu(), xr.grad) <= 1e-4
    assert _rel(m.kernel.grad.cpu(), wr.grad) <= 1e-3
    assert _rel(m.bias.grad.cpu(), br.grad) <= 1e-3


@pytest.mark.parametrize("mode", ["reflect", "replicate"])
def test_hexconvmodule_explicit_padding_is_one_launch(mode):
    """HexConvModule(padding_mode='reflect' | 'replicate') puts a padding layer in front of a padding = 0 conv
    (HexModules.py:185-190); here the frame is folded into the conv kernel: one launch in inference, same numbers as the
    module's own two-step route (pad kernel, then conv)."""
    from HyGrid import HexModules as hm
    from HyGrid import _native as nv
    torch.manual_seed(72)
    m = hm.HexConvModule(8, 12, 1, 2, padding=2, padding_mode=mode).cuda().eval()
    x = torch.randn(2, 8, 21, 26).cuda()
    with torch.no_grad():
        nv.reset_launch_count()
        y = m(x)
        assert nv.launch_count() == 1
        two_step = torch.relu(m.conv(m.padding_layer(x)))
    assert y.shape == two_step.shape and _rel(y.cpu(), two_step.cpu()) <= 1e-5
    xr = x.cpu()
    ref = torch.relu(HO.hexconv2d(torch.nn.functional.pad(xr, (2, 2, 2, 2), mode), m.conv.kernel.detach().cpu(), m.conv.bias.detach().cpu(),
                                  1, 2, 1, 0))
    assert _rel(y.cpu(), ref) <= 1e-4


# ----------------------------------------------------------------------------------------------------
# float32 callers on the tensor cores: three bfloat16 passes over split operands
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [
    # N, Cin, Cout, H, W, pad, offset, padding mode, bias
    (2, 64, 64, 70, 300, 1, 0, "constant", True),
    (1, 32, 64, 33, 136, 2, 1, "reflect", True),
    (2, 64, 128, 20, 128, 1, 0, "constant", False),
    (1, 128, 64, 24, 130, 1, 1, "constant", True),
    (8, 32, 64, 64, 63, 1, 0, "constant", False),        # the C5 network's second and third layers (odd widths: LDG staging)
    (8, 64, 128, 32, 31, 1, 0, "constant", False),
])
def test_fp32_tensor_core_route_keeps_fp32_accuracy(cfg):
    """HexConv2d called in float32 (no autocast) with a dense channel contraction runs as three bfloat16 tcgen05 passes over
    split operands: forward, dx, dW, db within the float32 contract (1e-4 of the range) of the oracle on the UNROUNDED
    operands -- and far closer than one bfloat16 pass (2e-3)."""
    from HyGrid import HexFrames as hf
    from HyGrid import _native as nv
    N, Cin, Cout, H, W, pad, off, mode, has_bias = cfg
    torch.manual_seed(81)
    m = hf.HexConv2d(Cin, Cout, off, 2, padding=pad, padding_mode=mode, bias=has_bias).cuda()
    x = torch.randn(N, Cin, H, W)
    xr = x.clone().requires_grad_()
    wr = m.kernel.detach().cpu().requires_grad_()
    br = m.bias.detach().cpu().requires_grad_() if has_bias else None
    ref = HO.hexconv2d(xr, wr, br, off, 2, 1, pad, 1, 1, padding_mode=mode)
    gy = torch.randn_like(ref)
    (ref * gy).sum().backward()
    xg = x.cuda().requires_grad_()
    y = m(xg)
    assert nv.last_launch().startswith("hexconv_umma"), nv.last_launch()
    assert y.dtype == torch.float32 and _rel(y.detach().cpu(), ref.detach()) <= 1e-4
    (y * gy.cuda()).sum().backward()
    assert _rel(xg.grad.cpu(), xr.grad) <= 1e-4
    assert _rel(m.kernel.grad.cpu(), wr.grad) <= 1e-4
    if has_bias:
        assert _rel(m.bias.grad.cpu(), br.grad) <= 1e-4
    # the switch: plain float32 FMAs on the CUDA cores
    hf.set_fp32_tensor_cores(False)
    try:
        y2 = m(x.cuda())
        assert nv.last_launch() == "hexconv_fwd_direct" and _rel(y2.detach().cpu(), ref.detach()) <= 1e-4
    finally:
        hf.set_fp32_tensor_cores(True)
    # inference epilogues on the split route: fused ReLU, fused BN affine
    with torch.no_grad():
        yr = m(x.cuda(), relu=True)
        assert _rel(yr.cpu(), ref.detach().clamp_min(0)) <= 1e-4
        sc, sh = torch.rand(Cout).cuda() + 0.5, torch.randn(Cout).cuda()
        ya = m(x.cuda(), relu=True, affine=(sc, sh))
        want = torch.relu((ref.detach() - (br.detach().view(1, -1, 1, 1) if has_bias else 0)) * sc.cpu().view(1, -1, 1, 1) + sh.cpu().view(1, -1, 1, 1))
        assert _rel(ya.cpu(), want) <= 1e-4
