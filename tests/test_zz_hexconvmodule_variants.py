"""``HexConvModule.forward`` against the reference's own forward loop (HexModules.py:275-288).

The reference runs the three layers one after the other in ``order``; this build fuses where it can (ReLU and eval-mode
BatchNorm into the conv epilogue, BatchNorm + ReLU into one streaming pass), which adds control flow the reference does
not have.  Here every combination of layer order, ``activate`` / ``norm`` flags, training / eval, grad / no-grad, norm
and activation type, explicit padding mode and spectral norm is compared with the plain loop evaluated on the module's
own sub-layers (the fused paths must be numerically equivalent: 1e-4 of the range for outputs, statistics and data
gradients -- SURVEY.md 8c's fp32 conv tolerance -- and 1e-3 for parameter gradients, which sum over every pixel with
float atomics) -- outputs, running statistics and gradients."""
import copy
import itertools

import pytest
import torch

pytestmark = pytest.mark.gpu

ORDERS = list(itertools.permutations(("conv", "norm", "act")))


def plain_forward(m, x, activate=True, norm=True):
    """HexModules.py:279-288, literally."""
    for layer in m.order:
        if layer == "conv":
            if m.with_explicit_padding:
                x = m.padding_layer(x)
            x = m.conv(x)
        elif layer == "norm" and norm and m.with_norm:
            x = m.norm(x)
        elif layer == "act" and activate and m.with_activation:
            x = m.activate(x)
    return x


def _close(a, b, what, rel=1e-4):
    """Fused and plain paths use different kernels for the same arithmetic (library BatchNorm with float64 sums vs cuDNN's,
    BN scale folded into the weights vs applied to the conv result, float atomics in the weight gradient): equal to ~1e-5
    of the range (tests/test_gpu_hexframes.py measures 2e-5 .. 1e-4 for those pairs); a wrong branch is off by O(1)."""
    scale = max(1.0, float(b.abs().max()))
    assert a.shape == b.shape and float((a - b).abs().max()) <= rel * scale, (what, float((a - b).abs().max()), scale)


def _check(cfg, training, grad, activate=True, norm=True, Cin=4, Cout=6):
    from HyGrid import HexModules as hm
    torch.manual_seed(11)
    norm_first = cfg.get("norm_cfg") is not None and cfg.get("order", ("conv", "norm", "act")).index("norm") < cfg.get("order", ("conv", "norm", "act")).index("conv")
    m = hm.HexConvModule(Cin, Cout, 0, 2, padding=1, **cfg).cuda()
    for p in m.parameters():                       # non-trivial affine parameters / statistics
        with torch.no_grad():
            p.add_(0.1 * torch.randn_like(p))
    if m.with_norm and getattr(m.norm, "running_mean", None) is not None:
        with torch.no_grad():
            m.norm.running_mean.normal_(0, 0.3)
            m.norm.running_var.uniform_(0.5, 1.5)
    m.train(training)
    ref = copy.deepcopy(m)
    x = torch.randn(2, Cin, 10, 12).cuda()
    xa, xb = x.clone().requires_grad_(grad), x.clone().requires_grad_(grad)
    what = (cfg, training, grad, activate, norm)
    with torch.set_grad_enabled(grad):
        ya = m(xa * 1.0, activate=activate, norm=norm)            # (* 1.0: an in-place ReLU may be the first layer)
        yb = plain_forward(ref, xb * 1.0, activate, norm)
    _close(ya.detach().float().cpu(), yb.detach().float().cpu(), what)
    if m.with_norm and getattr(m.norm, "running_mean", None) is not None:
        _close(m.norm.running_mean.cpu(), ref.norm.running_mean.cpu(), what + ("running_mean",))
        _close(m.norm.running_var.cpu(), ref.norm.running_var.cpu(), what + ("running_var",))
    if grad:
        g = torch.randn_like(ya)
        ya.backward(g)
        yb.backward(g)
        _close(xa.grad.cpu(), xb.grad.cpu(), what + ("dx",), 1e-4)
        for (na, pa), (_, pb) in zip(m.named_parameters(), ref.named_parameters()):
            if pa.grad is None and pb.grad is None:
                continue
            _close(pa.grad.cpu(), pb.grad.cpu(), what + (na,), 1e-3)
    del norm_first


@pytest.mark.parametrize("order", ORDERS)
def test_every_order_with_batchnorm_and_relu(order):
    cfg = dict(norm_cfg=dict(type="BN"), order=order)
    for training, grad in ((True, True), (False, False), (False, True), (True, False)):
        for activate, norm in ((True, True), (False, True), (True, False), (False, False)):
            _check(cfg, training, grad, activate, norm)


@pytest.mark.parametrize("cfg", [
    dict(),                                                             # conv (+bias, 'auto') -> ReLU
    dict(act_cfg=None),
    dict(act_cfg=dict(type="LeakyReLU", negative_slope=0.1)),
    dict(act_cfg=dict(type="PReLU")),
    dict(act_cfg=dict(type="Tanh")),
    dict(act_cfg=dict(type="GELU")),
    dict(norm_cfg=dict(type="GN", num_groups=2)),
    dict(norm_cfg=dict(type="IN")),
    dict(norm_cfg=dict(type="BN"), bias=True),
    dict(norm_cfg=dict(type="BN", momentum=None)),
    dict(norm_cfg=dict(type="BN", affine=False)),
    dict(norm_cfg=dict(type="BN", track_running_stats=False)),
    dict(norm_cfg=dict(type="BN"), act_cfg=dict(type="LeakyReLU")),
    dict(padding_mode="reflect"),
    dict(padding_mode="replicate", norm_cfg=dict(type="BN")),
    dict(padding_mode="circular"),
    dict(with_spectral_norm=True),
    dict(inplace=False, norm_cfg=dict(type="BN")),
    dict(groups=2, norm_cfg=dict(type="BN")),
    dict(stride=2, dilation=1, norm_cfg=dict(type="BN")),
    dict(conv_cfg=dict(type="HexConv2dAdaptivePadding"), norm_cfg=dict(type="BN")),     # eval: fused affine epilogue
    dict(conv_cfg=dict(type="HexConv2dAdaptivePadding")),                               # eval: fused ReLU epilogue
])
def test_layer_types_padding_modes_and_flags(cfg):
    for training, grad in ((True, True), (False, False)):
        _check(cfg, training, grad)


def test_tensor_core_sized_module_under_autocast():
    """64 -> 64 channels: eval-mode conv + BN + ReLU collapses into one tcgen05 launch; compared with the plain loop on
    the same module at the bf16 tolerance of SURVEY.md 8c (2e-2 relative)."""
    from HyGrid import HexModules as hm
    torch.manual_seed(3)
    m = hm.HexConvModule(64, 64, 0, 2, padding=1, norm_cfg=dict(type="BN")).cuda().eval()
    with torch.no_grad():
        m.norm.running_mean.normal_(0, 0.2)
        m.norm.running_var.uniform_(0.5, 1.5)
    x = torch.randn(2, 64, 24, 40).cuda()
    with torch.no_grad():
        want = plain_forward(copy.deepcopy(m), x.clone())
        with torch.autocast("cuda", dtype=torch.bfloat16):
            got = m(x.clone())
    assert got.shape == want.shape and float((got.float() - want).abs().max()) <= 2e-2 * float(want.abs().max())
