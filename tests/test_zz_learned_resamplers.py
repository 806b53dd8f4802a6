"""The three learned lattice resamplers the reference retired into "HyGrid/codes in old versions.txt"
(Hex_to_Square_Conv2d_by_Double_Stride :1-66, Square_to_Hex_Conv2d_by_Double_Stride :421-493,
Hex_to_Square_original_resolution :587-636).

Pin: tests/golden/make_resampler_golden.py exec'd the retired class bodies unmodified and stored inputs, kernels, outputs and
autograd gradients.  CPU: the oracle's closed forms (oracle/hexframes_oracle.py) against those fixtures.  GPU: the product
modules (HyGrid.HexResamplers over hg_dwtaps_*) against the fixtures -- forward 1e-5, data gradient 1e-5, kernel gradient
1e-4 of the range (float atomics) -- plus constructor contract and initial weights."""
import os

import numpy as np
import pytest
import torch

from oracle import hexframes_oracle as HO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resampler_golden.npz")
KIND = {0: "h2s", 1: "s2h", 2: "h2so"}
MODE = {0: "constant", 1: "reflect", 2: "replicate"}


def _cases():
    G = np.load(GOLDEN)
    for n in range(int(G["count"])):
        kind, C, off, f, pad, mode = (int(v) for v in G[f"{n}_cfg"])
        yield n, KIND[kind], C, off, f, pad, MODE[mode], G


def _oracle(kind, x, k, off, f, pad, mode):
    if kind == "h2s":
        return HO.hex_to_square_double_stride(x, k, off, f, pad, mode)
    if kind == "s2h":
        return HO.square_to_hex_double_stride(x, k, pad, mode)
    return HO.hex_to_square_original_resolution(x, k, off, pad, mode)


def test_oracle_matches_the_retired_classes():
    seen = set()
    for n, kind, C, off, f, pad, mode, G in _cases():
        x = torch.from_numpy(G[f"{n}_x"]).requires_grad_()
        k = torch.from_numpy(G[f"{n}_kernel"]).requires_grad_()
        y = _oracle(kind, x, k, off, f, pad, mode)
        assert tuple(y.shape) == G[f"{n}_y"].shape, (n, kind)
        np.testing.assert_allclose(y.detach().numpy(), G[f"{n}_y"], rtol=0, atol=2e-6)
        (y * torch.from_numpy(G[f"{n}_gy"])).sum().backward()
        np.testing.assert_allclose(x.grad.numpy(), G[f"{n}_dx"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(k.grad.numpy(), G[f"{n}_dk"], rtol=1e-5, atol=2e-5)
        init = {"h2s": lambda: HO.resampler_weight(f, "hex_to_square"), "s2h": lambda: HO.resampler_weight(2, "square_to_hex").reshape(-1),
                "h2so": lambda: HO.resampler_weight(2, "original_resolution").reshape(-1)}[kind]()
        assert np.array_equal(G[f"{n}_init"][0], init.numpy())
        seen.add(kind)
    assert seen == {"h2s", "s2h", "h2so"}
    with pytest.raises(ValueError):
        HO.hex_to_square_original_resolution(torch.zeros(1, 1, 2, 5), torch.ones(1, 4))


def test_module_contract_on_cpu():
    """Constructors, attributes, parameter shapes and initial weights need no GPU."""
    from HyGrid import HexResamplers as hr
    m = hr.Hex_to_Square_Conv2d_by_Double_Stride(3, 1, 4, padding=1)
    assert m.kernel.shape == (3, 4, 4) and m.stride == (4, 7) and m.k_w == 10 and m.padded_even_odd_offset == 0
    assert torch.equal(m.kernel.detach()[1], HO.resampler_weight(4, "hex_to_square"))
    with pytest.raises(Exception):
        hr.Hex_to_Square_Conv2d_by_Double_Stride(3, 0, 3)
    s = hr.Square_to_Hex_Conv2d_by_Double_Stride(2, 2)
    assert s.kernel.shape == (2, 4) and torch.allclose(s.kernel.detach()[0], torch.full((4,), 0.25))
    o = hr.Hex_to_Square_original_resolution(5, 1, padding=2)
    assert o.kernel.shape == (5, 4) and not o.kernel.requires_grad and o.offset == 1
    assert hr.Hex_to_Square_original_resolution(1, 0, trainable=True).kernel.requires_grad
    with pytest.raises(Exception):                       # CPU tensors: there is no CPU path
        m(torch.zeros(1, 3, 8, 8))


@pytest.mark.gpu
def test_gpu_modules_match_the_retired_classes():
    from HyGrid import HexResamplers as hr
    for n, kind, C, off, f, pad, mode, G in _cases():
        if kind == "h2s":
            m = hr.Hex_to_Square_Conv2d_by_Double_Stride(C, off, f, padding=pad, padding_mode=mode)
        elif kind == "s2h":
            m = hr.Square_to_Hex_Conv2d_by_Double_Stride(C, f, padding=pad, padding_mode=mode)
        else:
            m = hr.Hex_to_Square_original_resolution(C, off, padding=pad, padding_mode=mode, trainable=True)
        m = m.cuda()
        assert np.array_equal(m.kernel.detach().cpu().numpy(), G[f"{n}_init"]), (n, kind)
        with torch.no_grad():
            m.kernel.copy_(torch.from_numpy(G[f"{n}_kernel"]))
        x = torch.from_numpy(G[f"{n}_x"]).cuda().requires_grad_()
        y = m(x)
        assert tuple(y.shape) == G[f"{n}_y"].shape and y.dtype == torch.float32, (n, kind, y.shape)
        np.testing.assert_allclose(y.detach().cpu().numpy(), G[f"{n}_y"], rtol=0, atol=1e-5)
        (y * torch.from_numpy(G[f"{n}_gy"]).cuda()).sum().backward()
        np.testing.assert_allclose(x.grad.cpu().numpy(), G[f"{n}_dx"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(m.kernel.grad.cpu().numpy(), G[f"{n}_dk"], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_gpu_modules_vs_oracle_on_larger_lattices():
    """BASELINE-like lattices (256 x 256, 64 channels): forward and gradients against the oracle; error agreement."""
    from HyGrid import HexResamplers as hr
    torch.manual_seed(9)
    x = torch.randn(2, 64, 256, 256)
    for m, fn in ((hr.Hex_to_Square_Conv2d_by_Double_Stride(64, 1, 4, padding=1), lambda a, k: HO.hex_to_square_double_stride(a, k, 1, 4, 1)),
                  (hr.Square_to_Hex_Conv2d_by_Double_Stride(64, 2), lambda a, k: HO.square_to_hex_double_stride(a, k)),
                  (hr.Hex_to_Square_original_resolution(64, 0, trainable=True), lambda a, k: HO.hex_to_square_original_resolution(a, k, 0))):
        m = m.cuda()
        with torch.no_grad():
            m.kernel.add_(0.05 * torch.randn_like(m.kernel))
        xr, kr = x.clone().requires_grad_(), m.kernel.detach().cpu().requires_grad_()
        ref = fn(xr, kr)
        ref.square().sum().backward()
        xg = x.cuda().requires_grad_()
        y = m(xg)
        assert y.shape == ref.shape and float((y.detach().cpu() - ref.detach()).abs().max()) <= 1e-5 * float(ref.abs().max())
        y.square().sum().backward()
        assert float((xg.grad.cpu() - xr.grad).abs().max()) <= 1e-4 * float(xr.grad.abs().max())
        assert float((m.kernel.grad.cpu() - kr.grad).abs().max()) <= 1e-3 * float(kr.grad.abs().max())
    with pytest.raises(RuntimeError):
        hr.Square_to_Hex_Conv2d_by_Double_Stride(2, 4).cuda()(torch.zeros(1, 2, 16, 16).cuda())
    with pytest.raises(RuntimeError):
        hr.Hex_to_Square_original_resolution(2, 0).cuda()(torch.zeros(1, 2, 2, 9).cuda())
