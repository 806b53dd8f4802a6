"""HexConvTranspose2d (retired from the reference into "codes in old versions.txt":129-274).

tests/golden/retired_golden.npz holds outputs of the reference's own class on CPU tensors (the only place it runs: it
builds its canvas with ``torch.zeros(...)`` on the CPU, :193).  CPU: the oracle (the reference's route: canvas, dense
window, two strided convs, interleave) and the product's decomposition (zero-insert table -> stride-1 hex conv -> row /
column selection table, evaluated here with numpy gathers and the hex-conv oracle) both reproduce them to 1e-5.
GPU: the module -- gather kernel, conv kernels, gather kernel through the C ABI -- against fixture and oracle, plus the
backward pass against autograd through the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import hexframes_oracle as HO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "retired_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _cases(G):
    for n in range(int(G["ct_count"])):
        r, st, eo, g, cin, cout, has_bias = (int(v) for v in G[f"ct_{n}_cfg"])
        bias = torch.from_numpy(G[f"ct_{n}_bias"]) if has_bias else None
        yield (r, st, eo, g, cin, cout), torch.from_numpy(G[f"ct_{n}_in"]), torch.from_numpy(G[f"ct_{n}_kernel"]), bias, G[f"ct_{n}_out"]


def _gather(x, tab, hw):
    B, C = x.shape[:2]
    t = torch.from_numpy(np.asarray(tab))
    flat = x.reshape(B, C, -1)
    return torch.where(t >= 0, flat[:, :, t.clamp(min=0)], torch.zeros((), dtype=x.dtype)).reshape(B, C, *hw)


def _decomposed(x, kernel, bias, r, st, eo, g):
    from HyGrid import HexFrames as hf
    up, hu_wu, o_u, sel, ho_wo, hy_wy = hf.conv_transpose_tables(r, st, eo, x.shape[2], x.shape[3])
    Y = HO.hexconv2d(_gather(x, up, hu_wu), kernel, bias, o_u, r, 1, 0, 1, g)
    assert tuple(Y.shape[-2:]) == hy_wy
    return _gather(Y, sel, ho_wo)


def test_oracle_and_decomposition_reproduce_the_reference(golden):
    assert int(golden["ct_count"]) >= 8
    for (r, st, eo, g, cin, cout), x, kernel, bias, want in _cases(golden):
        got = HO.hex_conv_transpose2d(x, kernel, bias, eo, r, st, g).numpy()
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-5
        dec = _decomposed(x, kernel, bias, r, st, eo, g).numpy()
        assert dec.shape == want.shape and np.abs(dec - want).max() <= 1e-5


def test_decomposition_on_more_shapes_and_where_the_reference_raises():
    from HyGrid import HexFrames as hf
    torch.manual_seed(5)
    agree = 0
    for st in (1, 2, 3, 4):
        for r in (2, 3):
            for eo in (0, 1):
                for H, W in [(6, 7), (5, 5), (4, 9), (8, 6), (7, 4), (2, 2)]:
                    x, kernel, bias = torch.randn(1, 2, H, W), torch.randn(3, 2, 1, 3 * r * r - 3 * r + 1), torch.randn(3)
                    try:
                        want = HO.hex_conv_transpose2d(x, kernel, bias, eo, r, st)
                    except ValueError:
                        with pytest.raises(ValueError):
                            hf.conv_transpose_tables(r, st, eo, H, W)
                        continue
                    got = _decomposed(x, kernel, bias, r, st, eo, 1)
                    assert got.shape == want.shape and float((got - want).abs().max()) <= 1e-5
                    agree += 1
    assert agree >= 50
    # stride 1 is HexConv2d with padding r-1 (old versions :186-205 degenerate to pad + type1)
    x, kernel = torch.randn(2, 3, 7, 6), torch.randn(4, 3, 1, 7)
    assert torch.allclose(HO.hex_conv_transpose2d(x, kernel, None, 1, 2, 1), HO.hexconv2d(x, kernel, None, 1, 2, 1, 1), atol=1e-5)


def test_module_contract_without_a_gpu():
    from HyGrid import HexFrames as hf
    m = hf.HexConvTranspose2d(4, 6, 1, 3, stride=2, groups=2, bias=True)
    for name in ("in_channel", "out_channel", "even_odd_offset", "hexkernel_radius", "hexkernel_size", "k_w", "k_h",
                 "kernelnum", "sh", "sw", "out_even_odd_offset", "groups", "b"):
        assert hasattr(m, name), name
    assert (m.k_h, m.k_w, m.kernelnum, m.sh, m.sw) == (5, 9, 19, 2, 4)
    assert tuple(m.kernel.shape) == (6, 2, 1, 19) and set(m.state_dict()) == {"kernel", "bias"}
    assert set(hf.HexConvTranspose2d(4, 4, 0, 2).state_dict()) == {"kernel"}           # bias=False by default (:131)
    with pytest.raises(ValueError):
        hf.HexConvTranspose2d(3, 4, 0, 2, groups=2)
    assert "HexConvTranspose2d" in repr(m)


@pytest.mark.gpu
def test_gpu_module_matches_fixture_and_oracle(golden):
    from HyGrid import HexFrames as hf
    for (r, st, eo, g, cin, cout), x, kernel, bias, want in _cases(golden):
        m = hf.HexConvTranspose2d(cin, cout, eo, r, stride=st, groups=g, bias=bias is not None).cuda()
        with torch.no_grad():
            m.kernel.copy_(kernel)
            if bias is not None:
                m.bias.copy_(bias)
        xg = x.cuda().requires_grad_(True)
        y = m(xg)
        assert y.dtype == torch.float32 and tuple(y.shape) == want.shape
        scale = max(1.0, float(np.abs(want).max()))
        assert float((y.detach().cpu() - torch.from_numpy(want)).abs().max()) <= 1e-4 * scale
        # backward: autograd through the oracle's canvas + conv2d route is the checker
        gy = torch.randn_like(y)
        y.backward(gy)
        xo, ko = x.clone().requires_grad_(True), kernel.clone().requires_grad_(True)
        bo = bias.clone().requires_grad_(True) if bias is not None else None
        HO.hex_conv_transpose2d(xo, ko, bo, eo, r, st, g).backward(gy.cpu())
        # tolerances of tests/test_gpu_hexframes.py: data gradient 1e-4, weight / bias gradient (float atomics) 1e-3 of the range
        for got, ref, rel in ((xg.grad, xo.grad, 1e-4), (m.kernel.grad, ko.grad, 1e-3)) + (((m.bias.grad, bo.grad, 1e-3),) if bias is not None else ()):
            assert float((got.cpu() - ref).abs().max()) <= rel * max(1.0, float(ref.abs().max()))


def test_selection_table_is_split_into_injective_layers_for_the_adjoint():
    """For an even stride both output parities read even rows of the stride-1 result, so some elements are gathered
    twice; the scatter kernel does plain stores, hence the backward sums one scatter per injective layer."""
    from HyGrid import HexFrames as hf
    for st, want_layers in ((1, 1), (2, 2), (3, 1)):
        up, _, _, sel, _, _ = hf.conv_transpose_tables(2, st, 0, 6, 7)
        assert len(hf._injective_layers(up)) == 1                     # zero insertion reads every input cell once
        layers = hf._injective_layers(sel)
        assert len(layers) == want_layers
        total = np.zeros_like(sel)
        for t in layers:
            v = t[t >= 0]
            assert len(np.unique(v)) == len(v)
            total += (t >= 0)
            assert np.array_equal(t[t >= 0], sel[t >= 0])
        assert np.array_equal(total, (sel >= 0).astype(total.dtype))  # every entry lands in exactly one layer
    t = np.array([3, -1, 5, 3, 0, 5, 3, 7])
    assert [l.tolist() for l in hf._injective_layers(t)] == [[3, -1, 5, -1, 0, -1, -1, 7], [-1, -1, -1, 3, -1, 5, -1, -1],
                                                            [-1, -1, -1, -1, -1, -1, 3, -1]]
