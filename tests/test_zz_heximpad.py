"""``geometry_np.heximpad`` / ``hex_impad_to_multiple`` (geometry_np.py:683-749).

tests/golden/impad_golden.npz: outputs of the reference's own function (``numbers`` injected -- it is the one name the
reference forgets to import), 48 random cases over dtypes, bands, modes, padding forms and border values, plus
``hex_impad_to_multiple``.  CPU: the oracle reproduces them exactly.  GPU: the product (``hg_pad2d`` through the C ABI)
returns exactly the same arrays -- shapes, dtypes and OpenCV's quirks (a scalar border value reaches band 0 only, a
single-band axis is dropped, the top pad is rounded down to an even number of rows)."""
import os

import numpy as np
import pytest

from oracle import hygrid_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "impad_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _cases(G):
    for n in range(int(G["pad_count"])):
        yield G[f"pad_{n}_in"], eval(str(G[f"pad_{n}_kw"]), {"__builtins__": {}}), G[f"pad_{n}_out"]   # repr of a dict of ints / floats / tuples / str


def _same(a, b):
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    assert np.array_equal(a, b)


def test_oracle_reproduces_the_reference(golden):
    assert int(golden["pad_count"]) == 48
    for img, kw, want in _cases(golden):
        _same(O.heximpad(img, **kw), want)
    for k in range(int(golden["mul_count"])):
        _same(O.hex_impad_to_multiple(golden[f"mul_{k}_in"], int(golden[f"mul_{k}_div"]), int(golden[f"mul_{k}_val"])), golden[f"mul_{k}_out"])


def test_documented_quirks():
    img = np.ones((2, 2, 3))
    out = O.heximpad(img, padding=(1, 0, 0, 0), pad_val=7.5)
    assert out[:, 0].tolist() == [[7.5, 0.0, 0.0]] * 2                      # Scalar(7.5) = (7.5, 0, 0, 0)
    assert O.heximpad(np.ones((2, 2, 3), np.uint8), padding=(1, 0, 0, 0), pad_val=7.5)[0, 0, 0] == 8     # saturate_cast rounds
    assert O.heximpad(np.ones((3, 3, 1)), padding=1).shape == (5, 5)        # single-band axis dropped
    assert O.heximpad(np.ones((3, 3)), padding=(0, 3, 0, 0)).shape == (6, 3)
    assert O.heximpad(np.arange(9.0).reshape(3, 3), padding=(0, 3, 0, 0))[:2].sum() == 0   # top 3 -> 2 rows, 1 moved to the bottom
    with pytest.raises(Exception):
        O.heximpad(np.ones((2, 2, 5)), padding=1, pad_val=1.0)


@pytest.mark.gpu
def test_gpu_heximpad_returns_the_reference_arrays(golden):
    from HyGrid import geometry_np as gnp
    for img, kw, want in _cases(golden):
        _same(gnp.heximpad(img, **kw), want)
    for k in range(int(golden["mul_count"])):
        _same(gnp.hex_impad_to_multiple(golden[f"mul_{k}_in"], int(golden[f"mul_{k}_div"]), int(golden[f"mul_{k}_val"])), golden[f"mul_{k}_out"])
    img = np.ones((2, 2, 3))
    assert gnp.heximpad(img, padding=(1, 0, 0, 0), pad_val=7.5)[:, 0].tolist() == [[7.5, 0.0, 0.0]] * 2
    with pytest.raises(Exception):
        gnp.heximpad(np.ones((2, 2, 5)), padding=1, pad_val=1.0)
    with pytest.raises(TypeError):
        gnp.heximpad(img, padding=1, pad_val="0")
    with pytest.raises(ValueError):
        gnp.heximpad(img, padding=(1, 2, 3))
