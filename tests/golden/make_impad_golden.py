"""Golden vectors of ``geometry_np.heximpad`` / ``hex_impad_to_multiple`` (geometry_np.py:683-749) from the reference.

    python tests/golden/make_impad_golden.py        # needs /root/reference and cv2 (build container only)

The reference function raises NameError as shipped (it uses ``numbers`` without importing it); that one name is injected
into the module namespace, nothing else is touched.  Output: tests/golden/impad_golden.npz.
"""
import numbers
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402

OUT = os.path.join(HERE, "impad_golden.npz")
DTYPES = (np.uint8, np.float32, np.float64, np.int16, np.uint16, np.int32)


def random_case(rng):
    H, W, C = int(rng.integers(3, 14)), int(rng.integers(3, 14)), int(rng.integers(0, 5))
    dt = DTYPES[int(rng.integers(0, len(DTYPES)))]
    img = (rng.random((H, W) if C == 0 else (H, W, C)) * 200).astype(dt)
    mode = ("constant", "edge", "reflect", "symmetric")[int(rng.integers(0, 4))]
    kind = int(rng.integers(0, 4))
    kw = {"padding_mode": mode}
    if kind == 0:
        kw["shape"] = (H + int(rng.integers(0, 5)), W + int(rng.integers(0, 5)))
    elif kind == 1:
        kw["padding"] = int(rng.integers(0, 3))
    elif kind == 2:
        kw["padding"] = (int(rng.integers(0, 3)), int(rng.integers(0, 3)))
    else:
        kw["padding"] = tuple(int(v) for v in rng.integers(0, 3, 4))
    if mode == "constant":
        r = int(rng.integers(0, 3))
        if r == 1:
            kw["pad_val"] = float(rng.uniform(0, 300))
        elif r == 2 and C > 0:
            kw["pad_val"] = tuple(float(v) for v in rng.uniform(0, 250, C))
    return img, kw


def main():
    gnp = MG.load_reference()[0]
    gnp.numbers = numbers
    rng = np.random.default_rng(20260320)
    out, n = {}, 0
    while n < 48:
        img, kw = random_case(rng)
        try:
            res = gnp.heximpad(img.copy(), **kw)
        except Exception:
            continue
        out[f"pad_{n}_in"] = img
        out[f"pad_{n}_kw"] = np.array(repr(kw))
        out[f"pad_{n}_out"] = res
        n += 1
    out["pad_count"] = np.array(n)
    for k, (shape, div, val) in enumerate([((5, 7, 3), 4, 0), ((9, 6), 3, 9), ((8, 8, 2), 8, 0), ((1, 13, 4), 5, 0)]):
        img = (rng.random(shape) * 200).astype(np.float32)
        out[f"mul_{k}_in"], out[f"mul_{k}_div"], out[f"mul_{k}_val"] = img, np.array(div), np.array(val)
        out[f"mul_{k}_out"] = gnp.hex_impad_to_multiple(img.copy(), div, val)
    out["mul_count"] = np.array(4)
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
