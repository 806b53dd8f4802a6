"""Generate golden fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py          # needs /root/reference (build container only)

The reference ships no tests, so these fixtures -- outputs of
the reference's own functions on seeded inputs -- are what pins the oracle
(``oracle/``) and, through it, the CUDA kernels.  The GPU box has no
``/root/reference``; only the ``.npz`` files written here travel.

How the reference is made importable without touching it:
  * ``HyGrid.Image`` / ``HyGrid.HexImage`` ``sys.exit()`` when GDAL/mmcv/OpenGL/...
    are missing -> empty stub modules are registered in ``sys.modules`` first.
  * ``HyGrid.geometry_torch`` hard-codes ``device='cuda'``; its source text is
    exec'd with the literal ``'cuda'`` replaced by ``'cpu'`` (no other change).
  * ``HexAdaptivePool2d`` / ``HexGlobalPool2d`` reference an undefined global
    ``centroid_pooling``; that one name is injected into the module namespace.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mod("matplotlib"); mod("matplotlib.pyplot")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    mod("osgeo", gdal=types.SimpleNamespace())
    mod("mmcv")
    gl = mod("OpenGL.GL", GL_TEXTURE_2D=0, GL_RGB=0, GL_UNSIGNED_BYTE=0, shaders=types.SimpleNamespace())
    mod("OpenGL", GL=gl)
    mod("OpenGL.arrays"); mod("OpenGL.arrays.vbo", VBO=object)
    mod("glfw"); mod("tkinter")
    pil = mod("PIL"); mod("PIL.Image"); pil.Image = sys.modules["PIL.Image"]


def load_reference():
    _stub_modules()
    sys.path.insert(0, REF)
    import HyGrid.geometry_np as gnp
    import HyGrid.HexFrames as hf
    hf.centroid_pooling = lambda v: (_ for _ in ()).throw(NotImplementedError())
    from HyGrid.Image import IMAGE
    from HyGrid.HexImage import HEXIMAGE
    src = open(os.path.join(REF, "HyGrid", "geometry_torch.py")).read().replace("'cuda'", "'cpu'")
    gt = types.ModuleType("ref_geometry_torch_cpu")
    exec(compile(src, "geometry_torch.py[cuda->cpu]", "exec"), gt.__dict__)
    return gnp, gt, hf, IMAGE, HEXIMAGE


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def rand_img(rng, shape, dtype):
    if dtype == "u8":
        return rng.integers(0, 256, shape, dtype=np.uint8)
    a = rng.random(shape) * 255.0
    return a.astype(np.float32 if dtype == "f32" else np.float64)


def index_image(c_unused, h, w):
    """channels hold the exact row / column index: pushed through a *nearest*
    resampler it reveals the integer gather table of the reference."""
    ii, jj = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    return np.stack([ii + 1, jj + 1], 0).astype(np.float64)   # +1 so that zero-fill is visible


def main():
    gnp, gt, hf, IMAGE, HEXIMAGE = load_reference()
    rng = np.random.default_rng(20261018)
    G = {}

    # ---------------- R1 rect -> hex --------------------------------------
    r1_cases = [((3, 12, 10), None), ((3, 12, 10), (6, 5)), ((3, 17, 23), (9, 31)),
                ((2, 2, 2), (2, 2)), ((1, 9, 14), (20, 7)), ((3, 33, 65), (33, 65)),
                ((3, 64, 48), (32, 24))]
    n = 0
    for shape, dsize in r1_cases:
        for dt in ("u8", "f32", "f64"):
            img = rand_img(rng, shape, dt)
            for interp in ("nearest", "bilinear"):
                out = gnp.rect_to_hex_resample(img, dsize, interp)
                G[f"r1_{n}_img"] = img
                G[f"r1_{n}_dsize"] = np.array(dsize if dsize else (-1, -1))
                G[f"r1_{n}_interp"] = np.array(interp)
                G[f"r1_{n}_out"] = out
                n += 1
    G["r1_count"] = np.array(n)
    # integer gather tables via the index image
    n = 0
    for (h, w), dsize in [((12, 10), (12, 10)), ((17, 23), (9, 31)), ((64, 48), (32, 24)),
                          ((1080 // 8, 1920 // 8), (2160 // 8, 3840 // 8)), ((512, 512), (256, 256)),
                          ((1024, 1024), (1024, 1024))]:
        out = gnp.rect_to_hex_resample(index_image(2, h, w), dsize, "nearest")
        G[f"r1idx_{n}_hw"] = np.array([h, w, dsize[0], dsize[1]])
        # 1-D tables read off a valid middle column / row (index+1; 0 = zero-filled)
        G[f"r1idx_{n}_i"] = out[0].astype(np.int32)[:, out.shape[2] // 2]
        G[f"r1idx_{n}_j"] = out[1].astype(np.int32)[out.shape[1] // 2, :]
        G[f"r1idx_{n}_sha"] = np.array(sha(out.astype(np.int32)))
        n += 1
    G["r1idx_count"] = np.array(n)
    # config-1 sized case, hashed only (C1: 512x512 RGB u8 -> hex 256x256 nearest via IMAGE)
    img = rand_img(np.random.default_rng(0), (3, 512, 512), "u8")
    hexd = IMAGE(data=img).ConvertToHexagon()
    G["c1_hex_sha"] = np.array(sha(hexd)); G["c1_hex_shape"] = np.array(hexd.shape)
    G["c1_hex_dtype"] = np.array(str(hexd.dtype))
    back = gnp.hex_to_rect_resample(HEXIMAGE(data=hexd).HexagonImage, (512, 512), "linear")
    G["c1_back_sha"] = np.array(sha(back)); G["c1_back_sum"] = np.array(back.sum())
    G["c1_back_probe"] = back[:, ::37, ::41].copy()

    # ---------------- R2 / R4 hex -> rect, hexresize -----------------------
    n = 0
    r2_cases = [((3, 12, 10), None), ((3, 12, 10), (24, 20)), ((3, 17, 23), (9, 31)),
                ((1, 9, 14), (20, 7)), ((3, 32, 24), (64, 48)), ((2, 5, 4), (3, 11))]
    for shape, dsize in r2_cases:
        for dt in ("u8", "f32", "f64"):
            img = rand_img(rng, shape, dt)
            G[f"r2_{n}_img"] = img
            G[f"r2_{n}_dsize"] = np.array(dsize if dsize else (-1, -1))
            G[f"r2_{n}_np_linear"] = gnp.hex_to_rect_resample(img, dsize, "linear")
            G[f"r2_{n}_torch_linear"] = gt.hex_to_square_resample(img, dsize, "linear")
            G[f"r2_{n}_torch_nearest"] = gt.hex_to_square_resample(img, dsize, "nearest")
            ds = dsize if dsize else shape[1:]
            G[f"r2_{n}_resize_linear"] = gnp.hexresize(img, ds, "linear")
            n += 1
    G["r2_count"] = np.array(n)
    # index tables for the fp32-linspace twin vs the fp64 twin (SURVEY: they differ)
    n = 0
    for (h, w), dsize in [((12, 10), (24, 20)), ((135, 240), (270, 480)), ((256, 256), (512, 512))]:
        out = gt.hex_to_square_resample(index_image(2, h, w), dsize, "nearest")
        G[f"r2idx_{n}_hw"] = np.array([h, w, dsize[0], dsize[1]])
        G[f"r2idx_{n}_sha"] = np.array(sha(out.astype(np.int32)))
        G[f"r2idx_{n}_probe"] = out.astype(np.int32)[:, ::7, ::5].copy()
        n += 1
    G["r2idx_count"] = np.array(n)

    # ---------------- R3 warp ----------------------------------------------
    th = np.pi / 6
    Hs = [np.eye(3),
          np.array([[1.7, 0, 0], [0, 1.3, 0], [0, 0, 1.0]]),
          np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]]),
          np.array([[0.8, 0.3, 1.5], [-0.2, 1.1, -2.25], [0, 0, 1.0]])]
    n = 0
    for shape in [(3, 12, 10), (2, 9, 14)]:
        for dt in ("u8", "f32", "f64"):
            img = rand_img(rng, shape, dt)
            for Hm in Hs:
                G[f"r3_{n}_img"] = img
                G[f"r3_{n}_H"] = Hm
                G[f"r3_{n}_np_linear"] = gnp.image_geometric_transformation(img, Hm, "linear")
                G[f"r3_{n}_torch_linear"] = gt.image_geometric_transformation_gpu(img, Hm, "linear")
                G[f"r3_{n}_torch_nearest"] = gt.image_geometric_transformation_gpu(img, Hm, "nearest")
                n += 1
    G["r3_count"] = np.array(n)

    # ---------------- R5 doubled rasters -----------------------------------
    n = 0
    for shape in [(3, 6, 5), (1, 7, 4), (2, 1, 3)]:
        for off in (0, 1):
            img = rand_img(rng, shape, "f64")
            hx = HEXIMAGE(data=img, even_odd_offset=off)
            t1, g1 = hx.GenerateType1Image()
            t2, g2 = hx.GenerateType2Image()
            G[f"r5_{n}_img"] = img; G[f"r5_{n}_off"] = np.array(off)
            G[f"r5_{n}_t1"] = t1; G[f"r5_{n}_t2"] = t2
            G[f"r5_{n}_g1"] = np.array(g1); G[f"r5_{n}_g2"] = np.array(g2)
            G[f"r5_{n}_dec1"] = HEXIMAGE(data=t1, heximagetype=1).HexagonImage
            G[f"r5_{n}_dec2"] = HEXIMAGE(data=t2, heximagetype=2).HexagonImage
            x = torch.tensor(img, dtype=torch.float32)[None]
            G[f"r5_{n}_tt1"] = hf.heximage_to_type1(x, off).numpy()
            G[f"r5_{n}_tt2"] = hf.heximage_to_type2(x, off).numpy()
            G[f"r5_{n}_tdec"] = hf.type1_to_heximage(hf.heximage_to_type1(x, off), off)[0].numpy()
            n += 1
    G["r5_count"] = np.array(n)
    np.savez_compressed(os.path.join(OUT, "resample_golden.npz"), **G)

    # ---------------- C1/C2 hex conv ---------------------------------------
    torch.manual_seed(7)
    C = {}
    n = 0
    conv_cases = []
    for r in (2, 3):
        for s in (1, 2):
            for d in (1, 2):
                for pad in (0, 1, 2):
                    for off in (0, 1):
                        conv_cases.append(dict(r=r, s=s, d=d, pad=pad, off=off, g=1, bias=True, H=13, W=11, Cin=3, Cout=4))
    conv_cases += [dict(r=2, s=1, d=1, pad=1, off=0, g=2, bias=True, H=10, W=12, Cin=4, Cout=6),
                   dict(r=2, s=1, d=1, pad=1, off=1, g=4, bias=False, H=9, W=9, Cin=4, Cout=4),
                   dict(r=2, s=2, d=1, pad=1, off=0, g=1, bias=True, H=16, W=16, Cin=5, Cout=2),
                   dict(r=4, s=1, d=1, pad=3, off=0, g=1, bias=True, H=14, W=15, Cin=2, Cout=3)]
    for cs in conv_cases:
        m = hf.HexConv2d(cs["Cin"], cs["Cout"], cs["off"], cs["r"], stride=cs["s"], padding=cs["pad"],
                         dilation=cs["d"], groups=cs["g"], bias=cs["bias"])
        x = torch.randn(2, cs["Cin"], cs["H"], cs["W"], requires_grad=True)
        try:
            y = m(x)
        except Exception as e:          # shapes the reference itself cannot interleave
            continue
        gy = torch.randn_like(y)
        (y * gy).sum().backward()
        C[f"conv_{n}_cfg"] = np.array([cs[k] for k in ("r", "s", "d", "pad", "off", "g")] + [int(cs["bias"])])
        C[f"conv_{n}_x"] = x.detach().numpy(); C[f"conv_{n}_w"] = m.kernel.detach().numpy()
        if cs["bias"]:
            C[f"conv_{n}_b"] = m.bias.detach().numpy(); C[f"conv_{n}_db"] = m.bias.grad.numpy()
        C[f"conv_{n}_y"] = y.detach().numpy(); C[f"conv_{n}_gy"] = gy.numpy()
        C[f"conv_{n}_dx"] = x.grad.numpy(); C[f"conv_{n}_dw"] = m.kernel.grad.numpy()
        n += 1
    C["conv_count"] = np.array(n)
    n = 0
    for (H, W, r, s, d) in [(8, 8, 2, 1, 1), (9, 7, 2, 2, 1), (11, 10, 3, 1, 1), (12, 9, 2, 1, 2)]:
        m = hf.HexConv2dAdaptivePadding(3, 2, 0, r, stride=s, dilation=d)
        x = torch.randn(1, 3, H, W)
        try:
            y = m(x)
        except Exception:
            continue
        C[f"aconv_{n}_cfg"] = np.array([r, s, d]); C[f"aconv_{n}_x"] = x.numpy()
        C[f"aconv_{n}_w"] = m.kernel.detach().numpy(); C[f"aconv_{n}_b"] = m.bias.detach().numpy()
        C[f"aconv_{n}_y"] = y.detach().numpy()
        n += 1
    C["aconv_count"] = np.array(n)

    # ---------------- P1/P2 pooling ----------------------------------------
    n = 0
    pool_cases = []
    for method in ("max", "min", "average"):
        pool_cases += [dict(method=method, k=2, s=2, pad=0, ceil=False, cip=True, H=12, W=13, nan=False),
                       dict(method=method, k=2, s=2, pad=1, ceil=False, cip=True, H=9, W=10, nan=False),
                       dict(method=method, k=2, s=2, pad=0, ceil=True, cip=True, H=11, W=11, nan=False),
                       dict(method=method, k=2, s=2, pad=0, ceil=True, cip=False, H=11, W=11, nan=False),
                       dict(method=method, k=2, s=4, pad=0, ceil=False, cip=True, H=16, W=18, nan=False),
                       dict(method=method, k=3, s=3, pad=0, ceil=False, cip=True, H=12, W=14, nan=False),
                       dict(method=method, k=2, s=2, pad=0, ceil=False, cip=True, H=8, W=9, nan=True),
                       dict(method=method, k=(2, 3), s=(2, 4), pad=0, ceil=False, cip=True, H=10, W=15, nan=False)]
    for cs in pool_cases:
        m = hf.HexPool2d(cs["method"], cs["k"], cs["s"], padding=cs["pad"], ceil_mode=cs["ceil"],
                         count_include_pad=cs["cip"])
        x = torch.randn(2, 3, cs["H"], cs["W"])
        if cs["nan"]:
            x[torch.rand_like(x) < 0.3] = float("nan")
            x[0, 0, 0:2, 0:2] = float("nan")      # one all-NaN window
        x.requires_grad_(True)
        try:
            y = m(x)
        except Exception:
            continue
        gy = torch.randn_like(y)
        (torch.nan_to_num(y) * gy).sum().backward()
        k = cs["k"] if isinstance(cs["k"], tuple) else (cs["k"], cs["k"])
        s = cs["s"] if isinstance(cs["s"], tuple) else (cs["s"], cs["s"])
        C[f"pool_{n}_cfg"] = np.array([k[0], k[1], s[0], s[1], cs["pad"], int(cs["ceil"]), int(cs["cip"])])
        C[f"pool_{n}_method"] = np.array(cs["method"])
        C[f"pool_{n}_x"] = x.detach().numpy(); C[f"pool_{n}_y"] = y.detach().numpy()
        C[f"pool_{n}_gy"] = gy.numpy(); C[f"pool_{n}_dx"] = x.grad.numpy()
        n += 1
    C["pool_count"] = np.array(n)
    n = 0
    for method in ("max", "min", "average"):
        for outsize in (1, 2, 3, 4, 6):
            x = torch.randn(2, 3, 12, 13)
            m = hf.HexAdaptivePool2d(outsize, method)
            C[f"apool_{n}_method"] = np.array(method); C[f"apool_{n}_out"] = np.array(outsize)
            C[f"apool_{n}_x"] = x.numpy(); C[f"apool_{n}_y"] = m(x).numpy()
            n += 1
        x = torch.randn(2, 3, 7, 5)
        C[f"gpool_{method}_x"] = x.numpy(); C[f"gpool_{method}_y"] = hf.HexGlobalPool2d(method)(x).numpy()
    C["apool_count"] = np.array(n)
    np.savez_compressed(os.path.join(OUT, "hexframes_golden.npz"), **C)
    for f in ("resample_golden.npz", "hexframes_golden.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
