"""CPU oracle for the hex-lattice operators of HyGrid/HexFrames.py (torch CPU).

TEST INFRASTRUCTURE ONLY -- see the header of ``hygrid_oracle.py`` for who may
import this.  These are floating-point kernels, so the oracle is a plain
torch-CPU fp32/fp64 restatement (autograd supplies the backward oracle).

Parity pin: the reference has no tests; ``tests/golden/make_golden.py`` runs the
reference's own ``HexConv2d`` / ``HexPool2d`` / ``heximage_to_type1`` ... in the
build container and commits inputs+outputs; ``tests/test_oracle_golden.py``
checks the restatement below against them.

The reference evaluates the hex convolution by materialising a 2x-wide
"doubled" image and running two dense strided ``F.conv2d`` with a zero-stuffed
(2r-1)x(4r-3) kernel (HexFrames.py:96-169).  The restatement here is the closed
form in *offset* coordinates: one gather per hexagonal tap, no doubled image.

All ``file:line`` citations are into ``/root/reference/HyGrid/``.
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F

__all__ = [
    "hex_taps", "hexconv_out_shape", "hexconv2d", "adaptive_padding",
    "hexpool_out_shape", "hexpool2d", "hexadaptivepool2d", "hexglobalpool2d",
    "reduce_max", "reduce_min", "reduce_average",
    "heximage_to_type1", "heximage_to_type2", "type1_to_heximage", "hex_pixel_shuffle",
    "hex_conv_transpose2d", "resampler_weight", "hex_to_square_double_stride", "square_to_hex_double_stride",
    "hex_to_square_original_resolution",
]


# --------------------------------------------------------------------------
# hex convolution (HexFrames.py:22-185)
# --------------------------------------------------------------------------
def hex_taps(radius: int):
    """[(a, t, m)] for every weight index, in the reference's running-sum order
    (HexFrames.py:112-118): kernel row a, |a-(r-1)| = t, m-th cell of that row."""
    taps = []
    for a in range(2 * radius - 1):
        t = abs(a - radius + 1)
        for m in range(2 * radius - 1 - t):
            taps.append((a, t, m))
    return taps


def hexconv_out_shape(Hp, Wp, radius, stride, dilation):
    """Rows / cols of the interleaved output for a padded Hp x Wp input
    (HexFrames.py:127-162 with k_h, k_w from :82-83).  Returns (rows_even,
    rows_odd, cols); a count <= 0 means that conv is skipped (``None``)."""
    s, d = stride, dilation
    k_h = (2 * radius - 2) * d + 1
    k_w = 2 * d * (2 * radius - 2) + 1
    wt = 2 * Wp - s                       # width of both sliced doubled images
    cols = (wt - k_w) // (2 * s) + 1 if wt >= k_w else 0
    rows_e = (Hp - k_h) // (2 * s) + 1 if Hp >= k_h else 0
    rows_o = (Hp - s - k_h) // (2 * s) + 1 if Hp - s >= k_h else 0
    return rows_e, rows_o, cols


def hexconv2d(x, kernel, bias=None, even_odd_offset=0, radius=2, stride=1, padding=0,
              dilation=1, groups=1, padding_mode="constant", padding_value=0):
    """Closed form of HexConv2d.forward.

    With P = F.pad(x), o = (even_odd_offset + padding) % 2, output row R, col q:
        y[R, q] = bias + sum_taps w * D[s*R + a*d, 1 + (R%2)*s + 2*s*q + t*d + 2*d*m]
    where D is the doubled image: D[i, c] = P[i, (c - s_i) // 2] for
    s_i <= c < 2*Wp + s_i (s_i = (i + o) % 2) and 0 elsewhere
    (HexFrames.py:417-445 for D; :129-144 for the two strided convs whose
    interleave (:157-162) gives the single formula in R)."""
    while x.dim() < 4:
        x = x.unsqueeze(0)
    x = x.to(kernel.dtype)
    s, d, r = stride, dilation, radius
    P = F.pad(x, (padding,) * 4, padding_mode, padding_value)
    N, Cin, Hp, Wp = P.shape
    Cout = kernel.shape[0]
    o = (even_odd_offset + padding) % 2
    rows_e, rows_o, cols = hexconv_out_shape(Hp, Wp, r, s, d)
    if rows_e <= 0 or rows_o < 0 or cols <= 0:
        raise ValueError("input too small: the reference takes an undefined path here")
    # rows_o == 0 (padded height in [k_h, k_h + s)): the odd conv is skipped and the reference returns the even-row
    # result alone (HexFrames.py:163-164) -- one output row, which the formula below yields with Ho = 1
    if rows_e - rows_o not in (0, 1):
        raise ValueError("even/odd row counts cannot be interleaved (the reference raises too)")
    Ho = rows_e + rows_o
    R = torch.arange(Ho).view(Ho, 1)
    q = torch.arange(cols).view(1, cols)
    cin_g, cout_g = Cin // groups, Cout // groups
    y = torch.zeros(N, Cout, Ho, cols, dtype=kernel.dtype)
    for k, (a, t, m) in enumerate(hex_taps(r)):
        i = s * R + a * d                                   # (Ho,1) row in P
        c = 1 + (R % 2) * s + 2 * s * q + t * d + 2 * d * m  # (Ho,cols) doubled column
        si = (i + o) % 2
        ok = (c >= si) & (c < 2 * Wp + si)
        pc = torch.div(c - si, 2, rounding_mode="floor").clamp(0, Wp - 1)
        g = P[:, :, i.expand(Ho, cols), pc] * ok.to(P.dtype)  # (N,Cin,Ho,cols)
        wk = kernel[:, :, 0, k]                               # (Cout, Cin/g)
        for gi in range(groups):
            y[:, gi * cout_g:(gi + 1) * cout_g] += torch.einsum(
                "oc,nchw->nohw", wk[gi * cout_g:(gi + 1) * cout_g], g[:, gi * cin_g:(gi + 1) * cin_g])
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y


def adaptive_padding(Hin, Win, radius, stride, dilation):
    """(left, right, top, bottom) of HexConv2dAdaptivePadding (HexFrames.py:235-250)."""
    ks = 2 * radius - 1
    oh, ow = math.ceil(Hin / stride), math.ceil(Win / stride)
    pad_h = max((oh - 1) * stride + (ks - 1) * dilation + 1 - Hin, 0)
    pad_w = max(ow * stride + (ks - 1) * dilation + 1 - Win, 0)
    return (pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2)


# --------------------------------------------------------------------------
# pooling (HexFrames.py:255-414, 461-479)
# --------------------------------------------------------------------------
def reduce_max(v):   # HexFrames.py:461-463
    return torch.max(v.masked_fill(v.isnan(), float("-inf")), dim=-1)[0]


def reduce_min(v):   # HexFrames.py:464-466
    return torch.min(v.masked_fill(v.isnan(), float("inf")), dim=-1)[0]


def reduce_average(v):  # HexFrames.py:467-479
    nan = v.isnan()
    cnt = (~nan).to(v.dtype).sum(-1)
    tot = torch.where(nan, torch.zeros_like(v), v).sum(-1)
    out = tot / cnt
    return torch.where(cnt == 0, torch.full_like(out, float("nan")), out)


_REDUCE = {"max": reduce_max, "min": reduce_min, "average": reduce_average}


def hexpool_out_shape(h, w, kh, kw, sh, sw):
    """HexFrames.py:302-303."""
    return (h - kh) // sh + 1, (w - sw // 2) // sw


def _windows(P, hn, wn, kh, kw, sh, sw, shift):
    """(B,C,hn,wn,kh*kw) windows: top-left of out(I,J) is
    (sh*I, ((I%2)*shift)//2 + J*sw)  (HexFrames.py:318-325 / 385-392)."""
    B, C, h, w = P.shape
    I = torch.arange(hn).view(hn, 1, 1, 1)
    J = torch.arange(wn).view(1, wn, 1, 1)
    a = torch.arange(kh).view(1, 1, kh, 1)
    b = torch.arange(kw).view(1, 1, 1, kw)
    ii = (sh * I + a).expand(hn, wn, kh, kw)
    jj = (torch.div((I % 2) * shift, 2, rounding_mode="floor") + J * sw + b).expand(hn, wn, kh, kw)
    if hn > 0 and wn > 0 and (int(ii.max()) >= h or int(jj.max()) >= w):
        raise IndexError("pool window leaves the image (the reference raises IndexError here)")
    return P[:, :, ii, jj].reshape(B, C, hn, wn, kh * kw)


def hexpool2d(x, method, kernel_size=2, stride=None, padding=0, padding_mode="constant",
              padding_value=0, ceil_mode=False, count_include_pad=True):
    """HexPool2d.forward (HexFrames.py:286-336).  ``stride=None`` means
    stride = kernel_size (the evident intent of :273-274; the reference crashes)."""
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    if stride is None:
        stride = (kh, kw)
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    while x.dim() < 4:
        x = x.unsqueeze(0)
    P = F.pad(x, (padding,) * 4, padding_mode, padding_value)
    if ceil_mode:
        h, w = P.shape[-2:]
        hn0 = h // sh
        wn0 = (w - sw // 2 - sw) // sw + 1
        ph = (kh - h + hn0 * sh) % kh
        pw = (kw - w + (wn0 * sw + sw // 2)) % kw
        # literal argument order of HexFrames.py:298 -- (left,right,top,bottom)=(0,ph,0,pw)
        P = F.pad(P, (0, ph, 0, pw), mode="constant",
                  value=0 if count_include_pad else float("nan"))
    h, w = P.shape[-2:]
    hn, wn = hexpool_out_shape(h, w, kh, kw, sh, sw)
    return _REDUCE[method](_windows(P, hn, wn, kh, kw, sh, sw, sw))


def hexadaptivepool2d(x, outsize: int, method):
    """HexAdaptivePool2d.forward (HexFrames.py:362-401)."""
    while x.dim() < 4:
        x = x.unsqueeze(0)
    h, w = x.shape[-2:]
    hn = wn = outsize
    gh = int(h / hn)
    gw = int(w / (wn + 0.5)) if gh > 1 else int(w / wn)
    return _REDUCE[method](_windows(x, hn, wn, gh, gw, gh, gw, gw))


def hexglobalpool2d(x, method):
    """HexGlobalPool2d.forward (HexFrames.py:410-414): returns (B, C)."""
    while x.dim() < 4:
        x = x.unsqueeze(0)
    return _REDUCE[method](x.reshape(x.size(0), x.size(1), -1))


# --------------------------------------------------------------------------
# layout converters (HexFrames.py:417-458)
# --------------------------------------------------------------------------
def heximage_to_type1(x, even_odd_offset):
    while x.dim() < 4:
        x = x.unsqueeze(0)
    B, C, H, W = x.shape
    out = torch.zeros(B, C, H, 2 * W + 1, dtype=torch.float32)
    rep = x.repeat_interleave(2, dim=3).to(torch.float32)
    for i in range(H):
        s = (i + even_odd_offset) % 2
        out[:, :, i, s:s + 2 * W] = rep[:, :, i]
    return out


def heximage_to_type2(x, even_odd_offset):
    return heximage_to_type1(x, even_odd_offset).repeat_interleave(2, dim=2)


def type1_to_heximage(t1, even_odd_offset):
    return t1[:, :, :, 1::2], even_odd_offset


# --------------------------------------------------------------------------
# hex pixel shuffle (retired; "codes in old versions.txt":68-126)
# --------------------------------------------------------------------------
def hex_pixel_shuffle(x, upscale_factor):
    """(B, C*r*r, H, W) -> (B, C, r*H-r+1, r*W-ceil(r/2)) float32, written the way the reference does it: every
    sub-channel block is painted into a doubled (type1) canvas, cell by cell, in the reference's order (later writes
    overwrite earlier ones), then every second canvas column is kept and the frame is cropped.

    :84-87 canvas of r*H+r-1 rows and 2*(r*W+r//2)+1 sub-columns; :90-93 the second sub-column of a cell lies at +1
    (r even) or -1 (r odd); :101-104 sub-channel blocks enumerate hexagon rows i (r-t cells, t=|1+i-r|); :105-112
    even input rows -> canvas rows i+2r*a, sub-columns 1+t+2k+2r*b; :114-122 odd input rows r rows lower and r
    sub-columns further right; :125 type1 -> hex keeps sub-columns 1::2; :126 crop."""
    r = int(upscale_factor)
    while x.dim() < 4:
        x = x.unsqueeze(0)
    B, C, H, W = x.shape
    if C % (r * r) != 0:
        raise Exception("channel count must be a multiple of upscale_factor**2")   # :81-82
    cout = C // (r * r)
    rows_c, cells_c = r * H + r - 1, r * W + r // 2
    canvas = torch.zeros(B, cout, rows_c, 2 * cells_c + 1, dtype=torch.float32)
    pair = 1 if r % 2 == 0 else -1
    n = 0
    for i in range(2 * r - 1):
        t = abs(1 + i - r)
        for k in range(r - t):
            block = x[:, n * cout:(n + 1) * cout].to(torch.float32)
            for par in (0, 1):
                src = block[:, :, par::2]
                for a in range(src.shape[2]):
                    y = par * r + i + 2 * r * a
                    for b in range(W):
                        c = par * r + 1 + t + 2 * k + 2 * r * b
                        canvas[:, :, y, c] = src[:, :, a, b]
                        canvas[:, :, y, c + pair] = src[:, :, a, b]
            n += 1
    hexed = canvas[:, :, :, 1::2]
    return hexed[:, :, r - 1:rows_c - (r - 1), r // 2:hexed.shape[3] - (r + 1) // 2].contiguous()


# --------------------------------------------------------------------------
# hex transposed convolution (retired; "codes in old versions.txt":129-274)
# --------------------------------------------------------------------------
def hex_conv_transpose2d(x, kernel, bias=None, even_odd_offset=0, radius=2, stride=1, groups=1):
    """The reference's own route, step by step: paint the input into a zero canvas in doubled coordinates (:186-202),
    frame it (:203-204), expand the hex kernel into its dense (2r-1) x (4r-3) window (:216-224), run one strided
    correlation for the even output rows and one -- started ``stride`` rows lower and ``stride`` sub-columns further
    right -- for the odd ones (:232-243), trim to the common width (:244-263) and interleave (:265-270)."""
    r, s, eo = int(radius), int(stride), int(even_odd_offset)
    while x.dim() < 4:
        x = x.unsqueeze(0)
    B, C, H, W = x.shape
    p = r - 1
    w1 = 2 * s * W - s + 2 + (1 - s % 2)
    h1 = s * H - s + 1
    # the reference assigns whole strided slices, which only works when their lengths equal the input's (:194-202)
    for par, nrows in ((0, (H + 1) // 2), (1, H // 2)):
        start = (eo if par == 0 else 1 - eo) * s
        fits = (len(range(par * s, h1, 2 * s)) == nrows and len(range(start, w1 - 1, 2 * s)) == W
                and len(range(start + 1, w1, 2 * s)) == W)
        if nrows and not fits:
            raise ValueError("input does not fit the canvas (the reference raises a shape mismatch)")
    canvas = torch.zeros(B, C, h1 + 2 * p, w1 + 4 * p, dtype=torch.float32)
    for i in range(H):
        start = (eo if i % 2 == 0 else 1 - eo) * s
        y = p + s * i                                   # even rows 2a -> 2s*a, odd rows 2a+1 -> s + 2s*a
        for j in range(W):
            c = 2 * p + start + 2 * s * j
            canvas[:, :, y, c] = x[:, :, i, j].float()
            canvas[:, :, y, c + 1] = x[:, :, i, j].float()
    dense = torch.zeros(kernel.shape[0], kernel.shape[1], 2 * r - 1, 4 * r - 3, dtype=torch.float32)
    for k, (a, t, m) in enumerate(hex_taps(r)):
        dense[:, :, a, t + 2 * m] += kernel[:, :, 0, k].float()
    b = None if bias is None else bias.float()
    even = F.conv2d(canvas[:, :, :, 1:canvas.shape[3] - s], dense, b, stride=(2, 2), groups=groups)
    odd = F.conv2d(canvas[:, :, s:, s + 1:], dense, b, stride=(2, 2), groups=groups)
    wmin = min(even.shape[3], odd.shape[3])
    even, odd = even[..., :wmin], odd[..., :wmin]
    if even.shape[2] - odd.shape[2] not in (0, 1):
        raise ValueError("even / odd rows cannot be interleaved (the reference raises a shape mismatch)")
    out = torch.empty(B, kernel.shape[0], even.shape[2] + odd.shape[2], wmin, dtype=torch.float32)
    out[:, :, 0::2] = even
    out[:, :, 1::2] = odd
    return out


# --------------------------------------------------------------------------
# learned lattice resamplers (retired; "codes in old versions.txt":1-66, 421-493, 587-636)
# --------------------------------------------------------------------------
# All three are depthwise (one small kernel per channel) weighted gathers on DOUBLED coordinates.  Restated here as the
# closed form  y[R, J] = sum_t w[c, t] * A(sy*R + ry[t], sx*J + ex[t])  with A the doubled ("type1") view of the padded
# hex lattice -- A(i, c) = P[i, (c - s_i) // 2] for s_i <= c < 2*Wp + s_i (s_i = (i + o) % 2), 0 elsewhere -- or the
# plain padded image for the square source; differentiable (autograd supplies the backward oracle).
def _doubled(P, o, rows, dcols):
    """P: (N,C,Hp,Wp); rows (Ho,1) and doubled columns (Ho,Wo) index tensors -> (N,C,Ho,Wo) values of the type1 view."""
    Wp = P.shape[3]
    si = (rows + o) % 2
    ok = (dcols >= si) & (dcols < 2 * Wp + si)
    pc = torch.div(dcols - si, 2, rounding_mode="floor").clamp(0, Wp - 1)
    return P[:, :, rows.expand_as(dcols), pc] * ok.to(P.dtype)


def resampler_weight(f, kind):
    """Initial (C-independent) f x f weights of the three layers: inverse distance to the output sample, normalised.
    kind 'hex_to_square' (:36-49), 'square_to_hex' (:445-459), 'original_resolution' (:616-623)."""
    x = torch.arange(0, f).float()
    coor = torch.cartesian_prod(x, x).view(f, f, 2)
    a, b = coor[:, :, 0], coor[:, :, 1]
    if kind == "hex_to_square":
        d2 = (a - (f - 1) / 2) * (a - (f - 1) / 2) + (0.5 * a + b - 3 * (f - 1) / 4) * (0.5 * a + b - 3 * (f - 1) / 4)
    elif kind == "square_to_hex":
        d2 = (a - (f - 1) / 2) * (a - (f - 1) / 2) + (b - (f - 1) / 2) * (b - (f - 1) / 2)
    else:
        d2 = (a + b - (f - 1)) * (a + b - (f - 1)) + (0.5 * a - 0.5 * b) * (0.5 * a - 0.5 * b)
    dist = 1 / torch.sqrt(d2)
    return dist / dist.sum()


def hex_to_square_double_stride(x, kernel, even_odd_offset=0, downsample_factor=2, padding=0, padding_mode="constant",
                                padding_value=0):
    """Hex_to_Square_Conv2d_by_Double_Stride.forward (:50-64): kernel (C, f, f); tap (i, m) sits at type1 row i, sub-column
    i + 2m of the window (:53-54); the window moves f rows and 2f-1 sub-columns per output (:34) over type1[..., 1:]
    (one more column is cut on the right when the padded parity is odd, :60)."""
    f = int(downsample_factor)
    while x.dim() < 4:
        x = x.unsqueeze(0)
    P = F.pad(x, (padding,) * 4, padding_mode, padding_value)
    o = (even_odd_offset + padding) % 2
    Hp, Wp = P.shape[-2:]
    wt = 2 * Wp if o % 2 == 0 else 2 * Wp - 1
    Ho, Wo = (Hp - f) // f + 1, (wt - (3 * f - 2)) // (2 * f - 1) + 1
    R = torch.arange(Ho).view(Ho, 1)
    J = torch.arange(Wo).view(1, Wo)
    y = 0
    for i in range(f):
        for m in range(f):
            y = y + kernel[:, i, m].view(1, -1, 1, 1) * _doubled(P, o, f * R + i, 1 + (2 * f - 1) * J + i + 2 * m + 0 * R)
    return y


def square_to_hex_double_stride(x, kernel, padding=0, padding_mode="constant", padding_value=0):
    """Square_to_Hex_Conv2d_by_Double_Stride.forward (:461-489) for downsample_factor 2 (the unfold is hard-wired to a
    2 x 2 window, :468-469, so every other factor raises a size mismatch in the reference): kernel (C, 4);
    out[R, J] = sum_{i,j<2} k[2i+j] * P[2R + i, 2J + (R odd) + j] -- a learned 2 x 2 box whose odd rows start one pixel
    (half a hex cell) further right.  Even rows come from P[..., :-1], odd rows from P[:, 2:, 1:] (:468-469)."""
    while x.dim() < 4:
        x = x.unsqueeze(0)
    P = F.pad(x, (padding,) * 4, padding_mode, padding_value)
    Hp, Wp = P.shape[-2:]
    he, ho = math.ceil((Hp - 1) / 4), math.ceil((Hp - 3) / 4)
    Wo = int((Wp - 2) / 2)
    if he - ho not in (0, 1) or he < 1 or Wo < 1:
        raise ValueError("even / odd rows cannot be interleaved (the reference raises a shape mismatch)")
    Ho = he + ho
    R = torch.arange(Ho).view(Ho, 1)
    J = torch.arange(Wo).view(1, Wo)
    y = 0
    for i in range(2):
        for j in range(2):
            y = y + kernel[:, 2 * i + j].view(1, -1, 1, 1) * P[:, :, (2 * R + i).expand(Ho, Wo), 2 * J + (R % 2) + j]
    return y


def hex_to_square_original_resolution(x, kernel, even_odd_offset=0, padding=0, padding_mode="constant", padding_value=0):
    """Hex_to_Square_original_resolution.forward (:624-636): kernel (C, 4).  Even rows of the padded lattice are kept; odd
    rows 1, 3, ... (< Hp - 1; they sit half a cell to the side) are re-interpolated at the even rows' column positions
    from the rhombus {(R-1, 2J+2), (R, 2J+1), (R, 2J+3), (R+1, 2J+2)} of the type1 view (:630-632, unfold :645-669);
    the first column is dropped (:635)."""
    while x.dim() < 4:
        x = x.unsqueeze(0)
    P = F.pad(x, (padding,) * 4, padding_mode, padding_value)
    o = (even_odd_offset + padding) % 2
    Hp, Wp = P.shape[-2:]
    if Hp < 3:
        raise ValueError("fewer than three rows: the reference's unfold raises a shape mismatch")
    out = P[:, :, :, 1:].clone()
    rows = torch.arange(1, Hp - 1, 2).view(-1, 1)
    J = torch.arange(Wp - 1).view(1, -1)
    if rows.numel() and Wp > 1:
        tmp = 0
        for t, (dr, dc) in enumerate(((-1, 2), (0, 1), (0, 3), (1, 2))):
            tmp = tmp + kernel[:, t].view(1, -1, 1, 1) * _doubled(P, o, rows + dr, 2 * J + dc + 0 * rows)
        out[:, :, 1:Hp - 1:2] = tmp
    return out
