"""The hex CNN of BASELINE config 5 (builder-defined): HexConvModule 3->32->64->128 (BN + ReLU),
HexPool2d('max', 2, 2) between, global average, Linear -> 10, on 128 x 128 hex lattices.

``HexCNN``        the product model (HyGrid modules over libhygrid_b200.so) -- used by bench.py, tools/hexcnn_ddp.py and
                  the parity test.
``oracle_forward`` the same network evaluated with the CPU oracle's operators (oracle/hexframes_oracle.py: the closed
                  forms of HexConv2d / HexPool2d pinned against the live reference) and torch's own batch norm / linear, on
                  the product model's parameters.  Test infrastructure: only tests/ and bench.py's CPU leg call it.
"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


class HexCNN(nn.Module):
    def __init__(self, classes=10, widths=(32, 64, 128)):
        super().__init__()
        from HyGrid import HexFrames as hf
        from HyGrid.HexModules import HexConvModule
        bn = dict(type='BN')
        c1, c2, c3 = widths
        self.c1 = HexConvModule(3, c1, 0, 2, padding=1, norm_cfg=bn)
        self.c2 = HexConvModule(c1, c2, 0, 2, padding=1, norm_cfg=bn)
        self.c3 = HexConvModule(c2, c3, 0, 2, padding=1, norm_cfg=bn)
        self.pool = hf.HexPool2d('max', 2, 2)
        self.gap = hf.HexGlobalPool2d('average')
        self.fc = nn.Linear(c3, classes)

    def forward(self, x):
        x = self.pool(self.c1(x))
        x = self.pool(self.c2(x))
        x = self.c3(x)
        return self.fc(self.gap(x))


class _RoundGradBf16(torch.autograd.Function):
    """Identity whose backward rounds the gradient to bfloat16: the tensor-core data / weight gradient kernels read the
    float32 gradient of a conv output through a bfloat16 conversion on its way into shared memory."""

    @staticmethod
    def forward(ctx, y):
        return y.view_as(y)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _round_bf16(t):
    """Value rounded to bfloat16, gradient passed through unchanged (what a kernel that rounds its operand on load does)."""
    return t + (t.bfloat16().float() - t).detach()


def oracle_forward(params, x, eps=1e-5, autocast=False):
    """Forward of the same network on CPU tensors with the oracle's operators.  ``params``: dict name -> CPU tensor
    (requires_grad as wanted) with the product model's ``named_parameters()`` names; training-mode batch norm (batch
    statistics), conv without bias (HexConvModule's bias='auto' with a norm layer).  ``autocast``: restate where the
    product rounds to bfloat16 under torch.autocast -- every conv reads bfloat16 activations and output gradients; the
    tensor-core layers (c2, c3) also round their weights (the RGB layer runs the direct stencil with float32 weights)."""
    from oracle import hexframes_oracle as HO
    F = torch.nn.functional
    for blk in ("c1", "c2", "c3"):
        w = params[f"{blk}.conv.kernel"]
        if autocast:
            x = _round_bf16(x)
            if blk != "c1":
                w = _round_bf16(w)
        x = HO.hexconv2d(x, w, None, 0, 2, 1, 1)
        if autocast:
            x = _RoundGradBf16.apply(x)
        x = F.batch_norm(x, None, None, params[f"{blk}.bn.weight"], params[f"{blk}.bn.bias"], True, 0.1, eps)
        x = F.relu(x)
        if blk != "c3":
            x = HO.hexpool2d(x, "max", 2, 2)
    x = HO.hexglobalpool2d(x, "average")
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):    # nn.Linear is an autocast op in the product too
        return F.linear(x, params["fc.weight"], params["fc.bias"])
