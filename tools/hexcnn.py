"""The hex CNN of BASELINE config 5 (builder-defined): HexConvModule 3->32->64->128 (BN + ReLU),
HexPool2d('max', 2, 2) between, global average, Linear -> 10, on 128 x 128 hex lattices.

``HexCNN``: the product model (HyGrid modules over libhygrid_b200.so) -- used by bench.py, tools/hexcnn_ddp.py and the parity
test.  Its CPU twin built from the oracle's operators lives with the tests (tests/hexcnn_oracle.py).
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
