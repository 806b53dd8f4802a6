#!/bin/bash
# GPU run 3: bisect the TMA fault with the standalone probe, run the suite with the tiled kernel off,
# then the tcgen05 conv tests in their own process, then memcheck one tiled launch.
TAG=${1:-r1c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in 0 1 2 4 7; do timeout 60 tools/probes/tma_probe $V 132 35 -1 0 >> $OUT/probe.log 2>&1; echo "rc=$?" >> $OUT/probe.log; done
for B in "80 52 -1 0" "80 52 0 0" "128 32 0 0" "64 16 0 0" "132 35 200 190"; do timeout 60 tools/probes/tma_probe 0 $B >> $OUT/probe.log 2>&1; echo "rc=$?" >> $OUT/probe.log; done
cat $OUT/probe.log
HG_R2H_ROWS=0 timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "not tcgen05" > $OUT/pytest_notma.log 2>&1; echo "pytest(no tma, no tcgen05) rc=$?"; tail -3 $OUT/pytest_notma.log
HG_R2H_ROWS=0 timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "tcgen05" > $OUT/pytest_tcgen05.log 2>&1; echo "pytest(tcgen05) rc=$?"; tail -15 $OUT/pytest_tcgen05.log
HG_R2H_ROWS=0 timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
cat > /tmp/tma_case.py <<'PY'
import sys, torch
sys.path.insert(0, 'hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200')
from HyGrid import functional as Fn
x = torch.rand(1, 200, 300, device='cuda')
y = Fn.rect_to_hex(x, (129, 500), 'bilinear', out_dtype=torch.float32, math='fast')
torch.cuda.synchronize(); print('ok', float(y.sum()))
PY
HG_R2H_ROWS=4 timeout 600 compute-sanitizer --tool memcheck python /tmp/tma_case.py > $OUT/sanitizer_tma.log 2>&1; echo "sanitizer rc=$?"; grep -v "^$" $OUT/sanitizer_tma.log | head -40
ls $OUT
