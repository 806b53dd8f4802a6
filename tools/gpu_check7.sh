#!/bin/bash
TAG=${1:-r1h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "tcgen05" > $OUT/pytest_tcgen05.log 2>&1; echo "pytest(tcgen05) rc=$?"; tail -3 $OUT/pytest_tcgen05.log
timeout 900 python tools/bench_conv.py --reps 10 > $OUT/bench_conv.log 2>&1; echo "bench_conv rc=$?"; grep -v '"rows"' $OUT/bench_conv.log
HG_CONV_NO_TMA=1 timeout 900 python tools/bench_conv.py --reps 10 > $OUT/bench_conv_notma.log 2>&1; echo "bench_conv(no TMA) rc=$?"; grep -v '"rows"' $OUT/bench_conv_notma.log | head -3
timeout 900 python tools/bench_path.py --reps 10 > $OUT/bench_path.log 2>&1; echo "bench_path rc=$?"; grep -v '"rows"' $OUT/bench_path.log
