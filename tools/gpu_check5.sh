#!/bin/bash
TAG=${1:-r1f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "not tcgen05" > $OUT/pytest_main.log 2>&1; echo "pytest(main) rc=$?"; tail -4 $OUT/pytest_main.log
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "tcgen05" > $OUT/pytest_tcgen05.log 2>&1; echo "pytest(tcgen05) rc=$?"; tail -12 $OUT/pytest_tcgen05.log
HG_CONV_NO_TMA=1 timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "tcgen05" > $OUT/pytest_tcgen05_notma.log 2>&1; echo "pytest(tcgen05, no TMA) rc=$?"; tail -4 $OUT/pytest_tcgen05_notma.log
for M in fast exact; do
  timeout 600 python bench.py --steps 20 --warmup 5 --math $M --no-cpu --e2e-steps 2 > $OUT/bench_c2_$M.json 2> $OUT/bench_c2_$M.err
  python -c "import json;d=json.load(open('$OUT/bench_c2_$M.json'));print('c2','$M',round(d['value']),round(d['roofline']['frac'],3),round(d['e2e']['value']))"
done
timeout 900 python tools/bench_conv.py --reps 10 > $OUT/bench_conv.log 2>&1; echo "bench_conv rc=$?"; grep -v '"rows"' $OUT/bench_conv.log
HG_CONV_NO_TMA=1 timeout 900 python tools/bench_conv.py --reps 10 > $OUT/bench_conv_notma.log 2>&1; echo "bench_conv(no TMA) rc=$?"; grep -v '"rows"' $OUT/bench_conv_notma.log
