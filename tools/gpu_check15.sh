#!/bin/bash
TAG=${1:-r1q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for ws in 8 16; do
  HG_R2H_WS=$ws timeout 900 python -m pytest tests/test_gpu_resample.py -m gpu -q --timeout 600 -k "r1 or config1 or host" > $OUT/pytest_r1_ws$ws.log 2>&1; echo "pytest(r1, ws=$ws) rc=$?"; tail -3 $OUT/pytest_r1_ws$ws.log | cut -c1-200
done
for ws in 0 8 16; do
  for wl in c2 c4; do
    HG_R2H_WS=$ws timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 1 --workload $wl > $OUT/bench_${wl}_ws$ws.json 2> $OUT/bench_${wl}_ws$ws.err
    python -c "
import json;d=json.load(open('$OUT/bench_${wl}_ws$ws.json'));print('$wl ws=$ws',round(d['value']),round(d['roofline']['frac'],3),round(d['roofline']['kernel_ms'],4))"
  done
done
HG_R2H_WS=8 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 1 --math exact > $OUT/bench_c2_ws8_exact.json 2>$OUT/err; python -c "
import json;d=json.load(open('$OUT/bench_c2_ws8_exact.json'));print('c2 exact ws=8',round(d['value']),round(d['roofline']['frac'],3))"
B="python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1"
for ws in 8 16; do
HG_R2H_WS=$ws timeout 900 ncu --set full --clock-control none --import-source on -k regex:rect2hex_bilinear_ws -s 3 -c 1 -o $OUT/prof_rect2hex_ws$ws $B > $OUT/ncu_r2h_ws$ws.log 2>&1; echo "ncu ws=$ws rc=$?"
done
