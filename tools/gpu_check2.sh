#!/bin/bash
# GPU run 2: full parity suite, smoke, A/B of the rect->hex kernels (direct vs TMA-tiled, 4 / 8 rows per warp),
# ncu --set full of the tiled kernel.
TAG=${1:-r1b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
for R in 0 4 8; do
  for M in fast exact; do
    HG_R2H_ROWS=$R timeout 600 python bench.py --steps 20 --warmup 5 --math $M --no-cpu --e2e-steps 2 > $OUT/bench_c2_rows${R}_$M.json 2> $OUT/bench_c2_rows${R}_$M.err
    python -c "import json;d=json.load(open('$OUT/bench_c2_rows${R}_$M.json'));print('rows',$R,'$M',round(d['value']),round(d['roofline']['frac'],3),round(d['e2e']['value']))"
  done
done
HG_R2H_ROWS=4 timeout 600 python bench.py --steps 20 --warmup 5 --workload c4 --no-cpu --e2e-steps 1 > $OUT/bench_c4_rows4.json 2> $OUT/bench_c4_rows4.err
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cat $OUT/bench_c2.json
PROF="python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 1"
timeout 300 $PROF > $OUT/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv $PROF > $OUT/ncu_launches.log 2>&1
timeout 300 $PROF > $OUT/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rect2hex_bilinear -s 3 -c 2 -o $OUT/prof_rect2hex $PROF > $OUT/ncu_full.log 2>&1
ls -la $OUT
