#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list and one --set full capture of the
# dominant kernel.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"
cat $OUT/bench_c2.json
timeout 600 python bench.py --steps 20 --warmup 5 --math exact --no-cpu --e2e-steps 2 > $OUT/bench_c2_exact.json 2> $OUT/bench_c2_exact.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload c2half --no-cpu --e2e-steps 2 > $OUT/bench_c2half.json 2> $OUT/bench_c2half.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
PROF="python bench.py --steps 4 --warmup 3 --no-cpu --e2e-steps 1"
timeout 300 $PROF > $OUT/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv $PROF > $OUT/ncu_launches.log 2>&1
timeout 300 $PROF > $OUT/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rect2hex_bilinear -s 3 -c 2 -o $OUT/prof_rect2hex $PROF > $OUT/ncu_full.log 2>&1
ls -la $OUT
