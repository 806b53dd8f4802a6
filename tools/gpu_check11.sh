#!/bin/bash
TAG=${1:-r1l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for d in 0 2 4 8 1; do
  echo "== HG_WU_DBG=$d"; HG_WU_DBG=$d timeout 300 python tools/debug/wgrad_one.py 2>&1 | tail -6
done > $OUT/wgrad_dbg.log 2>&1
cat $OUT/wgrad_dbg.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python tools/debug/wgrad_one.py > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -40 $OUT/racecheck.log
timeout 600 compute-sanitizer --tool memcheck python tools/debug/wgrad_one.py > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -15 $OUT/memcheck.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "not tcgen05" > $OUT/pytest_main.log 2>&1; echo "pytest(main) rc=$?"; tail -5 $OUT/pytest_main.log
timeout 900 python tools/bench_path.py --reps 10 > $OUT/bench_path.log 2>&1; echo "bench_path rc=$?"; grep -v '"rows"' $OUT/bench_path.log
