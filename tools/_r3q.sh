OUT=gpurun_out/r3q; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/pytest_gpu.log | cut -c1-300 | head -30
timeout 900 python tools/sweep_kernels.py --what h2r --reps 10 > $OUT/sweep.jsonl 2> $OUT/sweep.err; echo "sweep rc=$?"; grep -E "h2r fast" $OUT/sweep.jsonl | grep -v promotion | cut -c1-250; tail -3 $OUT/sweep.err
timeout 600 python tools/bench_c3.py > $OUT/bench_c3.json 2>&1; tail -1 $OUT/bench_c3.json | cut -c1-300
