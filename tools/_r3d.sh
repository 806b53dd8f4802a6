OUT=gpurun_out/r3d; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_stream_kernel.py tests/test_zz_learned_resamplers.py tests/test_gpu_resample.py tests/test_zz_numpy_api.py -q --timeout 600 -x > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/pytest_sel.log | cut -c1-250 | head -30
timeout 900 python tools/sweep_kernels.py --reps 10 > $OUT/sweep.jsonl 2> $OUT/sweep.err; echo "sweep rc=$?"; grep -E "stream v|specialised|h2r" $OUT/sweep.jsonl | cut -c1-250; tail -3 $OUT/sweep.err
