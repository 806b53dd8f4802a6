OUT=gpurun_out/r4e; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_resample.py tests/test_gpu_baseline_sizes.py tests/test_gpu_stream_kernel.py tests/test_zz_numba_twin.py -x -q -m gpu > $OUT/pytest_h2r.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_h2r.log
P="python tools/bench_path.py --reps 10 --only"
timeout 600 $P "r3 hex warp" > $OUT/warp.jsonl 2>$OUT/warp.err; cat $OUT/warp.jsonl | cut -c1-250
timeout 600 $P "c4 hex->rect linear exact f32" > $OUT/plain.log 2>&1; cat $OUT/plain.log | cut -c1-250
P="python tools/bench_path.py --reps 2 --only"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexsrc_linear_tma -s 2 -c 1 -o $OUT/prof_h2r_exact $P "c4 hex->rect linear exact f32" > $OUT/ncu_exact.log 2>&1; echo "ncu exact rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexsrc_linear_fast -s 2 -c 1 -o $OUT/prof_warp_linear $P "r3 hex warp affine linear" > $OUT/ncu_warp.log 2>&1; echo "ncu warp rc=$?"
