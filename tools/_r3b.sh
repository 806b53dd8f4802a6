OUT=gpurun_out/r3b; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_stream_kernel.py -q --timeout 600 -k "c3_layer_full or c5_train or stream or other_geom" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/pytest_sel.log | cut -c1-250 | head -30
timeout 900 python tools/sweep_kernels.py --what r2h --reps 10 > $OUT/sweep_r2h.jsonl 2> $OUT/sweep_r2h.err; echo "sweep rc=$?"; grep -E "fast" $OUT/sweep_r2h.jsonl | cut -c1-230; tail -3 $OUT/sweep_r2h.err
P="python tools/bench_path.py --reps 2 --small --only"
HG_HEXSRC_SHARE=64 timeout 600 $P "c4 hex->rect linear exact" > $OUT/plain_hexsrc_exact.log 2>&1 && HG_HEXSRC_SHARE=64 timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexsrc_linear_tma -s 2 -c 1 -o $OUT/prof_hexsrc_exact $P "c4 hex->rect linear exact" > $OUT/ncu_hexsrc_exact.log 2>&1; echo "ncu hexsrc exact rc=$?"
timeout 600 $P "c2 rect->hex bilinear fast" > $OUT/plain_stream.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:rect2hex_bilinear_stream -s 2 -c 1 -o $OUT/prof_stream $P "c2 rect->hex bilinear fast" > $OUT/ncu_stream.log 2>&1; echo "ncu stream rc=$?"
ls -la $OUT
