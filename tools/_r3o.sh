OUT=gpurun_out/r3o; mkdir -p $OUT
timeout 900 python tools/sweep_kernels.py --what h2r --reps 10 > $OUT/sweep.jsonl 2> $OUT/sweep.err; echo "sweep rc=$?"; grep -E "promotion|heuristic" $OUT/sweep.jsonl | cut -c1-250; tail -3 $OUT/sweep.err
timeout 600 python -m pytest tests/test_zz_hex_mosaic.py tests/test_zz_pixel_shuffle.py -q -m gpu 2>&1 | tail -2
timeout 600 python tools/bench_path.py --reps 10 --only "hex mosaic" 2>&1 | grep -v rows | cut -c1-200
timeout 300 python tools/hexcnn_ddp.py --batch 64 --steps 30 --autocast --blocking 2>&1 | grep '^{' | cut -c1-200
