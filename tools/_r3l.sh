OUT=gpurun_out/r3l; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/pytest_gpu.log | cut -c1-300 | head -30
timeout 900 python tools/bench_conv.py --cudnn-only --reps 9 > $OUT/bench_conv_cudnn.jsonl 2> $OUT/bench_conv_cudnn.err; echo "cudnn route rc=$?"; grep -v '"rows"' $OUT/bench_conv_cudnn.jsonl | cut -c1-330; tail -3 $OUT/bench_conv_cudnn.err
timeout 300 python tools/hexcnn_ddp.py --batch 64 --steps 20 2>&1 | grep '^{' | cut -c1-300
timeout 300 python tools/hexcnn_ddp.py --batch 64 --steps 20 --autocast 2>&1 | grep '^{' | cut -c1-300
compute-sanitizer --version 2>&1 | head -3
