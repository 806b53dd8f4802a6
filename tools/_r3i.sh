OUT=gpurun_out/r3i; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/pytest_gpu.log | cut -c1-250 | head -30
timeout 900 python tools/bench_path.py --reps 10 > $OUT/bench_path.jsonl 2> $OUT/bench_path.err; echo "bench_path rc=$?"; grep -v '"rows"' $OUT/bench_path.jsonl | cut -c1-200; tail -3 $OUT/bench_path.err
P="python tools/bench_path.py --reps 2 --small --only"
timeout 600 $P "pixel shuffle" > $OUT/plain_gather.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:plane_gather -s 2 -c 1 -o $OUT/prof_gather $P "pixel shuffle" > $OUT/ncu_gather.log 2>&1; echo "ncu gather rc=$?"
timeout 600 python tools/bench_path.py --reps 2 --only "c2 rect->hex bilinear fast" > $OUT/plain_stream.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:rect2hex_stream -s 2 -c 1 -o $OUT/prof_stream_v3 python tools/bench_path.py --reps 2 --only "c2 rect->hex bilinear fast" > $OUT/ncu_stream.log 2>&1; echo "ncu stream rc=$?"
timeout 300 python tools/bench_pcie.py > $OUT/bench_pcie_n1.json 2>&1; cat $OUT/bench_pcie_n1.json | cut -c1-600
timeout 600 python tools/hexcnn_ddp.py --batch 64 --steps 20 --autocast 2>&1 | grep '^{' | cut -c1-400
