#!/bin/bash
TAG=${1:-r1p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest(gpu) rc=$?"; tail -6 $OUT/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench_c2.json'));print('c2',round(d['value']),round(d['roofline']['frac'],3),round(d['e2e']['value']),d['roofline']['traffic'])"
timeout 900 python tools/bench_path.py --reps 10 > $OUT/bench_path.log 2>&1; echo "bench_path rc=$?"; grep -v '"rows"' $OUT/bench_path.log | cut -c1-200
timeout 300 python tools/hexcnn_ddp.py --batch 64 --steps 10 --autocast > $OUT/hexcnn_1gpu_autocast.log 2>&1; echo "hexcnn autocast rc=$?"; tail -1 $OUT/hexcnn_1gpu_autocast.log
timeout 300 python tools/hexcnn_ddp.py --batch 64 --steps 10 > $OUT/hexcnn_1gpu.log 2>&1; echo "hexcnn fp32 rc=$?"; tail -1 $OUT/hexcnn_1gpu.log
B="python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rect2hex_bilinear_tma -s 3 -c 1 -o $OUT/prof_rect2hex_tma $B > $OUT/ncu_r2h.log 2>&1; echo "ncu r2h rc=$?"
P="python tools/bench_path.py --reps 2 --small --only"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexsrc_linear_tma -s 2 -c 1 -o $OUT/prof_hexsrc_tma $P "c4 hex->rect linear fast" > $OUT/ncu_hexsrc.log 2>&1; echo "ncu hexsrc rc=$?"
