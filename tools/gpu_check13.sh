#!/bin/bash
# tests + wgrad stress + conv bench + ncu captures of the memory-bound kernels
TAG=${1:-r1o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest(gpu) rc=$?"; tail -4 $OUT/pytest_gpu.log
for k in 0 1 2 3 4 5; do timeout 300 python tools/debug/wgrad_cfg.py 2>&1 | grep -c BAD; done > $OUT/wgrad_stress.log 2>&1; echo "wgrad stress (BAD counts per process):" $(cat $OUT/wgrad_stress.log | tr '\n' ' ')
timeout 900 python tools/bench_conv.py --reps 10 > $OUT/bench_conv.log 2>&1; echo "bench_conv rc=$?"; grep -v '"rows"' $OUT/bench_conv.log
timeout 300 python tools/hexcnn_ddp.py --batch 64 --steps 10 --autocast > $OUT/hexcnn_1gpu_autocast.log 2>&1; echo "hexcnn autocast rc=$?"; tail -1 $OUT/hexcnn_1gpu_autocast.log
# --- ncu: launch list of the bench, then one full capture per kernel (each only after the plain command exited 0)
B="python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1"
timeout 600 $B > $OUT/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/bench_launches.csv $B > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rect2hex_bilinear_tma -s 3 -c 1 -o $OUT/prof_rect2hex_tma $B > $OUT/ncu_r2h.log 2>&1; echo "ncu r2h rc=$?"
cap() {  # name, kernel regex, bench_path filter
  P="python tools/bench_path.py --reps 2 --small --only"
  timeout 600 $P "$3" > $OUT/plain_$1.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o $OUT/prof_$1 $P "$3" > $OUT/ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
cap hexsrc_tma hexsrc_linear_tma "c4 hex->rect linear fast"
cap hexsrc_nearest hexsrc_nearest "c4 hex->rect nearest"
cap pool_vec hexpool2x2_fwd "pool avg 2x2 level 0"
cap pool_any hexpool2x2c_fwd "pool avg 2x2 level 1"
cap pool_bwd hexpool2x2_bwd "pool max bwd"
cap type1 hex_to_type_vec "hex->type1"
cap r2h_nearest rect2hex_nearest "c2 rect->hex nearest"
# C5 launch list
H="python tools/hexcnn_ddp.py --batch 64 --steps 2 --autocast"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/hexcnn_launches.csv $H > $OUT/ncu_hexcnn.log 2>&1; echo "hexcnn launch list rc=$?"
ls -la $OUT | head -50
