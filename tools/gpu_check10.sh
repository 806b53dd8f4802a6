#!/bin/bash
TAG=${1:-r1k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "not tcgen05" > $OUT/pytest_main.log 2>&1; echo "pytest(main) rc=$?"; tail -8 $OUT/pytest_main.log
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "tcgen05" > $OUT/pytest_tcgen05.log 2>&1; echo "pytest(tcgen05) rc=$?"; tail -3 $OUT/pytest_tcgen05.log
timeout 600 python tools/debug/wgrad_cfg.py > $OUT/wgrad_debug.log 2>&1; echo "wgrad debug rc=$?"; cat $OUT/wgrad_debug.log | tail -30
timeout 900 python tools/bench_path.py --reps 10 > $OUT/bench_path.log 2>&1; echo "bench_path rc=$?"; grep -v '"rows"' $OUT/bench_path.log
HG_HEXSRC_NO_TMA=1 timeout 900 python tools/bench_path.py --reps 10 > $OUT/bench_path_notma.log 2>&1; echo "bench_path(no hexsrc TMA) rc=$?"; grep -v '"rows"' $OUT/bench_path_notma.log
timeout 300 python tools/hexcnn_ddp.py --batch 64 --steps 10 --autocast > $OUT/hexcnn_1gpu_autocast.log 2>&1; echo "hexcnn autocast rc=$?"; tail -2 $OUT/hexcnn_1gpu_autocast.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"; cat $OUT/bench_c2.json
