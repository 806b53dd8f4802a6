#!/bin/bash
TAG=${1:-r1j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "not tcgen05" > $OUT/pytest_main.log 2>&1; echo "pytest(main) rc=$?"; tail -3 $OUT/pytest_main.log
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "tcgen05" > $OUT/pytest_tcgen05.log 2>&1; echo "pytest(tcgen05) rc=$?"; tail -3 $OUT/pytest_tcgen05.log
HG_CONV_NO_TMA=1 timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "tcgen05" > $OUT/pytest_tcgen05_notma.log 2>&1; echo "pytest(tcgen05,no TMA) rc=$?"; tail -3 $OUT/pytest_tcgen05_notma.log
timeout 900 python tools/bench_conv.py --reps 10 > $OUT/bench_conv.log 2>&1; echo "bench_conv rc=$?"; grep -v '"rows"' $OUT/bench_conv.log
HG_CONV_NO_TMA=1 timeout 900 python tools/bench_conv.py --reps 10 --dtypes f32f32 > $OUT/bench_conv_notma.log 2>&1; echo "bench_conv(no TMA) rc=$?"; grep -v '"rows"' $OUT/bench_conv_notma.log | head -3
timeout 900 python tools/bench_path.py --reps 10 > $OUT/bench_path.log 2>&1; echo "bench_path rc=$?"; grep -v '"rows"' $OUT/bench_path.log
timeout 300 python tools/hexcnn_ddp.py --batch 64 --steps 10 > $OUT/hexcnn_1gpu.log 2>&1; echo "hexcnn rc=$?"; tail -2 $OUT/hexcnn_1gpu.log
timeout 300 python tools/hexcnn_ddp.py --batch 64 --steps 10 --autocast > $OUT/hexcnn_1gpu_autocast.log 2>&1; echo "hexcnn autocast rc=$?"; tail -2 $OUT/hexcnn_1gpu_autocast.log
