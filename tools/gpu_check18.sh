#!/bin/bash
TAG=${1:-r1u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "hexconv or tcgen05" > $OUT/pytest_conv.log 2>&1; echo "pytest(conv) rc=$?"; tail -4 $OUT/pytest_conv.log | cut -c1-300
HG_CONV_NO_TMA=1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "tcgen05" > $OUT/pytest_conv_notma.log 2>&1; echo "pytest(conv,no tma) rc=$?"; tail -2 $OUT/pytest_conv_notma.log | cut -c1-300
timeout 600 python tools/bench_conv.py --reps 10 > $OUT/bench_conv.log 2>&1; echo "bench_conv rc=$?"; grep -v '"rows"' $OUT/bench_conv.log | cut -c1-220
timeout 300 python tools/hexcnn_ddp.py --batch 64 --steps 10 --autocast > $OUT/hexcnn_1gpu_autocast.log 2>&1; echo "hexcnn autocast rc=$?"; tail -1 $OUT/hexcnn_1gpu_autocast.log
PROF="python tools/bench_conv.py --reps 2 --only fwd --dtypes f32f32"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexconv_umma -s 2 -c 1 -o $OUT/prof_conv_fwd_f32 $PROF > $OUT/ncu_conv.log 2>&1; echo "ncu fwd f32 rc=$?"
