#!/bin/bash
TAG=${1:-r1i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
PROF="python tools/bench_conv.py --reps 2 --only fwd --dtypes f32f32"
timeout 300 $PROF > $OUT/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexconv_umma -s 2 -c 1 -o $OUT/prof_conv_fwd $PROF > $OUT/ncu_conv.log 2>&1
tail -3 $OUT/ncu_conv.log
PROF2="python tools/bench_conv.py --reps 2 --only wgrad --dtypes f32f32"
timeout 300 $PROF2 > $OUT/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexconv_wgrad -s 2 -c 1 -o $OUT/prof_conv_wgrad $PROF2 > $OUT/ncu_wgrad.log 2>&1
tail -3 $OUT/ncu_wgrad.log
ls -la $OUT
