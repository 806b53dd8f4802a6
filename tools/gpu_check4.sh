#!/bin/bash
TAG=${1:-r1d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for B in "132 35 1 0" "132 35 -4 0" "132 35 3 5" "132 35 4 -1" "132 35 -4 -2" "132 35 296 0"; do timeout 60 tools/probes/tma_probe 0 $B >> $OUT/probe.log 2>&1; echo "rc=$?" >> $OUT/probe.log; done
cat $OUT/probe.log
timeout 900 python tools/bench_conv.py --reps 10 > $OUT/bench_conv.log 2>&1; echo "bench_conv rc=$?"; cat $OUT/bench_conv.log | grep -v '"rows"'
timeout 600 python tools/bench_conv.py --reps 5 --batch 16 --direct > $OUT/bench_conv_direct_b16.log 2>&1; echo "bench_conv direct rc=$?"; cat $OUT/bench_conv_direct_b16.log | grep -v '"rows"'
