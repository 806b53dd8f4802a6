#!/bin/bash
TAG=${1:-r1m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for k in 0 1 2 3 4 5 6 7 8 9 10 11; do
  echo "== process $k HG_WU_DBG=48"; HG_WU_DBG=48 timeout 300 python tools/debug/wgrad_loop.py 4 $k 2>&1 | tail -30
done > $OUT/wgrad_loop48.log 2>&1
grep -c "bad iterations" $OUT/wgrad_loop48.log; grep -v " 0/4 bad" $OUT/wgrad_loop48.log | head -80
for k in 0 1 2 3 4 5 6 7; do
  echo "== process $k HG_WU_DBG=0"; HG_WU_DBG=0 timeout 300 python tools/debug/wgrad_cfg.py 2>&1 | grep BAD
done > $OUT/wgrad_cfg0.log 2>&1
cat $OUT/wgrad_cfg0.log | head -60
