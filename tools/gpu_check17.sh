#!/bin/bash
TAG=${1:-r1t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python tools/bench_conv.py --reps 10 > $OUT/bench_conv.log 2>&1; echo "bench_conv rc=$?"; grep -v '"rows"' $OUT/bench_conv.log | cut -c1-220
PROF="python tools/bench_conv.py --reps 2 --only fwd --dtypes f32f32"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexconv_umma -s 2 -c 1 -o $OUT/prof_conv_fwd_f32 $PROF > $OUT/ncu_conv.log 2>&1; echo "ncu fwd f32 rc=$?"
PROF="python tools/bench_conv.py --reps 2 --only fwd --dtypes bf16bf16"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexconv_umma -s 2 -c 1 -o $OUT/prof_conv_fwd_bf16 $PROF > $OUT/ncu_conv2.log 2>&1; echo "ncu fwd bf16 rc=$?"
PROF2="python tools/bench_conv.py --reps 2 --only wgrad --dtypes f32f32"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexconv_wgrad -s 2 -c 1 -o $OUT/prof_conv_wgrad_f32 $PROF2 > $OUT/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > $OUT/bench_numa.json 2> $OUT/bench_numa.err; python -c "
import json;d=json.loads(open('$OUT/bench_numa.json').read().strip().splitlines()[-1]);print(d['config'].get('numa'), d['e2e']['value'])"
