#!/bin/bash
# 2-GPU checks: bench.py weak scaling, reference arm, hex CNN DDP step
TAG=${1:-r1s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L | head -4
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench n1 rc=$?"; cut -c1-400 $OUT/bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?"; tail -1 $OUT/bench_n2.json | cut -c1-400; tail -3 $OUT/bench_n2.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"; cut -c1-500 $OUT/bench_ref.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/hexcnn_ddp.py --batch 64 --steps 10 --autocast > $OUT/hexcnn_n2.log 2>&1; echo "hexcnn n2 rc=$?"; tail -1 $OUT/hexcnn_n2.log
timeout 300 python tools/hexcnn_ddp.py --batch 64 --steps 10 --autocast > $OUT/hexcnn_n1.log 2>&1; echo "hexcnn n1 rc=$?"; tail -1 $OUT/hexcnn_n1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
