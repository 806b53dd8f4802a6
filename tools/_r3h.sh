OUT=gpurun_out/r3h; mkdir -p $OUT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3h/bench_n2.json') if l.startswith('{')][-1])
print("value",d["value"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"], "n", d["n_gpus"])
for k,v in d["extra"].items(): print(k, json.dumps(v)[:1500])
PY
tail -5 $OUT/bench_n2.err
for v in "" "--blocking"; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/hexcnn_ddp.py --batch 64 --steps 30 --autocast $v 2>&1 | grep '^{' ; done
