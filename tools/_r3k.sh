OUT=gpurun_out/r3k; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_hexframes.py tests/test_gpu_baseline_sizes.py -q --timeout 600 -x > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/pytest_sel.log | cut -c1-250 | head
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3k/bench_n2.json') if l.startswith('{')][-1])
print("value",d["value"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"], "n", d["n_gpus"])
print(json.dumps(d["extra"]["c3"])[:1200]); print(json.dumps(d["extra"]["c5"])[:2500])
PY
tail -3 $OUT/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/bench_pcie.py > $OUT/bench_pcie_n2.json 2> $OUT/pcie.err; grep '^{' $OUT/bench_pcie_n2.json | cut -c1-700
timeout 600 python tools/bench_c3.py > $OUT/bench_c3.json 2>&1; tail -1 $OUT/bench_c3.json | cut -c1-300
