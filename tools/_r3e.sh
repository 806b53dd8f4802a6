OUT=gpurun_out/r3e; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/pytest_gpu.log | cut -c1-250 | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3e/bench_c2.json').read().strip().splitlines()[-1])
print("value",d["value"],"frac",d["roofline"]["frac"],d["roofline"]["kernel"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"])
for k,v in d["extra"].items(): print(k, json.dumps(v)[:900])
print("cpu", d.get("cpu_baseline",{}).get("value"))
PY
tail -3 $OUT/bench_c2.err
