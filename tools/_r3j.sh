OUT=gpurun_out/r3j; mkdir -p $OUT
timeout 1500 python bench.py --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3j/bench_c2.json') if l.startswith('{')][-1])
print("value",d["value"],"frac",d["roofline"]["frac"],d["roofline"]["kernel"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"])
print(json.dumps(d["extra"]["c5"])[:2000])
PY
tail -3 $OUT/bench_c2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench(reference) rc=$?"; cut -c1-300 $OUT/bench_ref.json
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --e2e-steps 1"
timeout 600 $B > $OUT/plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/bench_launches.csv $B > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
C3="python tools/bench_c3.py --reps 1"
timeout 300 $C3 > $OUT/plain_c3.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/c3_launches.csv $C3 > $OUT/ncu_c3.log 2>&1; echo "c3 launch list rc=$?"; tail -1 $OUT/plain_c3.log | cut -c1-300
timeout 900 python tools/bench_conv.py --reps 10 > $OUT/bench_conv.jsonl 2>&1; echo "bench_conv rc=$?"; grep -v '"rows"' $OUT/bench_conv.jsonl | cut -c1-220
