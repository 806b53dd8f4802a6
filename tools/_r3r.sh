OUT=gpurun_out/r3r; mkdir -p $OUT
timeout 1500 python bench.py --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3r/bench_c2.json') if l.startswith('{')][-1])
e=d['extra']
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])
print('exact f64', e['c2_exact_f64']['value'], e['c2_exact_f64']['ms_per_step'], e['c2_exact_f64']['roofline']['frac'], e['c2_exact_f64']['e2e']['ms_per_step'])
for k,v in e['c4']['legs'].items(): print(k, round(v['value']), round(v['ms_per_step'],3), round(v['roofline']['frac'],3))
for k,v in e['c3']['modes'].items(): print(k, round(v['ms_per_step'],2), round(v['hex_mpix_per_s_per_layer_step']), round(v['roofline']['frac'],3), round(v['tensor']['frac'],3))
for k,v in e['c5']['variants'].items(): print(k, round(v['ms_per_step'],3), round(v['images_per_s']))
PY
tail -3 $OUT/bench_c2.err
timeout 900 python tools/bench_path.py --reps 10 > $OUT/bench_path.jsonl 2> $OUT/bench_path.err; grep -E "hex->rect|rect->hex" $OUT/bench_path.jsonl | grep -v rows | cut -c1-200
P="python tools/bench_path.py --reps 2 --only"
timeout 600 $P "c4 hex->rect linear fast" > $OUT/plain_h2r.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexsrc_stream -s 2 -c 1 -o $OUT/prof_h2r_stream $P "c4 hex->rect linear fast" > $OUT/ncu_h2r.log 2>&1; echo "ncu h2r stream rc=$?"
