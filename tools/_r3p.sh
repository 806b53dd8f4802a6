OUT=gpurun_out/r3p; mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3p/bench_n8.json') if l.startswith('{')][-1])
print("value",d["value"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"], "n", d["n_gpus"])
e=d["extra"]
print("exact f64", e["c2_exact_f64"].get("value"), e["c2_exact_f64"].get("e2e",{}).get("ms_per_step"))
print("c4", {k:(round(v["value"]),round(v["ms_per_step"],3)) for k,v in e["c4"]["legs"].items()} if "legs" in e["c4"] else e["c4"])
print("c3", {k:round(v["ms_per_step"],2) for k,v in e["c3"]["modes"].items()} if "modes" in e["c3"] else e["c3"])
print("c5", {k:(round(v.get("ms_per_step",0),3),round(v.get("images_per_s",0)), v.get("ranks_in_sync"), v.get("error")) for k,v in e["c5"]["variants"].items()} if "variants" in e["c5"] else e["c5"])
PY
tail -3 $OUT/bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 tools/bench_pcie.py > $OUT/bench_pcie_n8.json 2> $OUT/pcie.err; grep '^{' $OUT/bench_pcie_n8.json | cut -c1-700
