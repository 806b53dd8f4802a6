mkdir -p gpurun_out/s10
run() {
python tools/bench_conv.py --batch 64 --cin 3 --cout 32 --hw 128 --dtypes f32f32 --only fwd,wgrad --reps 20 2>&1 | grep '"op"' | sed "s/^/$1 L1 /"
python tools/bench_conv.py --batch 64 --cin 32 --cout 64 --hw 64 --w 63 --dtypes f32f32 --reps 20 2>&1 | grep '"op"' | sed "s/^/$1 L2 /"
python tools/bench_conv.py --batch 64 --cin 64 --cout 128 --hw 32 --w 31 --dtypes f32f32 --reps 20 2>&1 | grep '"op"' | sed "s/^/$1 L3 /"
}
timeout 900 python -m pytest tests/test_gpu_hexframes.py tests/test_gpu_baseline_sizes.py tests/test_zz_hexconvmodule_variants.py -m gpu -x -q -k "conv or Conv or c3 or c5 or C3 or C5 or hexcnn or module" > gpurun_out/s10/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s10/pytest.log
run auto > gpurun_out/s10/ab.log 2>&1
python tools/bench_conv.py --dtypes f32f32 --reps 30 2>&1 | grep '"op"' | sed "s/^/auto C3 /" >> gpurun_out/s10/ab.log
python tools/bench_conv.py --batch 64 --cin 128 --cout 128 --hw 32 --w 32 --dtypes f32f32 --reps 20 2>&1 | grep '"op"' | sed "s/^/auto c128w32 /" >> gpurun_out/s10/ab.log
for k in 0 1; do timeout 300 python tests/stress/wgrad_cfg.py 2>&1 | grep -c BAD; done > gpurun_out/s10/stress.log 2>&1; echo "stress BAD counts:" $(tr '\n' ' ' < gpurun_out/s10/stress.log)
python tools/hexcnn_ddp.py --autocast --graph --steps 50 2>&1 | tail -1 | cut -c1-200
python tools/hexcnn_ddp.py --autocast --steps 50 2>&1 | tail -1 | cut -c1-200
grep -v config gpurun_out/s10/ab.log | python -c "
import sys,json
for l in sys.stdin:
    if '{' not in l: continue
    tag,js=l.split('{',1); d=json.loads('{'+js)
    print(tag, d['op'], d['ms'])
"
