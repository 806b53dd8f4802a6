OUT=gpurun_out/r4d; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_resample.py tests/test_gpu_baseline_sizes.py tests/test_gpu_stream_kernel.py tests/test_zz_numba_twin.py -x -q -m gpu > $OUT/pytest_h2r.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_h2r.log
timeout 900 python tools/sweep_kernels.py --what dist --reps 10 > $OUT/sweep_dist.jsonl 2> $OUT/sweep_dist.err; echo "sweep rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r4d/sweep_dist.jsonl'):
    d=json.loads(l); print(d['config'], d['variant'], d['env'], d['ms'], d['hbm_frac'])
PY
