mkdir -p gpurun_out/s6
L=hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200/lib/libhygrid_b200.so
cp $L /tmp/new.so
run() {
python tools/bench_conv.py --dtypes f32f32 --reps 30 2>&1 | grep '"op"' | sed "s/^/$1 C3 /"
python tools/bench_conv.py --dtypes bf16bf16 --reps 30 --only wgrad 2>&1 | grep '"op"' | sed "s/^/$1 C3bf16 /"
python tools/hexcnn_ddp.py --autocast --steps 50 2>&1 | grep -i "ms\|img" | tail -2 | sed "s/^/$1 C5eager /"
python tools/hexcnn_ddp.py --autocast --graph --steps 50 2>&1 | grep -i "ms\|img" | tail -2 | sed "s/^/$1 C5graph /"
}
for i in 1 2; do
run new$i >> gpurun_out/s6/ab.log 2>&1
cp tools/_ab/libold.so $L
run old$i >> gpurun_out/s6/ab.log 2>&1
cp /tmp/new.so $L
done
timeout 900 python -m pytest tests/test_gpu_hexframes.py tests/test_gpu_baseline_sizes.py tests/test_zz_hexconvmodule_variants.py -m gpu -x -q -k "conv or Conv or c3 or c5 or C3 or C5 or hexcnn or module" > gpurun_out/s6/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s6/pytest.log
cut -c1-220 gpurun_out/s6/ab.log | grep -v config
