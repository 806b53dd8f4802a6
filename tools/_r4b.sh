OUT=gpurun_out/r4b; mkdir -p $OUT
P="python tools/bench_path.py --reps 2 --only"
for D in 1 0; do
HG_HEXSRC_DIST=$D timeout 600 $P "c4 hex->rect linear exact f32" > $OUT/plain_$D.log 2>&1 && HG_HEXSRC_DIST=$D timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexsrc_linear_tma -s 2 -c 1 -o $OUT/prof_h2r_exact_dist$D $P "c4 hex->rect linear exact f32" > $OUT/ncu_$D.log 2>&1; echo "ncu dist=$D rc=$?"
done
