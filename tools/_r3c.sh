OUT=gpurun_out/r3c; mkdir -p $OUT
P="python tools/bench_path.py --reps 2 --only"
timeout 600 $P "c2 rect->hex bilinear fast" > $OUT/plain_stream.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:rect2hex_stream -s 2 -c 1 -o $OUT/prof_stream $P "c2 rect->hex bilinear fast" > $OUT/ncu_stream.log 2>&1; echo "ncu stream rc=$?"
HG_HEXSRC_SHARE=64 timeout 600 $P "c4 hex->rect linear exact" > $OUT/plain_hexsrc_exact.log 2>&1 && HG_HEXSRC_SHARE=64 timeout 900 ncu --set full --clock-control none --import-source on -k regex:hexsrc_linear_tma -s 2 -c 1 -o $OUT/prof_hexsrc_exact $P "c4 hex->rect linear exact" > $OUT/ncu_hexsrc_exact.log 2>&1; echo "ncu hexsrc exact rc=$?"
cat $OUT/plain_stream.log $OUT/plain_hexsrc_exact.log | cut -c1-200
ls -la $OUT
