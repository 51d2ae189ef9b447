set -u
OUT=gpurun_out/r2y; mkdir -p $OUT
DEC_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decrypt_tc -s 3 -c 1 -f -o $OUT/prof_dec python tools/bench_decrypt.py > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
ls -la $OUT
