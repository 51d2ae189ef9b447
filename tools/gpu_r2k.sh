#!/usr/bin/env bash
# round 2, call K: ncu --set full (source-level stall sampling) of the ring kernel, production build, n = 5 and n = 50
set -u
OUT=gpurun_out/${1:-r2k}; mkdir -p $OUT
for n in 5 50; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cloud_ring -s 2 -c 1 -f -o $OUT/prof_n$n \
    python bench.py --neighbors $n --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-parity --no-decrypt --sustain 0 > $OUT/ncu_n$n.log 2>&1; echo "ncu n=$n rc=$?"
done
timeout 300 python bench.py --no-cpu-baseline --no-parity --no-decrypt --sustain 0 > $OUT/bench_prod_n5.json 2>$OUT/bench.err; python -c "
import json; r=json.load(open('$OUT/bench_prod_n5.json')); print('prod lib n=5 kernel_ms', r['roofline']['kernel_ms'])"
ls -la $OUT
