#!/usr/bin/env bash
# round 2, call O: ncu source-level profile of the ring kernel at 335 samples (NUM_REGIONS = 3) / neighbors = 20; bench lines of the production build
set -u
OUT=gpurun_out/${1:-r2o}; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cloud_ring -s 2 -c 1 -f -o $OUT/prof_s335 \
    python bench.py --samples 335 --neighbors 20 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-parity --no-decrypt --sustain 0 > $OUT/ncu_s335.log 2>&1; echo "ncu rc=$?"
for cfg in "--neighbors 5" "--neighbors 20" "--neighbors 50" "--samples 335 --neighbors 20"; do
  timeout 300 python bench.py $cfg --no-cpu-baseline --no-parity --no-decrypt --sustain 0 2>>$OUT/bench.err | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('$cfg', 'kernel_ms', round(r['roofline']['kernel_ms'],4), 'frac', round(r['roofline']['frac'],4))"
done
