#!/usr/bin/env bash
# A/B timing of two builds of the library on the same box: bash tools/gpu_ab.sh <tag> <libA> <libB> ; alternates A B A B per workload
set -u
OUT=gpurun_out/${1:-ab}; mkdir -p $OUT
A=$2; B=$3
for cfg in "--neighbors 5" "--neighbors 20" "--neighbors 50" "--samples 335 --neighbors 20"; do
  for rep in 1 2; do for L in $A $B; do
    IDASH_B200_LIB=$L timeout 300 python bench.py $cfg --no-cpu-baseline --no-parity --no-decrypt --sustain 0 --e2e-steps 1 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('$cfg', '$L', 'kernel_ms', round(r['roofline']['kernel_ms'],4))"
  done; done
done
