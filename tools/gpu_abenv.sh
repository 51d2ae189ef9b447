#!/usr/bin/env bash
# A/B/C... timing of environment settings of the PROFILING build on one box, separate processes, two rounds:
#   bash tools/gpu_abenv.sh <tag> "<bench args>" "ENV=V ENV2=V2" "ENV=V" ...      ("-" = no variables)
set -u
OUT=gpurun_out/$1; mkdir -p $OUT; CFG=$2; shift 2
for rep in 1 2; do for V in "$@"; do
  E=""; [ "$V" != "-" ] && E="$V"
  env IDASH_B200_USE_PROFILE_LIB=1 $E timeout 300 python bench.py $CFG --no-cpu-baseline --no-parity --no-decrypt --sustain 0 --e2e-steps 1 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('$CFG |', '$V', '| kernel_ms', round(r['roofline']['kernel_ms'],4))"
done; done
