#!/usr/bin/env bash
# Knock-out timing of the ring kernel: which stage bounds it? (results are wrong when a stage is disabled)
set -u
OUT=gpurun_out/${1:-knock}; mkdir -p $OUT
for n in ${NEIGHBORS:-5 50}; do
 for ko in ${KOS:-0 1 2 4 6 8 3 7 15}; do
  IDASH_B200_KNOCKOUT=$ko timeout 200 python bench.py --steps 10 --warmup 3 --kernel ring --neighbors $n --no-cpu-baseline --e2e-steps 1 > $OUT/b_${n}_$ko.json 2>$OUT/b_${n}_$ko.err
  python - $OUT/b_${n}_$ko.json $n $ko <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("n=%s knockout=%s kernel_ms=%.4f"%(sys.argv[2],sys.argv[3],d["roofline"]["kernel_ms"]))
except Exception as e: print("fail",sys.argv[2:],e)
PY
 done
done
