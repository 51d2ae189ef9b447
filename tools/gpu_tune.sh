#!/usr/bin/env bash
# Timing of the ring kernel under the IDASH_B200_TUNE experiment switches (+ parity tests under one setting).
set -u
OUT=gpurun_out/${1:-tune}; mkdir -p $OUT
for n in ${NEIGHBORS:-5}; do
 for tu in ${TUNES:-0 1 3 7 8 9 11 15}; do
  IDASH_B200_TUNE=$tu timeout 200 python bench.py --steps 10 --warmup 3 --kernel ring --neighbors $n --no-cpu-baseline --e2e-steps 1 > $OUT/b_${n}_$tu.json 2>$OUT/b_${n}_$tu.err
  python - $OUT/b_${n}_$tu.json $n $tu <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("n=%s tune=%s kernel_ms=%.4f match=%s"%(sys.argv[2],sys.argv[3],d["roofline"]["kernel_ms"],d["e2e"]["matches_device_path"]))
except Exception as e: print("fail",sys.argv[2:],e)
PY
 done
done
if [ -n "${TEST_TUNE:-}" ]; then
  IDASH_B200_TUNE=$TEST_TUNE timeout 600 python -m pytest tests/test_gpu_cloud.py -x -q > $OUT/pytest_tune.log 2>&1; tail -3 $OUT/pytest_tune.log
fi
