#!/usr/bin/env bash
set -u
OUT=gpurun_out/${1:-r2r}; mkdir -p $OUT
timeout 900 python tools/ring_sweep.py --workloads 1004:50 --settings "${SETTINGS}" --out $OUT/sweep.jsonl > $OUT/sweep.log 2>&1; echo "sweep rc=$?"
python - <<PY
import json
for l in open("$OUT/sweep.jsonl"):
    r=json.loads(l); print(r['S'],r['n'],r['setting'],r['kernel_ms'],r['same_bits'])
PY
