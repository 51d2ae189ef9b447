#!/usr/bin/env bash
# round 2, call J: packed burst epilogue (EPI 2), try_wait MMA waits, weight-stationary combinations
set -u
OUT=gpurun_out/${1:-r2j}; mkdir -p $OUT
timeout 240 python -m pytest tests/test_gpu_cloud.py -m gpu -x -q > $OUT/pytest_cloud.log 2>&1; echo "pytest cloud rc=$?"; tail -3 $OUT/pytest_cloud.log
timeout 1200 python tools/ring_sweep.py --settings "${SETTINGS:-456;456,ko=32;328;328,ko=32}" --out $OUT/sweep.jsonl --trace-dir $OUT/traces > $OUT/sweep.log 2>&1; echo "sweep rc=$?"
python - <<PY
import json
from collections import defaultdict
t=defaultdict(dict)
for l in open("$OUT/sweep.jsonl"):
    r=json.loads(l); t[r['setting']][(r['S'],r['n'])]=(r['kernel_ms'], r['same_bits'])
for s,v in t.items(): print(f"{s:22s}", ' '.join(f"{k[0]}:{k[1]}={x[0]:.4f}{'' if x[1] else '!'}" for k,x in v.items()))
PY
for f in $OUT/traces/*n50*.txt; do python tools/trace_ring.py $f > ${f%.txt}.tbl 2>&1; echo $f; tail -1 ${f%.txt}.tbl; done
