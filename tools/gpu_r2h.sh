#!/usr/bin/env bash
set -u
OUT=gpurun_out/${1:-r2h}; mkdir -p $OUT
timeout 900 python tools/ring_sweep.py --settings "328;329;330;332;333;335;360" --out $OUT/sweep.jsonl > $OUT/sweep.log 2>&1; echo "sweep rc=$?"
cat $OUT/sweep.log
