#!/usr/bin/env bash
set -u
OUT=gpurun_out/${1:-r2g}; mkdir -p $OUT
timeout 900 python tools/ring_sweep.py --workloads "1004:5,1004:50,335:20" --settings "328;328,ko=4;328,ko=2" --steps 10 --out $OUT/sweep.jsonl --trace-dir $OUT/traces > $OUT/sweep.log 2>&1; echo "sweep rc=$?"
cat $OUT/sweep.log
for f in $OUT/traces/*.txt; do python tools/trace_ring.py $f > ${f%.txt}.tbl 2>&1; echo $f; sed -n 45,52p ${f%.txt}.tbl; tail -1 ${f%.txt}.tbl; done
