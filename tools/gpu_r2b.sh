#!/usr/bin/env bash
# round 2, call B: GPU tests, weight-stationary MMA sweep, layout compile timing on the box
set -u
OUT=gpurun_out/${1:-r2b}; mkdir -p $OUT
nproc > $OUT/host.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|NUMA" >> $OUT/host.txt; free -g | head -2 >> $OUT/host.txt
cat /sys/kernel/mm/transparent_hugepage/enabled >> $OUT/host.txt
timeout 240 python -m pytest tests/test_gpu_cloud.py -m gpu -x -q > $OUT/pytest_cloud.log 2>&1; rc=$?; echo "pytest cloud rc=$rc"; tail -5 $OUT/pytest_cloud.log
if [ $rc -ne 0 ]; then tail -60 $OUT/pytest_cloud.log; fi
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 1200 python tools/ring_sweep.py --settings "${SETTINGS:-8;136;200;232;168}" --out $OUT/sweep.jsonl --trace-dir $OUT/traces > $OUT/sweep.log 2>&1; echo "sweep rc=$?"
cat $OUT/sweep.log | tail -40
for f in $OUT/traces/*n50*.txt; do python tools/trace_ring.py $f > ${f%.txt}.tbl 2>&1; echo $f; tail -1 ${f%.txt}.tbl; done
for t in 1 8 16 64; do for n in 5 50; do echo "threads $t n $n"; PYTHONPATH=. IDASH_B200_THREADS=$t IDASH_B200_LAYOUT_TIMING=1 python tools/layout_time.py $n 2>&1 | tail -11; done; done > $OUT/layout_time.log 2>&1
grep -E "threads|total" $OUT/layout_time.log
