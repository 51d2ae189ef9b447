#!/usr/bin/env bash
# round 2, call D: TMEM load micro-benchmark, wide-load epilogue sweep
set -u
OUT=gpurun_out/${1:-r2d}; mkdir -p $OUT
timeout 120 ./tools/tmem_ld_rate > $OUT/tmem_ld_rate.txt 2>&1; cat $OUT/tmem_ld_rate.txt
timeout 200 python -m pytest tests/test_gpu_cloud.py -m gpu -x -q > $OUT/pytest_cloud.log 2>&1; echo "pytest cloud rc=$?"; tail -3 $OUT/pytest_cloud.log
timeout 900 python tools/ring_sweep.py --settings "8;264;328;72;264,ko=2;8,ko=2" --out $OUT/sweep.jsonl --trace-dir $OUT/traces > $OUT/sweep.log 2>&1; echo "sweep rc=$?"
cat $OUT/sweep.log
