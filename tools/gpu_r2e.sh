#!/usr/bin/env bash
# round 2, call E: burst epilogue sweep
set -u
OUT=gpurun_out/${1:-r2e}; mkdir -p $OUT
timeout 900 python tools/ring_sweep.py --settings "${SETTINGS:-328;584;616;584,ko=2;584,slots=9}" --out $OUT/sweep.jsonl --trace-dir $OUT/traces > $OUT/sweep.log 2>&1; echo "sweep rc=$?"
cat $OUT/sweep.log
IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_TUNE=584 timeout 300 python -m pytest tests/test_gpu_cloud.py -m gpu -x -q > $OUT/pytest_cloud_584.log 2>&1; echo "pytest cloud (tune 584) rc=$?"; tail -3 $OUT/pytest_cloud_584.log
