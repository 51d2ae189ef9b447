#!/usr/bin/env bash
# round 2, call A: GPU parity tests (incl. the full-size ones), smoke, ring-kernel schedule sweep with traces
set -u
OUT=gpurun_out/${1:-r2a}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; free -g | head -2 >> $OUT/host.txt
timeout 240 python -m pytest tests/test_gpu_cloud.py -m gpu -x -q > $OUT/pytest_cloud.log 2>&1; rc=$?; echo "pytest cloud rc=$rc"; tail -5 $OUT/pytest_cloud.log
if [ $rc -ne 0 ]; then tail -60 $OUT/pytest_cloud.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 1200 python tools/ring_sweep.py --settings "${SETTINGS:-8;40;72;104;104,slots=9;8,slots=9}" --out $OUT/sweep.jsonl --trace-dir $OUT/traces > $OUT/sweep.log 2>&1; echo "sweep rc=$?"
cat $OUT/sweep.log | tail -40
for f in $OUT/traces/*n50*.txt $OUT/traces/*S335*.txt; do python tools/trace_ring.py $f > ${f%.txt}.tbl 2>&1; echo $f; tail -1 ${f%.txt}.tbl; done
