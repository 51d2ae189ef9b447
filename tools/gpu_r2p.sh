#!/usr/bin/env bash
# round 2, call P: full GPU test suite + bench lines of the production build at the four workloads
set -u
OUT=gpurun_out/${1:-r2p}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
for cfg in "--neighbors 5" "--neighbors 20" "--neighbors 50" "--samples 335 --neighbors 20"; do
  tag=$(echo $cfg | tr -d ' -'); timeout 300 python bench.py $cfg --no-cpu-baseline --no-decrypt --sustain 0 > $OUT/bench_$tag.json 2>>$OUT/bench.err; python -c "
import json,sys; r=json.load(open('$OUT/bench_$tag.json')); print('$cfg', 'kernel_ms', round(r['roofline']['kernel_ms'],4), 'frac', round(r['roofline']['frac'],4), 'parity', r['parity'].get('equal'), r['parity'].get('checked_words'))"
done
