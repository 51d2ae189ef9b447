#!/usr/bin/env bash
# Diagnosis run: store-pattern micro-benchmarks, knock-out timing and a per-tile trace of the ring kernel.
set -u
OUT=gpurun_out/${1:-diag}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 120 tools/ubench_store > $OUT/ubench_store.log 2>&1; cat $OUT/ubench_store.log
timeout 120 tools/ubench > $OUT/ubench.log 2>&1; cat $OUT/ubench.log
NEIGHBORS="5" KOS="0 1 2 4 8 3 6 7 15" bash tools/gpu_knock.sh ${1:-diag}/knock
NEIGHBORS="50" KOS="0 1 2 4" bash tools/gpu_knock.sh ${1:-diag}/knock
for cta in 0 70; do
  IDASH_B200_TRACE=$cta IDASH_B200_TRACE_FILE=$OUT/trace_$cta.txt timeout 200 python bench.py --steps 3 --warmup 3 --kernel ring --no-cpu-baseline --e2e-steps 1 > $OUT/trace_$cta.json 2> $OUT/trace_$cta.err
  python tools/trace_ring.py $OUT/trace_$cta.txt > $OUT/trace_$cta.tbl 2>&1; tail -3 $OUT/trace_$cta.tbl
done
