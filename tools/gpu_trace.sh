#!/usr/bin/env bash
# Per-tile timeline of one ring-kernel CTA (IDASH_B200_TRACE) under a tune / knockout setting.
# usage: gpu_trace.sh <tag> ; env: NEIGHBORS, TUNE, KNOCKOUT, CTAS
set -u
OUT=gpurun_out/${1:-trace}; mkdir -p $OUT
for cta in ${CTAS:-0 70}; do
  f=$OUT/trace_n${NEIGHBORS:-5}_t${TUNE:-0}_k${KNOCKOUT:-0}_$cta
  IDASH_B200_TUNE=${TUNE:-0} IDASH_B200_KNOCKOUT=${KNOCKOUT:-0} IDASH_B200_TRACE=$cta IDASH_B200_TRACE_FILE=$f.txt timeout 200 python bench.py --steps 3 --warmup 3 --kernel ring --neighbors ${NEIGHBORS:-5} --no-cpu-baseline --e2e-steps 1 > $f.json 2> $f.err
  python tools/trace_ring.py $f.txt > $f.tbl 2>&1; sed -n 30,42p $f.tbl; tail -1 $f.tbl
done
