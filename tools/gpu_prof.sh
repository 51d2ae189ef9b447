#!/usr/bin/env bash
# ncu captures of the dominant kernel under bench.py: a launch list and one --set full capture.
# Usage: bash tools/gpu_prof.sh <tag> <kernel-regex> [bench args...]
set -u
TAG=$1; KRE=$2; shift 2
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline "$@" > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 2 -f -o "$OUT/prof" \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline "$@" > "$OUT/ncu_full.log" 2>&1; echo "ncu full rc=$?"
ls -la "$OUT"
