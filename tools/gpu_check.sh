#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, micro-benchmarks, a bench run, an ncu launch list and one
# `ncu --set full` capture of the dominant kernel.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh [tag] [extra bench args...]
set -u
TAG=${1:-run}
shift || true
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/host.txt"; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> "$OUT/host.txt"; free -g | head -2 >> "$OUT/host.txt"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
tail -5 "$OUT/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/smoke.log"
tail -3 "$OUT/smoke.log"
if [ -x tools/ubench ]; then timeout 120 tools/ubench > "$OUT/ubench.log" 2>&1; cat "$OUT/ubench.log"; fi
timeout 900 python bench.py --steps 20 --warmup 3 "$@" > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"
cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 "$@" > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "bench ref rc=$?"
cat "$OUT/bench_ref.json"
# launch list of the same command (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline "$@" > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?"
# full capture of the dominant kernel (3 launches after the warm-up one)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cloud_eval -s 1 -c 2 -f -o "$OUT/prof_cloud" \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline "$@" > "$OUT/ncu_full.log" 2>&1; echo "ncu full rc=$?"
ls -la "$OUT"
