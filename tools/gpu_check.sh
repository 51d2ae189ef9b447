#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, micro-benchmarks, a short bench and an ncu launch list.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh [tag]
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/host.txt"; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> "$OUT/host.txt"; free -g | head -2 >> "$OUT/host.txt"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
tail -5 "$OUT/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/smoke.log"
tail -3 "$OUT/smoke.log"
[ -x tools/ubench ] && timeout 120 tools/ubench > "$OUT/ubench.log" 2>&1; cat "$OUT/ubench.log"
timeout 900 python bench.py --steps 20 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"
cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
