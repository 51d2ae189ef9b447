#!/usr/bin/env bash
# GPU iteration on the decrypt kernels: parity tests (short timeout: a deadlocked persistent kernel must not hang the box),
# the at-scale timing of both kernels, then optionally the whole GPU suite.
# Usage: bash tools/gpu_dec.sh <tag> [full]
set -u
TAG=${1:-dec}; FULL=${2:-}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 240 python -m pytest tests/test_gpu_decrypt.py -m gpu -x -q > "$OUT/pytest_dec.log" 2>&1; rc=$?; echo "pytest decrypt rc=$rc" | tee -a "$OUT/pytest_dec.log"
tail -30 "$OUT/pytest_dec.log"
if [ $rc -eq 0 ]; then
  timeout 300 python tools/bench_decrypt.py > "$OUT/decrypt.json" 2> "$OUT/decrypt.err"; echo "bench_decrypt rc=$?"
  cat "$OUT/decrypt.json"; tail -5 "$OUT/decrypt.err"
fi
if [ -n "$FULL" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_decrypt.py > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rest rc=$?" | tee -a "$OUT/pytest_gpu.log"
  tail -8 "$OUT/pytest_gpu.log"
  timeout 600 python bench.py --steps 20 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"
  cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
fi
