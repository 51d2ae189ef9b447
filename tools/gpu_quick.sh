#!/usr/bin/env bash
# Quick GPU iteration: parity tests + short bench for the kernels named on the command line.
# Usage: bash tools/gpu_quick.sh <tag> [kernel ...]   (kernel = auto | imad | tensor)
set -u
TAG=${1:-quick}; shift || true
KERNELS=${*:-tensor}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout ${PYTEST_TIMEOUT:-400} python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
tail -25 "$OUT/pytest_gpu.log"
for k in $KERNELS; do
  for n in ${NEIGHBORS:-5}; do
    timeout 240 python bench.py --steps 20 --warmup 3 --kernel $k --neighbors $n --no-cpu-baseline --e2e-steps 2 > "$OUT/bench_${k}_n$n.json" 2> "$OUT/bench_${k}_n$n.err"
    echo "bench $k n=$n rc=$?"; python - "$OUT/bench_${k}_n$n.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]; print({k:d[k] for k in ("value","ms_per_step")}, {k:r[k] for k in ("kernel","kernel_ms","achieved","frac")}, d["e2e"]["matches_device_path"])
except Exception as e:
    print("no bench line:", e); print(open(sys.argv[1].replace(".json",".err")).read()[-2000:])
PY
  done
done
