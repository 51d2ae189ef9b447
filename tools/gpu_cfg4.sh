#!/usr/bin/env bash
# BASELINE configs[3]: population-stratified models, 335 samples per population (NUM_REGIONS = 3, REGION_SIZE = 341), neighbors = 20
set -u
TAG=${1:-cfg4}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for k in ${KERNELS:-auto imad tile}; do
  timeout 300 python bench.py --steps 20 --warmup 3 --samples 335 --neighbors 20 --kernel $k --no-cpu-baseline --e2e-steps 2 > "$OUT/bench_${k}.json" 2> "$OUT/bench_${k}.err"
  echo "cfg4 $k rc=$?"; python - "$OUT/bench_${k}.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]; print({k:d[k] for k in ("value","ms_per_step")}, {k:r[k] for k in ("kernel","kernel_ms","achieved","frac","algorithmic_bytes")}, d["e2e"]["matches_device_path"], d["config"]["in_ct_per_gpu_batch"], d["config"]["out_ct_per_gpu_batch"])
except Exception as e:
    print("no bench line:", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
