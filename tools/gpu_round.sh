#!/usr/bin/env bash
# One gpurun call for a round checkpoint: GPU parity tests, smoke, the default bench + reference arm, a kernel x
# neighbors sweep, the ncu launch list and `ncu --set full` captures of the dominant kernels.
# Usage (repo root on the GPU box): bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-round}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/host.txt"; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> "$OUT/host.txt"; free -g | head -2 >> "$OUT/host.txt"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
tail -5 "$OUT/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/smoke.log"
tail -3 "$OUT/smoke.log"
timeout 900 python bench.py --steps 20 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"
cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "bench ref rc=$?"
cat "$OUT/bench_ref.json"
for n in ${NEIGHBORS:-5 20 50}; do
  for k in ${KERNELS:-imad tile ring}; do
    timeout 300 python bench.py --steps 20 --warmup 3 --kernel $k --neighbors $n --no-cpu-baseline --e2e-steps 2 > "$OUT/bench_${k}_n$n.json" 2> "$OUT/bench_${k}_n$n.err"
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
# launch list of the default bench command (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?"
# full captures of the dominant kernel (2nd launch on), one per kernel variant at n=5 and the ring kernel at n=50
for spec in ${NCU_SPECS:-auto:5 imad:5 ring:50}; do
  k=${spec%%:*}; n=${spec##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:cloud_(eval|tc|ring)' -s 1 -c 1 -f -o "$OUT/prof_${k}_n$n" \
      python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --kernel $k --neighbors $n > "$OUT/ncu_full_${k}_n$n.log" 2>&1; echo "ncu full $k n=$n rc=$?"
done
ls -la "$OUT"
# the decrypt kernels at iDASH scale + one full capture of the tensor-core one
timeout 300 python tools/bench_decrypt.py > "$OUT/decrypt.json" 2> "$OUT/decrypt.err"; echo "bench_decrypt rc=$?"; cat "$OUT/decrypt.json"
DEC_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decrypt_tc -s 2 -c 1 -f -o "$OUT/prof_decrypt_tc" \
    python tools/bench_decrypt.py > "$OUT/ncu_full_decrypt.log" 2>&1; echo "ncu decrypt rc=$?"
# BASELINE configs[3] geometry (335 samples, NUM_REGIONS = 3, neighbors = 20)
timeout 300 python bench.py --steps 20 --warmup 3 --samples 335 --neighbors 20 --no-cpu-baseline --e2e-steps 2 > "$OUT/bench_cfg4.json" 2> "$OUT/bench_cfg4.err"; echo "bench cfg4 rc=$?"
ls -la "$OUT"
