#!/usr/bin/env bash
# Round-2 checkpoint in one gpurun call: GPU parity tests, smoke, the default bench + reference arm, bench lines of the other BASELINE
# configurations, the ncu launch list and `ncu --set full` captures of the dominant kernels.  Usage: bash tools/gpu_round2.sh <tag>
set -u
TAG=${1:-round2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/host.txt"; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> "$OUT/host.txt"; free -g | head -2 >> "$OUT/host.txt"
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
tail -4 "$OUT/pytest_gpu.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/smoke.log"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "bench ref rc=$?"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"
for cfg in "n20:--neighbors 20" "n50:--neighbors 50" "s335_n20:--samples 335 --neighbors 20"; do
  tag=${cfg%%:*}; args=${cfg#*:}
  timeout 900 python bench.py $args --no-decrypt > "$OUT/bench_$tag.json" 2> "$OUT/bench_$tag.err"; echo "bench $tag rc=$?"
done
for f in bench bench_n20 bench_n50 bench_s335_n20; do python - "$OUT/$f.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]; print(sys.argv[1], {k:d[k] for k in ("value","ms_per_step")}, {k:r[k] for k in ("kernel","kernel_ms","frac")}, d.get("parity",{}).get("equal"), d.get("sustained"), (d.get("decrypt") or {}).get("roofline"))
except Exception as e:
    print("no bench line:", sys.argv[1], e)
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-parity --sustain 0 > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?"
for spec in "n5:--neighbors 5" "n50:--neighbors 50" "s335:--samples 335 --neighbors 20"; do
  tag=${spec%%:*}; args=${spec#*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:cloud_ring' -s 2 -c 1 -f -o "$OUT/prof_ring_$tag" \
      python bench.py $args --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-parity --no-decrypt --sustain 0 > "$OUT/ncu_full_$tag.log" 2>&1; echo "ncu full $tag rc=$?"
done
timeout 300 python tools/bench_decrypt.py > "$OUT/decrypt.json" 2> "$OUT/decrypt.err"; echo "bench_decrypt rc=$?"; cat "$OUT/decrypt.json"
DEC_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decrypt_pair -s 3 -c 1 -f -o "$OUT/prof_decrypt_pair" \
    python tools/bench_decrypt.py > "$OUT/ncu_full_decrypt.log" 2>&1; echo "ncu decrypt rc=$?"
timeout 600 python tools/population_bench.py > "$OUT/populations.json" 2> "$OUT/populations.err"; echo "population bench rc=$?"; tail -3 "$OUT/populations.json"
ls -la "$OUT"
