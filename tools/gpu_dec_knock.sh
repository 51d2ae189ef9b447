#!/usr/bin/env bash
# Knock-out timing of the tensor-core decrypt kernel (which stage bounds it?) + one ncu --set full capture.
# Usage: bash tools/gpu_dec_knock.sh <tag>
set -u
TAG=${1:-deck}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for ko in ${KOS:-0 1 2 4 6 8 16 24 32 35 63}; do
  IDASH_B200_DECRYPT_KNOCKOUT=$ko DEC_QUICK=1 timeout 120 python tools/bench_decrypt.py > "$OUT/ko_$ko.json" 2> "$OUT/ko_$ko.err"
  python - "$OUT/ko_$ko.json" $ko <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])["kernels"]["decrypt_tc_kernel"]
    print("knockout", sys.argv[2], "kernel_ms %.4f" % d["kernel_ms"], "ok" if d["sample_matches_exact_oracle"] else "")
except Exception as e:
    print("knockout", sys.argv[2], "failed", e)
PY
done
if [ "${NCU:-1}" != "0" ]; then
  DEC_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decrypt_tc -s 2 -c 1 -f -o "$OUT/prof_decrypt_tc" \
      python tools/bench_decrypt.py > "$OUT/ncu_full.log" 2>&1; echo "ncu rc=$?"
fi
ls -la "$OUT"
