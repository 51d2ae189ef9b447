#!/usr/bin/env bash
# round 2, call I: all GPU tests with the new host layer, the binaries end to end at iDASH scale (model cache cold / warm), bench lines
set -u
OUT=gpurun_out/${1:-r2i}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 900 python tools/pipeline_at_scale.py --decrypt ${GPUS:+--gpus $GPUS} > $OUT/pipeline.json 2> $OUT/pipeline.err; echo "pipeline rc=$?"; tail -3 $OUT/pipeline.err
python - <<PY
import json
r = json.loads(open("$OUT/pipeline.json").read().strip().splitlines()[-1])
for k, v in r.items():
    if isinstance(v, dict) and "phases" in v:
        print(k, {x: y for x, y in v.items() if x != "phases"})
        for p in v["phases"]: print("   ", p)
    elif k.startswith("ref_"): print(k, v)
PY
timeout 600 python bench.py > $OUT/bench_n5.json 2> $OUT/bench_n5.err; echo "bench rc=$?"; cat $OUT/bench_n5.json
