#!/usr/bin/env bash
# 8-GPU call: multi-GPU tests (4 / 8 GPUs), in-process sharded evaluation, torchrun bench at N = 8, the cloud binary with IDASH_GPUS
set -u
OUT=gpurun_out/mg_r2_8; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc >> $OUT/topo.txt
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_host.py -m gpu -q -k "sharded or several_gpus" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 600 python tools/multi_gpu_bench.py --gpus 1,2,4,8 > $OUT/multi_device.jsonl 2> $OUT/multi_device.err; echo "multi rc=$?"; cut -c1-200 $OUT/multi_device.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --no-decrypt > $OUT/bench_8.json 2> $OUT/bench_8.err; echo "bench rc=$?"
python - <<PY
import json
try:
    r=json.loads(open("$OUT/bench_8.json").read().strip().splitlines()[-1])
    print({k:r[k] for k in ("value","n_gpus","ms_per_step")}, r["roofline"]["frac"], r["parity"]["equal"], r.get("nvlink"), {k:r["e2e"][k] for k in ("value","ms_per_step","d2h_gbs_per_gpu","bare_d2h_gbs_per_gpu")})
except Exception as e:
    print("bench line unreadable", e); print(open("$OUT/bench_8.err").read()[-3000:])
PY
timeout 900 python tools/pipeline_at_scale.py --gpus 0,1,2,3,4,5,6,7 > $OUT/pipeline.json 2> $OUT/pipeline.err; echo "pipeline rc=$?"
python - <<PY
import json
r = json.loads(open("$OUT/pipeline.json").read().strip().splitlines()[-1])
for k, v in r.items():
    if isinstance(v, dict) and "phases" in v:
        print(k, {x: y for x, y in v.items() if x != "phases"})
        for p in v["phases"]: print("   ", p)
    elif k.startswith("ref_"): print(k, v)
PY
