#!/usr/bin/env bash
# multi-GPU call (gpurun --gpus N): multi-GPU tests, the in-process sharded evaluation, the torchrun bench line at N GPUs
set -u
N=${1:-2}; OUT=gpurun_out/mg_r2_$N; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_host.py -m gpu -x -q -k "sharded or several_gpus" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 600 python tools/multi_gpu_bench.py --gpus 1,2,4,8 > $OUT/multi_device.jsonl 2> $OUT/multi_device.err; echo "multi rc=$?"; cat $OUT/multi_device.jsonl | cut -c1-220
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > $OUT/bench_$N.json 2> $OUT/bench_$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    r=json.loads(open("$OUT/bench_$N.json").read().strip().splitlines()[-1])
    print({k:r[k] for k in ("value","n_gpus","ms_per_step")}, r["roofline"]["frac"], r["parity"], r.get("nvlink"), {k:r["e2e"][k] for k in ("value","ms_per_step","d2h_gbs_per_gpu","bare_d2h_gbs_per_gpu")}, (r.get("decrypt") or {}).get("sharded"))
except Exception as e:
    print("bench line unreadable", e); print(open("$OUT/bench_$N.err").read()[-3000:])
PY
