set -u
OUT=gpurun_out/r3m; mkdir -p $OUT
timeout 120 python tools/pair_check.py 2>&1 | tail -12; echo "rc=$?"
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv | tail -1
