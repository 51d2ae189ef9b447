set -u
OUT=gpurun_out/r3e; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,temperature.gpu --format=csv
timeout 600 python tools/ring_sweep.py --workloads 1004:5,1004:50 --settings "456;456,chunks=8;456,chunks=7;456;456,chunks=8" --steps 10 --out $OUT/a.jsonl 2>&1 | tail -10
timeout 600 python tools/ring_sweep.py --workloads 335:20 --settings "328;328,chunks=12;328,chunks=11;328" --steps 10 --out $OUT/b.jsonl 2>&1 | tail -4
