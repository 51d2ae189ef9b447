set -u
OUT=gpurun_out/r3k; mkdir -p $OUT
timeout 600 python tools/ring_sweep.py --workloads 1004:5,1004:20,1004:50 --settings "456;1480;456;1480" --steps 20 --out $OUT/a.jsonl 2>&1 | tail -12
timeout 600 python tools/ring_sweep.py --workloads 335:20 --settings "328;1352;328;1352" --steps 20 --out $OUT/b.jsonl 2>&1 | tail -4
