set -u
OUT=gpurun_out/r3g; mkdir -p $OUT
timeout 600 python tools/ring_sweep.py --workloads 1004:5,1004:20,1004:50 --settings "456,extra=0;456,fill=2;456,fill=4;456,fill=6;456,fill=8;456,fill=12;456,extra=0;456,fill=4;456,fill=6" --steps 10 --out $OUT/a.jsonl 2>&1 | tail -27
timeout 600 python tools/ring_sweep.py --workloads 335:20 --settings "328,extra=0;328,fill=2;328,fill=4;328,fill=6;328,fill=8;328,fill=12;328,extra=0;328,fill=4;328,fill=6" --steps 10 --out $OUT/b.jsonl 2>&1 | tail -9
