set -u
OUT=gpurun_out/r3b; mkdir -p $OUT
timeout 600 python tools/ring_sweep.py --workloads 1004:50 --settings "456;488;456,ko=1;456,ko=2;456,ko=4;456,ko=6;456,ko=8;456,ko=10;456,ko=14;456,ko=32;456,ko=34" --steps 10 --out $OUT/n50.jsonl 2>&1 | tail -12
timeout 600 python tools/ring_sweep.py --workloads 335:20 --settings "328;360;328,ko=1;328,ko=2;328,ko=4;328,ko=6;328,ko=8;328,ko=10;328,ko=16;328,ko=18;328,ko=32;328,ko=34" --steps 10 --out $OUT/s335.jsonl 2>&1 | tail -13
timeout 600 python tools/ring_sweep.py --workloads 1004:5 --settings "456;456,ko=1;456,ko=2;456,ko=4;456,ko=8;456,ko=32" --steps 10 --out $OUT/n5.jsonl 2>&1 | tail -7
