set -u
OUT=gpurun_out/r3c; mkdir -p $OUT
timeout 600 python tools/ring_sweep.py --workloads 1004:50 --settings "456,ko=14;456,ko=14,bchunks=14;456,ko=14,bchunks=16;456,ko=14,prefetch=0;488,ko=14;328,ko=14;460,ko=14;1480,ko=14;392,ko=14;456,ko=10;488,ko=10;456,ko=6;488,ko=6;456;460;1480" --steps 10 --out $OUT/n50.jsonl 2>&1 | tail -17
