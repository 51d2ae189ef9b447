#!/usr/bin/env bash
# round 2, call C: knock-out timings of the ring kernel (which stage bounds neighbors = 50 and NUM_REGIONS = 3)
set -u
OUT=gpurun_out/${1:-r2c}; mkdir -p $OUT
timeout 600 python tools/ring_sweep.py --workloads 1004:50 --settings "8;8,ko=1;8,ko=2;8,ko=4;8,ko=8;8,ko=6;8,ko=14;200,ko=2;200,ko=6;40,ko=6" --out $OUT/ko_n50.jsonl > $OUT/ko_n50.log 2>&1; cat $OUT/ko_n50.log
timeout 600 python tools/ring_sweep.py --workloads 335:20 --settings "8;8,ko=16;8,ko=2;8,ko=1;8,ko=4;8,ko=8;8,slices=16;8,ko=18" --out $OUT/ko_s335.jsonl > $OUT/ko_s335.log 2>&1; cat $OUT/ko_s335.log
timeout 600 python tools/ring_sweep.py --workloads 1004:5 --settings "8;8,ko=1;8,ko=2;8,ko=4;8,ko=8;8,ko=6" --out $OUT/ko_n5.jsonl > $OUT/ko_n5.log 2>&1; cat $OUT/ko_n5.log
