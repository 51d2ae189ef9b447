set -u
OUT=gpurun_out/r3d; mkdir -p $OUT
timeout 600 python tools/ring_sweep.py --workloads 1004:50 --settings "456;12776;16872;20968;25064;29160;21992;26088;456,ko=14;20968,ko=14;25064,ko=14;21992,ko=14;26088,ko=14" --steps 10 --out $OUT/n50.jsonl 2>&1 | tail -14
timeout 600 python tools/ring_sweep.py --workloads 1004:20,335:20 --settings "456;20968;25064;21992;328;20840;24936" --steps 10 --out $OUT/n20.jsonl 2>&1 | tail -15
