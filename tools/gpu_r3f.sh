set -u
OUT=gpurun_out/r3f; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_cloud.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/ring_sweep.py --workloads 1004:5,1004:20,1004:50,335:20 --settings "456,extra=0;456;456,fill=0;456,fill=4;456,extra=0;456" --steps 10 --out $OUT/a.jsonl 2>&1 | tail -24
