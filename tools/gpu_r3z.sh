set -u
OUT=gpurun_out/r3z; mkdir -p $OUT
timeout 300 compute-sanitizer --tool memcheck python tools/race_probe2.py 400 300 4 2>&1 | grep -v "^=========" | grep "bad rows"
timeout 300 compute-sanitizer --tool memcheck python tools/race_probe.py 400 300 6 2>&1 | grep -v "^=========" | tail -6
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_cloud.py -m gpu -x -q -k "ring or batched or outlier or multi_model or population or golden" > $OUT/memcheck_cloud.log 2>&1; echo "memcheck cloud rc=$?"; tail -4 $OUT/memcheck_cloud.log
timeout 600 python tools/ring_sweep.py --workloads 1004:5,1004:50,335:20 --settings "456;456" --steps 10 2>&1 | tail -6 | cut -c1-110
