set -u
OUT=gpurun_out/r3v; mkdir -p $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_cloud.py -m gpu -x -q -k "ring or batched or outlier or multi_model or population or golden" > $OUT/memcheck_cloud.log 2>&1; echo "memcheck cloud rc=$?"; tail -4 $OUT/memcheck_cloud.log
