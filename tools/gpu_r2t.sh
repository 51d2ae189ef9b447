set -u
OUT=gpurun_out/r2t; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
bash tools/gpu_ab.sh r2t gpurun_ab/lib_4bcb5c9.so idash2019_2_b200/lib/libidash_b200.so
