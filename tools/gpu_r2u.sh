set -u
OUT=gpurun_out/r2u; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_cloud.py -m gpu -x -q > $OUT/pytest_cloud.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_cloud.log
timeout 600 python tools/population_bench.py > $OUT/population.json 2> $OUT/population.err; echo "pop rc=$?"; cat $OUT/population.json; tail -3 $OUT/population.err
