set -u
OUT=gpurun_out/mgt; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_host.py -m gpu -q -k "sharded or several_gpus" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
