set -u
timeout 300 python -m pytest tests/test_gpu_cloud.py -m gpu -x -q -k "slow_producers" 2>&1 | tail -3
IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_KNOCKOUT=64 timeout 120 python tools/race_probe2.py 400 300 2 2>&1 | grep "bad rows"
