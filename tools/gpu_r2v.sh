set -u
OUT=gpurun_out/r2v; mkdir -p $OUT
for rep in 1 2; do for ko in 0 64; do
  IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_DECRYPT_KNOCKOUT=$ko DEC_QUICK=1 timeout 300 python tools/bench_decrypt.py 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); k=r['kernels']['decrypt_tc_kernel']; print('ko=$ko', k)"
done; done
IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_DECRYPT_KNOCKOUT=64 timeout 300 python -m pytest tests/test_gpu_decrypt.py -m gpu -x -q 2>&1 | tail -3
