set -u
OUT=gpurun_out/r3h; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_decrypt.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do DEC_QUICK=1 timeout 300 python tools/bench_decrypt.py 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); k=r['kernels']['decrypt_tc_kernel']; print('prod', round(k['kernel_ms'],4), k['sample_matches_exact_oracle'])"; done
for cfg in "IDASH_B200_DECRYPT_KNOCKOUT=0" "IDASH_B200_DECRYPT_KNOCKOUT=1" "IDASH_B200_DECRYPT_KNOCKOUT=25" "IDASH_B200_DECRYPT_KNOCKOUT=27" "IDASH_B200_DECRYPT_KNOCKOUT=6" "IDASH_B200_DECRYPT_KNOCKOUT=38"; do
  env IDASH_B200_USE_PROFILE_LIB=1 $cfg DEC_QUICK=1 timeout 300 python tools/bench_decrypt.py 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); k=r['kernels']['decrypt_tc_kernel']; print('$cfg', round(k['kernel_ms'],4), k['sample_matches_exact_oracle'])"
done
