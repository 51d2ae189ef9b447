set -u
OUT=gpurun_out/r3p; mkdir -p $OUT
timeout 120 python tools/pair_check.py 2>&1 | tail -7
for cfg in "IDASH_B200_DECRYPT_KNOCKOUT=0" "IDASH_B200_DECRYPT_KNOCKOUT=1" "IDASH_B200_DECRYPT_KNOCKOUT=38" "IDASH_B200_DECRYPT_KNOCKOUT=25" "IDASH_B200_DECRYPT_KNOCKOUT=27" "IDASH_B200_DECRYPT_SLOTS=20" "IDASH_B200_DECRYPT_SLOTS=12"; do
  env IDASH_B200_USE_PROFILE_LIB=1 $cfg DEC_KERNELS=decrypt_pair_kernel DEC_QUICK=1 timeout 120 python tools/bench_decrypt.py 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); k=list(r['kernels'].values())[0]; print('$cfg', round(k['kernel_ms'],4), k['sample_matches_exact_oracle'])"
done
