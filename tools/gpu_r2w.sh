set -u
OUT=gpurun_out/r2w; mkdir -p $OUT
for ko in 64 65 66 68 72 80 96 71 89 121 103; do
  IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_DECRYPT_KNOCKOUT=$ko DEC_QUICK=1 timeout 300 python tools/bench_decrypt.py 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); k=r['kernels']['decrypt_tc_kernel']; print('ko=$ko', round(k['kernel_ms'],4))"
done
