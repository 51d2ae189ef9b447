set -u
OUT=gpurun_out/r3i; mkdir -p $OUT
for cfg in "IDASH_B200_DECRYPT_SLOTS=9 IDASH_B200_DECRYPT_BSTAGES=5" "IDASH_B200_DECRYPT_SLOTS=10 IDASH_B200_DECRYPT_BSTAGES=3" "IDASH_B200_DECRYPT_SLOTS=10 IDASH_B200_DECRYPT_BSTAGES=2" "IDASH_B200_DECRYPT_SLOTS=9 IDASH_B200_DECRYPT_BSTAGES=4" "IDASH_B200_DECRYPT_SLOTS=9 IDASH_B200_DECRYPT_BSTAGES=3" "IDASH_B200_DECRYPT_KNOCKOUT=64"; do
  env IDASH_B200_USE_PROFILE_LIB=1 $cfg DEC_QUICK=1 timeout 300 python tools/bench_decrypt.py 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); k=r['kernels']['decrypt_tc_kernel']; print('$cfg', round(k['kernel_ms'],4), k['sample_matches_exact_oracle'])"
done
DEC_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decrypt_tc -s 3 -c 1 -f -o $OUT/prof_dec python tools/bench_decrypt.py > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
