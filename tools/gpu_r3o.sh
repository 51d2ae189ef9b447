set -u
OUT=gpurun_out/r3o; mkdir -p $OUT
for g in 148 144 140 136 132 128 120 112 96 64; do
  env IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_DECRYPT_GRID=$g DEC_KERNELS=decrypt_pair_kernel DEC_QUICK=1 timeout 120 python tools/bench_decrypt.py 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); k=list(r['kernels'].values())[0]; print('grid $g', round(k['kernel_ms'],4), k['sample_matches_exact_oracle'])"
done
