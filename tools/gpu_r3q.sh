set -u
OUT=gpurun_out/r3q; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_decrypt.py -m gpu -x -q 2>&1 | tail -4
IDASH_B200_USE_PROFILE_LIB=1 timeout 600 python -m pytest tests/test_gpu_decrypt.py -m gpu -x -q -k many_groups 2>&1 | tail -3
for rep in 1 2 3; do for k in decrypt_pair_kernel decrypt_tc_kernel; do
  DEC_KERNELS=$k DEC_QUICK=1 timeout 120 python tools/bench_decrypt.py 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); k=list(r['kernels'].values())[0]; print('$k', round(k['kernel_ms'],4), round(k['min_ms'],4), k['sample_matches_exact_oracle'])"
done; done
