set -u
OUT=gpurun_out/r3a; mkdir -p $OUT
for rep in 1 2; do for cfg in "IDASH_B200_DECRYPT_SLOTS=9 IDASH_B200_DECRYPT_BSTAGES=5" "IDASH_B200_DECRYPT_SLOTS=10 IDASH_B200_DECRYPT_BSTAGES=3" "IDASH_B200_DECRYPT_SLOTS=10 IDASH_B200_DECRYPT_BSTAGES=2" "IDASH_B200_DECRYPT_SLOTS=9 IDASH_B200_DECRYPT_BSTAGES=3"; do
  env IDASH_B200_USE_PROFILE_LIB=1 $cfg DEC_QUICK=1 timeout 300 python tools/bench_decrypt.py 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); k=r['kernels']['decrypt_tc_kernel']; print('$cfg', round(k['kernel_ms'],4), k['sample_matches_exact_oracle'])"
done; done
export IDASH_B200_USE_PROFILE_LIB=1
NEIGHBORS=50 TUNE=456 CTAS="0 70" bash tools/gpu_trace.sh r3a
for cta in 0 70; do
  f=$OUT/trace_s335_$cta
  IDASH_B200_TUNE=328 IDASH_B200_TRACE=$cta IDASH_B200_TRACE_FILE=$f.txt timeout 200 python bench.py --steps 3 --warmup 3 --kernel ring --samples 335 --neighbors 20 --no-cpu-baseline --no-parity --no-decrypt --sustain 0 --e2e-steps 1 > $f.json 2> $f.err
  python tools/trace_ring.py $f.txt > $f.tbl 2>&1; sed -n 30,42p $f.tbl; tail -1 $f.tbl
done
