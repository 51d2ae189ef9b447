set -u
OUT=gpurun_out/r4g; mkdir -p $OUT
for cfg in "--neighbors 5" "--neighbors 50"; do
  for rep in 1 2; do for L in idash2019_2_b200/lib/libidash_b200.so gpurun_ab/lib_stg1.so gpurun_ab/lib_stg2.so gpurun_ab/lib_stg4.so; do
    IDASH_B200_LIB=$L timeout 60 python bench.py $cfg --no-cpu-baseline --no-parity --no-decrypt --sustain 0 --e2e-steps 1 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('$cfg', '$L', 'kernel_ms', round(r['roofline']['kernel_ms'],4))"
  done; done; done
