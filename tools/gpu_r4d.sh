set -u
OUT=gpurun_out/r4d; mkdir -p $OUT
timeout 60 python bench.py --neighbors 5 --no-cpu-baseline --no-decrypt --sustain 0 --e2e-steps 1 2>&1 | tail -1 | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('n5 prod', round(r['roofline']['kernel_ms'],4), r['parity']['equal'])" || exit 1
timeout 120 compute-sanitizer --tool memcheck python tools/race_probe2.py 400 300 3 2>&1 | grep -v "^=========" | grep "bad rows"
for cfg in "--neighbors 5" "--neighbors 20" "--neighbors 50" "--samples 335 --neighbors 20"; do
  for L in gpurun_ab/lib_nogate.so idash2019_2_b200/lib/libidash_b200.so gpurun_ab/lib_nogate.so idash2019_2_b200/lib/libidash_b200.so; do
    IDASH_B200_LIB=$L timeout 60 python bench.py $cfg --no-cpu-baseline --no-parity --no-decrypt --sustain 0 --e2e-steps 1 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('$cfg', '$L', 'kernel_ms', round(r['roofline']['kernel_ms'],4))"
  done; done
timeout 600 python -m pytest tests/test_gpu_cloud.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -2
