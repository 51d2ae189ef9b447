set -u
OUT=gpurun_out/r3l; mkdir -p $OUT
for ko in 0 1 2 4 8; do
  IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_KNOCKOUT=$ko timeout 300 python bench.py --neighbors 5 --no-cpu-baseline --no-parity --no-decrypt --sustain 3 --e2e-steps 1 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); s=r['sustained']; print('n5 ko=$ko burst', round(r['roofline']['kernel_ms'],4), 'sustained', round(s['ms_per_step'],4), s['clocks'])"
done
for ko in 0 1; do
  IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_KNOCKOUT=$ko timeout 300 python bench.py --neighbors 50 --no-cpu-baseline --no-parity --no-decrypt --sustain 3 --e2e-steps 1 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); s=r['sustained']; print('n50 ko=$ko burst', round(r['roofline']['kernel_ms'],4), 'sustained', round(s['ms_per_step'],4), s['clocks'])"
done
