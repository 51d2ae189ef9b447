set -u
OUT=gpurun_out/r3u; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_cloud.py tests/test_gpu_scale.py tests/test_host.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do timeout 600 python bench.py --no-decrypt --no-cpu-baseline --sustain 0 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); e=r['e2e']; print('e2e ms', round(e['ms_per_step'],3), 'bare d2h ms', round(e['bare_d2h_ms_per_step'],3), 'd2h gbs', round(e['d2h_gbs_per_gpu'],2), e['matches_device_path'], r['parity']['equal'])"; done
