set -u
OUT=gpurun_out/r3j; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_cloud.py tests/test_gpu_scale.py tests/test_gpu_decrypt.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/ring_sweep.py --workloads 335:20 --settings "344;328;344;328;328,ko=16" --steps 10 --out $OUT/a.jsonl 2>&1 | tail -5
for rep in 1 2; do DEC_QUICK=1 timeout 300 python tools/bench_decrypt.py 2>>$OUT/err.log | python -c "
import json,sys; r=json.loads(sys.stdin.read()); k=r['kernels']['decrypt_tc_kernel']; print('decrypt prod', round(k['kernel_ms'],4), k['sample_matches_exact_oracle'])"; done
