set -u
OUT=gpurun_out/r3s; mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_decrypt.py -m gpu -x -q -k "pair and (matches_exact or edge_keys or many_groups)" > $OUT/memcheck_decrypt_pair.log 2>&1; echo "memcheck rc=$?"; tail -5 $OUT/memcheck_decrypt_pair.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - $OUT/bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["parity"]["equal"])
print(json.dumps(d["decrypt"])[:1500])
PY
timeout 300 python tools/bench_decrypt.py > $OUT/decrypt.json 2> $OUT/decrypt.err; cat $OUT/decrypt.json
DEC_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decrypt_pair -s 3 -c 1 -f -o $OUT/prof_decrypt_pair python tools/bench_decrypt.py > $OUT/ncu_full_decrypt.log 2>&1; echo "ncu decrypt rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
