timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --sustain 0 2>/dev/null | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print(round(r['roofline']['kernel_ms'],4), round(r['roofline']['frac'],4), r['parity']['equal'], r['decrypt']['roofline']['kernel'], round(r['decrypt']['roofline']['frac'],4), r['gpu_launches'])"
