set -u
timeout 300 compute-sanitizer --tool memcheck python tools/race_probe.py 400 300 6 2>&1 | grep -v "^=========" | tail -8
timeout 300 compute-sanitizer --tool memcheck python tools/race_probe.py 335 300 4 2>&1 | grep -v "^=========" | tail -6
timeout 300 compute-sanitizer --tool memcheck python tools/race_probe.py 1004 300 4 2>&1 | grep -v "^=========" | tail -6
timeout 100 python tools/race_probe.py 400 300 40 2>&1 | sort | uniq -c | tail -5
