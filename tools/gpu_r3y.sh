timeout 300 compute-sanitizer --tool memcheck python tools/race_probe2.py 400 300 3 2>&1 | grep -v "^=========" | tail -40
