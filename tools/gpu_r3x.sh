set -u
echo "== old library (bc5adac)"
IDASH_B200_LIB=gpurun_ab/lib_bc5adac.so timeout 300 compute-sanitizer --tool memcheck python tools/race_probe.py 400 300 5 2>&1 | grep -v "^=========" | tail -6
echo "== new library, profile build, legacy zero fill (tune 344)"
IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_TUNE=344 timeout 300 compute-sanitizer --tool memcheck python tools/race_probe.py 400 300 5 2>&1 | grep -v "^=========" | tail -6
echo "== new library, profile build, bulk zero fill (tune 328)"
IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_TUNE=328 timeout 300 compute-sanitizer --tool memcheck python tools/race_probe.py 400 300 5 2>&1 | grep -v "^=========" | tail -6
echo "== new library, no zero fill (ko 16)"
IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_KNOCKOUT=16 timeout 300 compute-sanitizer --tool memcheck python tools/race_probe.py 400 300 5 2>&1 | grep -v "^=========" | tail -6
