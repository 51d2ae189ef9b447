set -u
echo "== slow producers, gate on"; IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_KNOCKOUT=64 timeout 120 python tools/race_probe2.py 400 300 2 2>&1 | grep "bad rows"
echo "== slow producers, gate OFF"; IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_KNOCKOUT=192 timeout 120 python tools/race_probe2.py 400 300 2 2>&1 | grep "bad rows"
