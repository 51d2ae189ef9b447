set -u
timeout 90 python bench.py --neighbors 5 --no-cpu-baseline --no-parity --no-decrypt --sustain 0 --e2e-steps 1 2>&1 | tail -1 | cut -c1-300; echo "prod rc=$?"
IDASH_B200_USE_PROFILE_LIB=1 timeout 90 python bench.py --neighbors 5 --no-cpu-baseline --no-parity --no-decrypt --sustain 0 --e2e-steps 1 2>&1 | tail -1 | cut -c1-300; echo "prof rc=$?"
IDASH_B200_USE_PROFILE_LIB=1 IDASH_B200_TUNE=456 timeout 90 python bench.py --neighbors 5 --no-cpu-baseline --no-parity --no-decrypt --sustain 0 --e2e-steps 1 2>&1 | tail -1 | cut -c1-300; echo "prof tune rc=$?"
