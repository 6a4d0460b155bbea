set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py -m gpu -q --timeout 600 -k "minmax or golden or fuzz or bad" 2>&1 | tail -3
timeout 600 python tools/microbench.py next 2>&1 | grep "minmaximum" | cut -c1-250
