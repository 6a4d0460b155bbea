set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "scan" 2>&1 | tail -5
