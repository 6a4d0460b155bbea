cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py -m gpu -q --timeout 600 2>&1 | tail -12
timeout 900 python tools/sweep.py 2>/dev/null | grep "byte\|short\|op " | grep "sumover\|average\|op "
