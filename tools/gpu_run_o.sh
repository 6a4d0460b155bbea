set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py -m gpu -q --timeout 600 2>&1 | tail -3
timeout 900 python tools/sweep.py > gpurun_out/r2q_sweep.txt 2> gpurun_out/r2q_sweep.err
grep "byte\|short\|op " gpurun_out/r2q_sweep.txt | grep "minimum\|maximum\|op "; grep "divide.*double" gpurun_out/r2q_sweep.txt
