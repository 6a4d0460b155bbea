set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "minmax or bad" > gpurun_out/r2l_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_pytest.log
tail -5 gpurun_out/r2l_pytest.log
timeout 600 python tools/microbench.py next 2>&1 | grep "minmaximum" | cut -c1-250
timeout 900 bash tools/ncu_summary.sh gpurun_out/ncu_r2l scan1d scan1d_bad > /dev/null 2>&1
