set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest.log
timeout 900 python tools/sweep.py > gpurun_out/r2d_sweep.txt 2> gpurun_out/r2d_sweep.err
timeout 900 bash tools/ncu_summary.sh gpurun_out/r2d_ncu ew:divide:float:good ew:divide:float:bad ew:sqrt:float:good ew:divide:double:good minmaximum rd:maximum_ind:sbyte:bad ew:lt:sbyte:bad > gpurun_out/r2d_ncu.log 2>&1
grep -v "^\.\|^$" gpurun_out/r2d_pytest.log | tail -40
cat gpurun_out/r2d_sweep.txt
