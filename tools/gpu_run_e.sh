set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest.log
timeout 900 python tools/sweep.py > gpurun_out/r2e_sweep.txt 2> gpurun_out/r2e_sweep.err
timeout 600 python tools/microbench.py cfg4 > gpurun_out/r2e_cfg4.jsonl 2> gpurun_out/r2e_cfg4.err
PDLB200_DMMA=ws timeout 600 python tools/microbench.py cfg4 > gpurun_out/r2e_cfg4_cpasync.jsonl 2>&1
grep -v "^\.\|^$" gpurun_out/r2e_pytest.log | tail -40
cat gpurun_out/r2e_sweep.txt | grep -v "long \|longlong"
cat gpurun_out/r2e_cfg4.jsonl gpurun_out/r2e_cfg4_cpasync.jsonl
