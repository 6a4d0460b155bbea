set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x --timeout 300 -k "scan" > gpurun_out/r2k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_pytest.log
tail -3 gpurun_out/r2k_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_golden.py -m gpu -q --timeout 300 > gpurun_out/r2k_pytest2.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_pytest2.log
tail -5 gpurun_out/r2k_pytest2.log
timeout 600 python tools/scan_bench.py > gpurun_out/r2k_scan.txt 2> gpurun_out/r2k_scan.err
cat gpurun_out/r2k_scan.txt; tail -5 gpurun_out/r2k_scan.err
timeout 600 python tools/microbench.py next 2>&1 | grep "minmaximum\|cumusum" | cut -c1-250
