set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_pytest.log
grep -v "^\.\|^$" gpurun_out/r2m_pytest.log | tail -15
timeout 600 python tools/scan_bench.py both > gpurun_out/r2m_scan.txt 2> gpurun_out/r2m_scan.err
cat gpurun_out/r2m_scan.txt
timeout 600 python tools/microbench.py next 2>&1 | grep "minmaximum\|cumusum" | cut -c1-250
