set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_pytest.log
K="regex:reduce_rows_kernel|reduce_finish_kernel|ew_tile_kernel|ew_kernel|collapse_records|mm_dmma|mm_exact|inner_|cx_kernel"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2h_bench_under_ncu.log 2>&1
grep -v "^\.\|^$" gpurun_out/r2h_pytest.log | tail -30
tail -3 gpurun_out/r2h_launches.csv | cut -c1-300
