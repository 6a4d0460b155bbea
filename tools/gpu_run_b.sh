set -x
cd $GRAFT_REPO_ROOT
R=oracle/_ref/blib; S=perl/PDL-B200/blib
INC="-I$R/lib -I$R/arch -I$S/lib -I$S/arch -Ioracle/shim"
timeout 300 perl $INC perl/PDL-B200/t/store.t > gpurun_out/r2b_store.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_store.log
timeout 600 perl $INC perl/PDL-B200/t/parity.t > gpurun_out/r2b_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_parity.log
timeout 900 python -m pytest tests/test_gpu_perl_shim.py -q -x --timeout 600 > gpurun_out/r2b_shim.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_shim.log
timeout 300 perl $INC perl/PDL-B200/bench_ops.pl --reps 50 > gpurun_out/r2b_perl_bench.json 2> gpurun_out/r2b_perl_bench.err
grep -c "^ok" gpurun_out/r2b_store.log; grep "^not ok" gpurun_out/r2b_store.log | head -20; tail -5 gpurun_out/r2b_store.log
grep -c "^ok" gpurun_out/r2b_parity.log; grep "^not ok" gpurun_out/r2b_parity.log | head -20; tail -3 gpurun_out/r2b_parity.log
tail -15 gpurun_out/r2b_shim.log
cat gpurun_out/r2b_perl_bench.json; tail -3 gpurun_out/r2b_perl_bench.err
