set -x
cd $GRAFT_REPO_ROOT
R=oracle/_ref/blib; S=perl/PDL-B200/blib
INC="-I$R/lib -I$R/arch -I$S/lib -I$S/arch -Ioracle/shim"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest.log
timeout 300 perl $INC perl/PDL-B200/bench_ops.pl --reps 50 > gpurun_out/r2c_perl_bench.json 2> gpurun_out/r2c_perl_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
grep -v "^\.\|^$" gpurun_out/r2c_pytest.log | tail -60
cat gpurun_out/r2c_perl_bench.json; tail -3 gpurun_out/r2c_perl_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_n1.json').read().strip().splitlines()[-1])
print(json.dumps(d['extra']['cfg1'])[:900])
print(json.dumps(d['extra'].get('perl'))[:1500])
PY
tail -5 gpurun_out/r2c_bench_n1.err
