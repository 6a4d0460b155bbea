set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_pytest.log
timeout 900 python tools/sweep.py > gpurun_out/r2j_sweep.txt 2> gpurun_out/r2j_sweep.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err
grep -v "^\.\|^$" gpurun_out/r2j_pytest.log | tail -20
grep "byte\|short\|op " gpurun_out/r2j_sweep.txt | grep "minimum\|_ind\|op "
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'])
for k in ('cfg5','cfg5_strong','cfg1'): print(k, json.dumps(d['extra'][k])[:700])
PY
