set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_pytest.log
grep -v "^\.\|^$" gpurun_out/r2p_pytest.log | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; tail -3 gpurun_out/r2p_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'roofline',d['roofline'],'cpu',d.get('cpu_baseline'))
for k in ('cfg1','cfg3','cfg4','next_rows','perl'): print(k, json.dumps(d['extra'].get(k))[:900])
PY
timeout 600 python tools/microbench.py next 2>&1 | grep "minmaximum" | cut -c1-250
