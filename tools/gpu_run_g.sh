set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err
timeout 900 python tools/sweep.py > gpurun_out/r2g_sweep.txt 2> gpurun_out/r2g_sweep.err
timeout 900 python tools/microbench.py next > gpurun_out/r2g_micro_next.jsonl 2> gpurun_out/r2g_micro.err
timeout 1500 bash tools/ncu_summary.sh gpurun_out/r2g_ncu matmult ew:divide:float:good ew:sqrt:float:good rd:maximum_ind:sbyte:good ew:plus:sbyte:bad minimum sumover average plus minmaximum matmult_exact_float > gpurun_out/r2g_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2g_bench_under_ncu.log 2>&1
grep -v "^\.\|^$" gpurun_out/r2g_pytest.log | tail -20
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'roof',d['roofline']['frac'], 'cpu', d['cpu_baseline'])
for k,v in d['extra'].items(): print(k, json.dumps(v)[:1800])
print(open('gpurun_out/r2g_bench_ref.json').read()[-900:])
PY
grep "divide\|sqrt\|_ind\|op " gpurun_out/r2g_sweep.txt | grep "float\|double\|byte\|short\|op "
cat gpurun_out/r2g_micro_next.jsonl | cut -c1-260
