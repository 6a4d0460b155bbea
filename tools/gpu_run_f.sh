set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest.log
timeout 900 python tools/sweep.py > gpurun_out/r2f_sweep.txt 2> gpurun_out/r2f_sweep.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --nccl-gather > gpurun_out/r2f_bench_n2_nccl.json 2> gpurun_out/r2f_bench_n2_nccl.err
grep -v "^\.\|^$" gpurun_out/r2f_pytest.log | tail -30
grep "float\|double\|_ind\|op " gpurun_out/r2f_sweep.txt
python - <<'PY'
import json
for f in ('gpurun_out/r2f_bench_n2.json','gpurun_out/r2f_bench_n2_nccl.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['e2e'])
        print(json.dumps(d['extra']['cfg5'])); print(json.dumps(d['extra']['cfg5_strong']))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
