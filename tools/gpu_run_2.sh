set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_parallel.py -q --timeout 600 > gpurun_out/r2r_pytest2.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_pytest2.log
tail -4 gpurun_out/r2r_pytest2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_bench_ref_n2.json 2> gpurun_out/r2r_bench_ref_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r2r_bench_n2.json','gpurun_out/r2r_bench_ref_n2.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', d['value'], 'ms', d.get('ms_per_step'), 'e2e', d['e2e']['value'])
        if 'extra' in d: print(json.dumps(d['extra']['cfg5'])[:600]); print(json.dumps(d['extra']['cfg5_strong'])[:600])
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
