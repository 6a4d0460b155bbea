set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_parallel.py -q --timeout 600 -k "nccl" > gpurun_out/r2i_pytest8.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest8.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2i_bench_n$n.json 2> gpurun_out/r2i_bench_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 8 --steps 10 --warmup 3 --nccl-gather > gpurun_out/r2i_bench_n8_nccl.json 2> gpurun_out/r2i_bench_n8_nccl.err
tail -5 gpurun_out/r2i_pytest8.log
python - <<'PY'
import json
for f in ('gpurun_out/r2i_bench_n8.json','gpurun_out/r2i_bench_n4.json','gpurun_out/r2i_bench_n8_nccl.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('numa_node'))
        print(json.dumps(d['extra']['cfg5'])); print(json.dumps(d['extra']['cfg5_strong']))
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
