# gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_run_8.sh'   (charged 8x: keep it short)
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_parallel.py -q --timeout 500 -k "nccl" > gpurun_out/r2n8_pytest8.log 2>&1; echo "rc=$?" >> gpurun_out/r2n8_pytest8.log
tail -3 gpurun_out/r2n8_pytest8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2n8_bench_n8.json 2> gpurun_out/r2n8_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2n8_bench_n4.json 2> gpurun_out/r2n8_bench_n4.err
python - <<'PY'
import json
for f in ('gpurun_out/r2n8_bench_n8.json','gpurun_out/r2n8_bench_n4.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
        print(json.dumps(d['extra']['cfg5'])[:700]); print(json.dumps(d['extra']['cfg5_strong'])[:700])
    except Exception as e:
        print(f, 'ERR', e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
