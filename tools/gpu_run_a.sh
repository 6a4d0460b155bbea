set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2a_bench_n2.json 2> gpurun_out/r2a_bench_n2.err
tail -3 gpurun_out/r2a_pytest.log; tail -2 gpurun_out/r2a_smoke.log; tail -c 600 gpurun_out/r2a_bench_n1.err; tail -c 600 gpurun_out/r2a_bench_n2.err
