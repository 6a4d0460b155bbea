set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "matmult" 2>&1 | tail -2
PDLB200_MM_STREAMK=1 timeout 600 python tools/microbench.py cfg4 > gpurun_out/r2o_cfg4_sk1.jsonl 2>&1; cat gpurun_out/r2o_cfg4_sk1.jsonl | cut -c1-330
