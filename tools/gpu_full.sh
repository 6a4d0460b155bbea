#!/usr/bin/env bash
# One gpurun call = the round's evidence on ONE B200:
#   gpurun --timeout 3000 -- 'bash tools/gpu_full.sh <tag>'
# full `pytest -m gpu`, the N=1 bench line with every extra leg, the op x type sweep, the scan bench and the `next`
# microbench rows; outputs under gpurun_out/<tag>_*.  (tools/gpu_run_2.sh / gpu_run_8.sh: the multi-GPU runs.)
set -x
cd "$GRAFT_REPO_ROOT"
TAG="${1:-full}"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -v "^\.\|^$" gpurun_out/${TAG}_pytest.log | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -2 gpurun_out/${TAG}_bench_n1.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 900 python tools/sweep.py > gpurun_out/${TAG}_sweep.txt 2> gpurun_out/${TAG}_sweep.err
timeout 600 python tools/scan_bench.py both > gpurun_out/${TAG}_scan.txt 2> gpurun_out/${TAG}_scan.err
timeout 900 python tools/microbench.py next > gpurun_out/${TAG}_microbench_next.jsonl 2> gpurun_out/${TAG}_microbench_next.err
timeout 900 python tools/microbench.py cfg4 > gpurun_out/${TAG}_microbench_cfg4.jsonl 2> gpurun_out/${TAG}_microbench_cfg4.err
tail -1 gpurun_out/${TAG}_bench_n1.json | cut -c1-600
