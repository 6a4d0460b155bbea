cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py -m gpu -q --timeout 600 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'roofline',d['roofline']['frac'],d['roofline']['kernel']); print({k:(round(v['ms'],4),round(v['frac'],3)) for k,v in d['per_op'].items()})"
python - <<'PY'
import json, os, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tools')
import torch
import pdl_b200 as P
from pdl_b200 import types as T
from microbench import wrap, timeit, PEAK
eng = P.CudaEngine(0)
n_all = 2 ** 28
x = torch.randint(-8, 9, (n_all,), device="cuda").float()
x[torch.rand(n_all, device="cuda") < 0.01] = -3.4028234663852886e38
for op in ("minimum", "maximum_ind"):
    for n in (8192, 12288, 16384, 32768, 65536, 131072):
        rows = n_all // n
        px = wrap(eng, x[:rows*n], T.F, [n, rows]); px.badflag = True
        out = P.PDL.empty(T.IND if op.endswith("_ind") else T.F, [rows], eng)
        f = P.prepare_op(op, [px], [out])
        print(op, 4*n, rows, round(4 * n * rows / timeit(f, 10) / 1e6 / PEAK, 3))
PY
