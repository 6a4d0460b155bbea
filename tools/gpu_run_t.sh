cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'roofline',d['roofline']['frac']); print({k:(round(v['ms'],4),round(v['frac'],3)) for k,v in d['per_op'].items()})
for k in ('cfg5','cfg5_strong'): print(k, d['extra'][k]['ms_per_step'], d['extra'][k]['frac_per_gpu'], d['extra'][k]['verified'])"
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
for op in ("sumover", "average", "minimum", "maximum_ind"):
    for n in (16384, 65536, 262144, 4194304, 2**28):
        rows = n_all // n
        px = wrap(eng, x[:rows*n], T.F, [n, rows]); px.badflag = True
        out = P.PDL.empty(T.IND if op.endswith("_ind") else T.F, [rows], eng)
        f = P.prepare_op(op, [px], [out])
        print(op, 4*n, rows, round(4 * n * rows / timeit(f, 10) / 1e6 / PEAK, 3))
PY
