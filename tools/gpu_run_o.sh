set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py -m gpu -q --timeout 600 -k "matmult or golden or fuzz" 2>&1 | tail -12
python - <<'PY'
import sys, json
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import torch, config_legs
import pdl_b200 as P
eng = P.CudaEngine(0)
r = config_legs.cfg4(eng, torch.device('cuda', 0))
print(json.dumps({k: r[k] for k in ('exact_float_4096', 'exact_bad_double_4096')}))
PY
