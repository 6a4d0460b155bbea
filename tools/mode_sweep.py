"""GPU: warp-per-row (mode 1) against CTA-per-row (mode 2) for row reductions, by row size (1 GiB of float rows)."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import torch
import pdl_b200 as P
from pdl_b200 import types as T
from microbench import wrap, timeit, PEAK
eng = P.CudaEngine(0)
n_all = 2 ** 28
x = torch.randint(-8, 9, (n_all,), device="cuda").float()
x[torch.rand(n_all, device="cuda") < 0.01] = -3.4028234663852886e38
for op in sys.argv[1:] or ["minimum", "sumover", "average", "maximum_ind"]:
    for n in (4096, 8192, 16384, 32768, 65536, 131072, 262144, 1048576):
        rows = n_all // n
        for bad in (True, False):
            px = wrap(eng, x, T.F, [n, rows]); px.badflag = bad
            out = P.PDL.empty(T.IND if op.endswith("_ind") else T.F, [rows], eng)
            res = {}
            for mode in ("1", "2"):
                os.environ["PDLB200_REDUCE_MODE"] = mode
                f = P.prepare_op(op, [px], [out])
                res[mode] = round(4 * n_all / timeit(f, 10) / 1e6 / PEAK, 3)
            os.environ.pop("PDLB200_REDUCE_MODE")
            print(json.dumps({"op": op, "row_bytes": 4 * n, "bad": bad, "warp_per_row": res["1"], "cta_per_row": res["2"]}), flush=True)
