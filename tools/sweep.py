#!/usr/bin/env python
"""Bandwidth sweep: op x type x {good, BAD} on 1 GiB operands (CUDA events, prepared descriptors).
Prints a table of GB/s (algorithmic bytes) and fraction of the measured HBM peak."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pdl_b200 as P  # noqa: E402
from pdl_b200 import types as T  # noqa: E402

PEAK = 6538.3
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
eng = P.CudaEngine(0)
dev = torch.device("cuda", 0)
NBYTES = 1 << 30
TT = {T.SB: torch.int8, T.B: torch.uint8, T.S: torch.int16, T.L: torch.int32, T.LL: torch.int64, T.F: torch.float32, T.D: torch.float64}


def wrap(t, typ, dims):
    return P.PDL(eng, eng.wrap(t.data_ptr(), t.numel() * t.element_size(), t), typ, dims)


def timeit(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows = []
for t, tt in TT.items():
    n = NBYTES // T.SIZE[t]
    if tt.is_floating_point:
        a = (torch.randint(-8, 9, (n,), device=dev)).to(tt)
        b = (torch.randint(1, 9, (n,), device=dev)).to(tt)
    else:
        a = torch.randint(0 if tt == torch.uint8 else -8, 9, (n,), device=dev, dtype=torch.int32).to(tt)
        b = torch.randint(1, 9, (n,), device=dev, dtype=torch.int32).to(tt)
    c = torch.empty_like(a)
    ncols = 16384
    for bad in (False, True):
        pa, pb, pc = wrap(a, t, [n]), wrap(b, t, [n]), wrap(c, t, [n])
        pa.badflag = pb.badflag = bad
        for op in ("plus", "mult", "divide", "lt"):
            f = P.prepare_op(op, [pa, pb], [pc])
            ms = timeit(f)
            rows.append((op, T.NAMES[t], bad, 3 * NBYTES / ms / 1e6))
        for op in ("sqrt", "_rabs"):
            f = P.prepare_op(op, [pa], [pc])
            ms = timeit(f)
            rows.append((op, T.NAMES[t], bad, 2 * NBYTES / ms / 1e6))
        p2 = wrap(a, t, [ncols, n // ncols])
        p2.badflag = bad
        for op in ("sumover", "average", "minimum", "maximum_ind", "prodover", "orover"):
            spec = P.SPECS[op]
            ot = P.trans.par_type(spec.pars[1], t)
            out = P.PDL.empty(ot, [n // ncols], eng)
            f = P.prepare_op(op, [p2], [out])
            ms = timeit(f)
            rows.append((op, T.NAMES[t], bad, NBYTES / ms / 1e6))
    del a, b, c
print(f"{'op':12s} {'type':9s} {'good GB/s':>10s} {'frac':>6s} {'BAD GB/s':>10s} {'frac':>6s}")
seen = {}
for op, tn, bad, gbs in rows:
    seen.setdefault((op, tn), {})[bad] = gbs
for (op, tn), d in seen.items():
    print(f"{op:12s} {tn:9s} {d[False]:10.0f} {d[False] / PEAK:6.2f} {d[True]:10.0f} {d[True] / PEAK:6.2f}")
