#!/usr/bin/env python
"""Launch ONE op of the bench workload a few times (for `ncu --set full -k regex:...`).
Usage: python tools/prof_one.py {sumover|average|minimum|plus|mult_cfg3|matmult|setbadif|inner|minmaximum|cumusumover|sequence} [reps]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import pdl_b200 as P  # noqa: E402
from pdl_b200 import types as T  # noqa: E402

op = sys.argv[1] if len(sys.argv) > 1 else "minimum"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
eng = P.CudaEngine(0)
dev = torch.device("cuda", 0)


def wrap(t, typ, dims):
    return P.PDL(eng, eng.wrap(t.data_ptr(), t.numel() * t.element_size(), t), typ, dims)


if op in ("sumover", "average", "minimum", "maximum"):
    rows = 65536
    x = bench.generate_device(torch, 0, rows, dev)
    a = wrap(x, T.F, [bench.N_DIM, rows]).set_badflag(True)
    out = P.PDL.empty(T.F, [rows], eng)
    f = P.prepare_op(op, [a], [out])
elif op == "plus":
    n = 2048 * 2048
    y, c, o = (torch.rand(n, dtype=torch.float64, device=dev) for _ in range(3))
    f = P.prepare_op("plus", [wrap(y, T.D, [2048, 2048]), wrap(c, T.D, [2048, 2048])], [wrap(o, T.D, [2048, 2048])])
elif op == "mult_cfg3":
    N = 32768
    b1, b2 = torch.rand(2 * N, dtype=torch.float64, device=dev), torch.rand(2 * N, dtype=torch.float64, device=dev)
    pr = torch.empty(N * N, dtype=torch.float64, device=dev)
    f = P.prepare_op("mult", [wrap(b1, T.D, [2 * N]).slice("0:-1:2").dummy(1, 1),
                              wrap(b2, T.D, [2 * N]).slice("0:-1:2").dummy(0, 1)], [wrap(pr, T.D, [N, N])])
elif op == "matmult_exact_float":
    n = 4096
    A, B, Cc = (torch.randint(-8, 8, (n, n), device=dev).float() for _ in range(3))
    f = P.prepare_op("matmult", [wrap(A, T.F, [n, n]), wrap(B, T.F, [n, n])], [wrap(Cc, T.F, [n, n])])
elif op == "matmult_2048":          # 256 tiles on 148 SMs: the stream-K persistent kernel
    n = 2048
    A, B, Cc = (torch.rand((n, n), dtype=torch.float64, device=dev) for _ in range(3))
    f = P.prepare_op("matmult", [wrap(A, T.D, [n, n]), wrap(B, T.D, [n, n])], [wrap(Cc, T.D, [n, n])])
elif op == "matmult_exact_bad_double":
    n = 4096
    A, B, Cc = (torch.randint(-64, 64, (n, n), device=dev).double() / 64 for _ in range(3))
    A[17, 33] = -1.7976931348623157e308
    f = P.prepare_op("matmult", [wrap(A, T.D, [n, n]).set_badflag(True), wrap(B, T.D, [n, n])], [wrap(Cc, T.D, [n, n])])
elif op == "matmult":
    n = 4096
    A, B, Cc = (torch.rand((n, n), dtype=torch.float64, device=dev) for _ in range(3))
    f = P.prepare_op("matmult", [wrap(A, T.D, [n, n]), wrap(B, T.D, [n, n])], [wrap(Cc, T.D, [n, n])])
elif op in ("setbadif", "inner", "minmaximum", "cumusumover", "sequence"):
    n = 2 ** 28
    x = torch.randint(-8, 9, (n,), device=dev).float()
    px = wrap(x, T.F, [n])
    of = P.PDL.empty(T.F, [n], eng)
    if op == "setbadif":
        m = (torch.rand(n, device=dev) < 0.1).int()
        f = P.prepare_op("setbadif", [px, wrap(m, T.L, [n])], [of])
    elif op == "inner":
        y = torch.randint(-8, 9, (n,), device=dev).float()
        f = P.prepare_op("inner", [px, wrap(y, T.F, [n])], [P.PDL.empty(T.F, [], eng)])
    elif op == "minmaximum":
        x2 = wrap(x, T.F, [16384, n // 16384])
        f = P.prepare_op("minmaximum", [x2], [P.PDL.empty(T.F, [n // 16384], eng), P.PDL.empty(T.F, [n // 16384], eng),
                                              P.PDL.empty(T.IND, [n // 16384], eng), P.PDL.empty(T.IND, [n // 16384], eng)])
    elif op == "cumusumover":
        f = P.prepare_op("cumusumover", [wrap(x, T.F, [16384, n // 16384])], [wrap(torch.empty_like(x), T.F, [16384, n // 16384])])
    else:
        f = P.prepare_op("axisvalues", [of], [of])
elif op in ("scan1d", "scan1d_bad", "minmax1d"):
    n = 2 ** 28
    x = torch.randint(-8, 9, (n,), device=dev).float()
    px = wrap(x, T.F, [n])
    if op == "minmax1d":
        f = P.prepare_op("minmaximum", [px], [P.PDL.empty(T.F, [], eng), P.PDL.empty(T.F, [], eng),
                                              P.PDL.empty(T.IND, [], eng), P.PDL.empty(T.IND, [], eng)])
    else:
        px.badflag = op.endswith("bad")
        f = P.prepare_op("cumusumover", [px], [P.PDL.empty(T.F, [n], eng)])
elif op.startswith(("ew:", "rd:")):
    # generic: ew:<op>:<type>:<good|bad> / rd:<op>:<type>:<good|bad> on 1 GiB operands (as tools/sweep.py)
    kind, name, tname, mode = op.split(":")
    t = T.NAMES.index(tname)
    tt = {T.SB: torch.int8, T.B: torch.uint8, T.S: torch.int16, T.L: torch.int32, T.LL: torch.int64, T.F: torch.float32, T.D: torch.float64}[t]
    n = (1 << 30) // T.SIZE[t]
    a = torch.randint(0 if tt == torch.uint8 else -8, 9, (n,), device=dev, dtype=torch.int32).to(tt)
    b = torch.randint(1, 9, (n,), device=dev, dtype=torch.int32).to(tt)
    c = torch.empty_like(a)
    bad = mode == "bad"
    if kind == "ew":
        pa, pb, pc = wrap(a, t, [n]), wrap(b, t, [n]), wrap(c, t, [n])
        pa.badflag = pb.badflag = bad
        f = P.prepare_op(name, [pa, pb] if len(P.SPECS[name].pars) == 3 else [pa], [pc])
    else:
        p2 = wrap(a, t, [16384, n // 16384])
        p2.badflag = bad
        ot = P.trans.par_type(P.SPECS[name].pars[1], t)
        f = P.prepare_op(name, [p2], [P.PDL.empty(ot, [n // 16384], eng)])
else:
    raise SystemExit(f"unknown op {op}")
for _ in range(reps):
    f()
torch.cuda.synchronize()
print("done", op, eng.last_kernel())
