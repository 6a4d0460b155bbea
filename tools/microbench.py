#!/usr/bin/env python
"""Per-config kernel timings (CUDA events, inputs resident, L2-cold by rotation or size) for
the BASELINE.json configs other than the bench.py headline.  Prints one JSON line per config.
`next` times the SURVEY.md §8(f) rows (Bad.pd ops, constructors, inner, scans).
Usage: python tools/microbench.py [cfg1] [cfg3] [cfg4] [cfg5] [next] [--reps N]"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pdl_b200 as P  # noqa: E402
from pdl_b200 import types as T, ufunc  # noqa: E402

PEAK = 6538.3
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def wrap(eng, t, typ, dims):
    return P.PDL(eng, eng.wrap(t.data_ptr(), t.numel() * t.element_size(), t), typ, dims)


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    reps = 20
    if "--reps" in sys.argv:
        reps = int(sys.argv[sys.argv.index("--reps") + 1])
    want = set(args) or {"cfg1", "cfg3", "cfg4", "cfg5", "next"}
    eng = P.CudaEngine(0)
    P.set_default_engine(eng)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)

    if "cfg1" in want:
        # $x = $y + $c, 2048x2048 double; 8 rotating buffer sets (768 MB) so every launch is L2-cold
        n, sets = 2048 * 2048, 8
        ys = [(torch.randint(-2**20, 2**20, (n,), device=dev, generator=g).double() / 1024) for _ in range(sets)]
        cs = [(torch.randint(-2**20, 2**20, (n,), device=dev, generator=g).double() / 1024) for _ in range(sets)]
        xs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(sets)]
        py = [wrap(eng, t, T.D, [2048, 2048]) for t in ys]
        pc = [wrap(eng, t, T.D, [2048, 2048]) for t in cs]
        px = [wrap(eng, t, T.D, [2048, 2048]) for t in xs]
        k = [0]
        prep = [P.prepare_op("plus", [py[i], pc[i]], [px[i]]) for i in range(sets)]
        # launch-bound inner loop -> capture one rotation of the 8 buffer sets in a CUDA graph
        cap = torch.cuda.Stream()
        eng.stream = cap.cuda_stream
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(cap):
            for q in prep:
                q()
            cap.synchronize()
            with torch.cuda.graph(graph, stream=cap):
                for q in prep:
                    q()
        eng.stream = None
        ms = timeit(graph.replay, reps) / sets
        by = 3 * 8 * n
        print(json.dumps({"cfg": "cfg1 plus 2048x2048 double (CUDA graph of 8 L2-cold launches)", "ms": ms, "gbs": by / ms / 1e6,
                          "frac": by / ms / 1e6 / PEAK, "elements_per_sec": n / ms * 1e3}))

        def f():
            i = k[0] % sets
            k[0] += 1
            prep[i]()
        ms = timeit(f, reps * 8)
        ok = bool(torch.equal(xs[0], ys[0] + cs[0]))
        by = 3 * 8 * n
        print(json.dumps({"cfg": "cfg1 plus 2048x2048 double (prepared descriptor, stream launches)", "ms": ms, "gbs": by / ms / 1e6,
                          "frac": by / ms / 1e6 / PEAK, "elements_per_sec": n / ms * 1e3, "bitexact_vs_torch": ok}))

        def f2():
            i = k[0] % sets
            k[0] += 1
            return py[i] + pc[i]
        ms = timeit(f2, reps * 8)
        print(json.dumps({"cfg": "cfg1 $x=$y+$c fresh output (pool alloc, no memset)", "ms": ms, "gbs": by / ms / 1e6,
                          "frac": by / ms / 1e6 / PEAK}))
        del ys, cs, xs, py, pc, px

    if "cfg3" in want:
        N = M = 32768
        big1 = torch.randint(-1024, 1024, (2 * N,), device=dev, generator=g).double() / 256
        big2 = torch.randint(-1024, 1024, (2 * M,), device=dev, generator=g).double() / 256
        a = wrap(eng, big1, T.D, [2 * N]).slice("0:-1:2").dummy(1, 1)
        b = wrap(eng, big2, T.D, [2 * M]).slice("0:-1:2").dummy(0, 1)
        prod = torch.empty(N * M, dtype=torch.float64, device=dev)
        pp = wrap(eng, prod, T.D, [N, M])
        out = P.PDL.empty(T.D, [M], eng)
        ms_m = timeit(lambda: P.run_op("mult", [a, b], [pp]), reps)
        ms_s = timeit(lambda: P.run_op("sumover", [pp], [out]), reps)
        by_m, by_s = 8 * N * M + 2 * 8 * N, 8 * N * M + 8 * M
        ref = (big1[::2].sum() * big2[::2]).cpu().numpy()
        got = out.to_numpy()
        print(json.dumps({"cfg": "cfg3 [N,1]*[1,M] strided+dummy mult, N=M=32768 double", "ms": ms_m, "gbs": by_m / ms_m / 1e6,
                          "frac": by_m / ms_m / 1e6 / PEAK}))
        print(json.dumps({"cfg": "cfg3 sumover of the 8 GiB product", "ms": ms_s, "gbs": by_s / ms_s / 1e6,
                          "frac": by_s / ms_s / 1e6 / PEAK, "bitexact_vs_closed_form": bool(np.array_equal(got, ref))}))
        del prod, pp

    if "cfg4" in want:
        for n in (2048, 4096, 8192):
            A = (torch.randint(-64, 64, (n, n), device=dev, generator=g).double() / 64)
            B = (torch.randint(-64, 64, (n, n), device=dev, generator=g).double() / 64)
            C = torch.empty((n, n), dtype=torch.float64, device=dev)
            pa, pb, pc = wrap(eng, A, T.D, [n, n]), wrap(eng, B, T.D, [n, n]), wrap(eng, C, T.D, [n, n])
            r = max(2, reps // (4 if n >= 8192 else 1))
            ms = timeit(lambda: P.run_op("matmult", [pa, pb], [pc]), r, warm=1)
            ms_cublas = timeit(lambda: torch.matmul(A, B), r, warm=1)
            ok = bool(torch.equal(C, torch.matmul(A, B)))
            fl = 2.0 * n ** 3
            print(json.dumps({"cfg": f"cfg4 matmult {n}^3 double ({eng.last_kernel()})", "ms": ms, "tflops": fl / ms / 1e9,
                              "cublas_dgemm_ms": ms_cublas, "cublas_tflops": fl / ms_cublas / 1e9,
                              "frac_of_cublas": ms_cublas / ms, "bitexact_vs_cublas_exact_inputs": ok}))
            del A, B, C

    if "cfg5" in want:
        n = 2 ** 33 if torch.cuda.mem_get_info()[0] > 40 * 2**30 else 2 ** 30
        x = torch.empty(n, dtype=torch.float32, device=dev)
        step = 2 ** 28
        for i in range(0, n, step):
            x[i:i + step] = torch.randint(-1, 2, (step,), device=dev, generator=g).float()
        x[12345678] = 7.0
        px = wrap(eng, x, T.F, [n])
        o1, o2 = P.PDL.empty(T.F, [], eng), P.PDL.empty(T.F, [], eng)
        ms_s = timeit(lambda: P.run_op("sumover", [px], [o1]), max(2, reps // 4), warm=1)
        ms_m = timeit(lambda: P.run_op("maximum", [px], [o2]), max(2, reps // 4), warm=1)
        ref = float(x.double().sum().item()) if n <= 2**33 else None
        by = 4 * n
        print(json.dumps({"cfg": f"cfg5 sum of float[2^{int(np.log2(n))}] on 1 GPU", "ms": ms_s, "gbs": by / ms_s / 1e6,
                          "frac": by / ms_s / 1e6 / PEAK, "sum": o1.sclr(), "exact": o1.sclr() == ref}))
        print(json.dumps({"cfg": f"cfg5 max of float[2^{int(np.log2(n))}] on 1 GPU", "ms": ms_m, "gbs": by / ms_m / 1e6,
                          "frac": by / ms_m / 1e6 / PEAK, "max": o2.sclr()}))

    if "next" in want:
        from pdl_b200 import bad as B, basic
        n = 2 ** 28          # 1 GiB of float per operand: far beyond L2
        x = torch.randint(-8, 9, (n,), device=dev, generator=g).float()
        x[torch.rand(n, device=dev, generator=g) < 0.01] = -3.4028234663852886e38
        m = (torch.rand(n, device=dev, generator=g) < 0.1).int()
        px, pm = wrap(eng, x, T.F, [n]), wrap(eng, m, T.L, [n])
        px.badflag = True
        of, ol = P.PDL.empty(T.F, [n], eng), P.PDL.empty(T.L, [n], eng)

        def row(name, fn, by, r=reps, **extra):
            ms = timeit(fn, r)
            print(json.dumps({"cfg": name, "ms": ms, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / PEAK,
                              "kernel": eng.last_kernel(), **extra}))
        row("next setbadif float[2^28] + int mask", lambda: P.run_op("setbadif", [px, pm], [of]), 12 * n)
        row("next setbadtoval float[2^28], 1% BAD", lambda: P.run_op("setbadtoval", [px], [of], param=0.0), 8 * n)
        row("next isbad float[2^28] -> long", lambda: P.run_op("isbad", [px], [ol]), 8 * n)
        row("next copybad float[2^28]", lambda: P.run_op("copybad", [of, px], [of]), 12 * n)
        row("next sequence(float, 2^28) (axisvalues, write only)", lambda: P.run_op("axisvalues", [of], [of]), 4 * n)
        od = P.PDL.empty(T.D, [n // 2], eng)
        row("next sequence(double, 2^27)", lambda: P.run_op("axisvalues", [od], [od]), 8 * (n // 2))
        py_ = wrap(eng, torch.randint(-8, 9, (n,), device=dev, generator=g).float(), T.F, [n])
        px.badflag = False
        o0 = P.PDL.empty(T.F, [], eng)
        row("next inner of two float[2^28] (dot product)", lambda: P.run_op("inner", [px, py_], [o0]), 8 * n)
        row("next cumusumover float[2^28] (one row, single-pass look-back scan)", lambda: P.run_op("cumusumover", [py_], [of]), 8 * n)
        mm = [P.PDL.empty(T.F, [], eng), P.PDL.empty(T.F, [], eng), P.PDL.empty(T.IND, [], eng), P.PDL.empty(T.IND, [], eng)]
        # minmaximum ends in a flag read-back + stream sync (float rows may be all-NaN), so host-side descriptor
        # preparation is not hidden behind the kernel: time the prepared descriptor (one C-ABI call)
        row("next minmaximum float[2^28] (minmax of a flat ndarray; prepared)", P.prepare_op("minmaximum", [py_], mm), 4 * n)
        row("next magnover float[2^28]", lambda: P.run_op("magnover", [py_], [o0]), 4 * n)
        x2 = wrap(eng, x, T.F, [16384, n // 16384])
        mm2 = [P.PDL.empty(T.F, [n // 16384], eng), P.PDL.empty(T.F, [n // 16384], eng),
               P.PDL.empty(T.IND, [n // 16384], eng), P.PDL.empty(T.IND, [n // 16384], eng)]
        row("next minmaximum float[16384,16384] (per row; prepared)", P.prepare_op("minmaximum", [x2], mm2), 4 * n)
        o2 = wrap(eng, torch.empty(n, dtype=torch.float32, device=dev), T.F, [16384, n // 16384])
        row("next cumusumover float[16384,16384] (warp per row)", lambda: P.run_op("cumusumover", [x2], [o2]), 8 * n)
        del x, m, px, pm, of, ol, od, py_, x2, o2
        # cfg3 fused: inner([N,1] strided, [1,M] strided) == sumover(mult(...)) without the 8 GiB intermediate
        N = M = 32768
        big1 = torch.randint(-1024, 1024, (2 * N,), device=dev, generator=g).double() / 256
        big2 = torch.randint(-1024, 1024, (2 * M,), device=dev, generator=g).double() / 256
        a = wrap(eng, big1, T.D, [2 * N]).slice("0:-1:2").dummy(1, 1)
        b = wrap(eng, big2, T.D, [2 * M]).slice("0:-1:2").dummy(0, 1)
        out = P.PDL.empty(T.D, [M], eng)
        ms = timeit(lambda: P.run_op("inner", [a, b], [out]), reps)
        ref = (big1[::2].sum() * big2[::2]).cpu().numpy()
        print(json.dumps({"cfg": "next cfg3 FUSED: inner([N,1],[1,M]) N=M=32768 double (reported apart from the roofline figure)",
                          "ms": ms, "madds_per_sec": N * M / ms * 1e3, "traffic_bytes": 8 * (N + 2 * M),
                          "bitexact_vs_closed_form": bool(np.array_equal(out.to_numpy(), ref))}))


if __name__ == "__main__":
    main()
