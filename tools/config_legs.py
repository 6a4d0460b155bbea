"""The BASELINE.json configurations other than the headline one (cfg2), as functions bench.py calls to put
them under the driver's clock (`extra.cfg1 / cfg3 / cfg4`).  Every leg: inputs resident in HBM, CUDA events on
the launch stream, L2-cold by rotation or by size, result verified against a closed form / an independent
computation outside the timed region.  Algorithmic bytes and flops are SURVEY.md §8(d)'s."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pdl_b200 as P  # noqa: E402
from pdl_b200 import types as T  # noqa: E402


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


def cfg1(eng, dev, peak, reps=25):
    """$x = $y + $c on two 2048x2048 double ndarrays.  8 rotating buffer sets (768 MB > L2) so every launch is
    L2-cold.  Three ways to issue the same op: a CUDA graph of prepared descriptors (kernel floor), prepared
    descriptors from the stream, and the operator surface with a FRESH output per op (`py + pc`: descriptor
    cache + size-class free list in front of the pool)."""
    g = torch.Generator(device=dev).manual_seed(1)
    n, sets = 2048 * 2048, 8
    ys = [(torch.randint(-2**20, 2**20, (n,), device=dev, generator=g).double() / 1024) for _ in range(sets)]
    cs = [(torch.randint(-2**20, 2**20, (n,), device=dev, generator=g).double() / 1024) for _ in range(sets)]
    xs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(sets)]
    py = [wrap(eng, t, T.D, [2048, 2048]) for t in ys]
    pc = [wrap(eng, t, T.D, [2048, 2048]) for t in cs]
    px = [wrap(eng, t, T.D, [2048, 2048]) for t in xs]
    prep = [P.prepare_op("plus", [py[i], pc[i]], [px[i]]) for i in range(sets)]
    by = 3 * 8 * n
    out = {"workload": "cfg1: $x = $y + $c, 2048x2048 double, L2-cold (8 rotating buffer sets)",
           "algorithmic_bytes": by}
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
    out["graph"] = {"ms": ms, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
    k = [0]

    def f():
        prep[k[0] % sets]()
        k[0] += 1
    ms = timeit(f, reps * sets)
    out["prepared"] = {"ms": ms, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}

    def f2():
        i = k[0] % sets
        k[0] += 1
        return py[i] + pc[i]
    ms = timeit(f2, reps * sets, warm=2 * sets)
    out["fresh_output_operator"] = {"ms": ms, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
    r = py[0] + pc[0]
    out["verified"] = bool(torch.equal(xs[0], ys[0] + cs[0])) and \
        bool(np.array_equal(r.to_numpy().reshape(-1), (ys[0] + cs[0]).cpu().numpy()))
    out["ms"], out["frac"] = out["prepared"]["ms"], out["prepared"]["frac"]
    return out


def cfg3(eng, dev, peak, reps=10):
    """[N,1] * [1,M] dummy-dim multiply of strided slices, then sumover; N = M = 32768 double.  Unfused (the
    intermediate is a real 8 GiB ndarray, the roofline figure) and fused (one launch, reported apart)."""
    g = torch.Generator(device=dev).manual_seed(3)
    N = M = 32768
    big1 = torch.randint(-1024, 1024, (2 * N,), device=dev, generator=g).double() / 256
    big2 = torch.randint(-1024, 1024, (2 * M,), device=dev, generator=g).double() / 256
    a = wrap(eng, big1, T.D, [2 * N]).slice("0:-1:2").dummy(1, 1)
    b = wrap(eng, big2, T.D, [2 * M]).slice("0:-1:2").dummy(0, 1)
    prod = torch.empty(N * M, dtype=torch.float64, device=dev)
    pp = wrap(eng, prod, T.D, [N, M])
    o = P.PDL.empty(T.D, [M], eng)
    ms_m = timeit(lambda: P.run_op("mult", [a, b], [pp]), reps)
    ms_s = timeit(lambda: P.run_op("sumover", [pp], [o]), reps)
    by_m, by_s = 8 * N * M + 2 * 8 * N, 8 * N * M + 8 * M
    ref = (big1[::2].sum() * big2[::2]).cpu().numpy()
    unfused_bits = o.to_numpy().copy()
    res = {"workload": "cfg3: ([N,1] * [1,M])->sumover, N=M=32768 double, strided slice + dummy-dim inputs",
           "mult": {"ms": ms_m, "gbs": by_m / ms_m / 1e6, "frac": by_m / ms_m / 1e6 / peak},
           "sumover": {"ms": ms_s, "gbs": by_s / ms_s / 1e6, "frac": by_s / ms_s / 1e6 / peak},
           "ms": ms_m + ms_s, "algorithmic_bytes": by_m + by_s,
           "frac": (by_m + by_s) / (ms_m + ms_s) / 1e6 / peak,
           "verified": bool(np.array_equal(unfused_bits, ref))}
    del prod, pp
    # the operator surface: ($a * $b)->sumover with the product consumed only by the reduction -> fused
    from pdl_b200 import ufunc
    fused = lambda: ufunc.sumover(a.flowing() * b)        # noqa: E731
    got = fused()
    ms_f = timeit(fused, reps)
    res["fused_operator"] = {"ms": ms_f, "launches": 1, "traffic_bytes": 8 * (N + 2 * M),
                             "bit_identical_to_unfused": bool(np.array_equal(got.to_numpy(), unfused_bits)),
                             "note": "reported apart from the roofline figure (SURVEY.md 8(d))"}
    return res


def cfg4(eng, dev, reps=4):
    """matmult 8192^3 double on the FP64 tensor cores, cuBLAS DGEMM beside it; the exact-order kernel (float, and
    double with BAD values) at 4096^3."""
    g = torch.Generator(device=dev).manual_seed(4)
    res = {"workload": "cfg4: matmult 8192x8192 double (FP64 tensor-core path) + exact-order kernel at 4096^3"}
    for n in (2048, 8192):
        A = (torch.randint(-64, 64, (n, n), device=dev, generator=g).double() / 64)
        B = (torch.randint(-64, 64, (n, n), device=dev, generator=g).double() / 64)
        C = torch.empty((n, n), dtype=torch.float64, device=dev)
        pa, pb, pc = wrap(eng, A, T.D, [n, n]), wrap(eng, B, T.D, [n, n]), wrap(eng, C, T.D, [n, n])
        prep = P.prepare_op("matmult", [pa, pb], [pc])
        ms = timeit(prep, reps if n >= 8192 else 4 * reps, warm=1)
        kern = eng.last_kernel()
        ms_cublas = timeit(lambda: torch.matmul(A, B), reps if n >= 8192 else 4 * reps, warm=1)
        fl = 2.0 * n ** 3
        res[f"dmma_{n}"] = {"ms": ms, "tflops": fl / ms / 1e9, "kernel": kern, "cublas_dgemm_ms": ms_cublas,
                            "cublas_tflops": fl / ms_cublas / 1e9, "vs_cublas": ms_cublas / ms,
                            "verified": bool(torch.equal(C, torch.matmul(A, B)))}     # exact inputs: order-independent
        del A, B, C
    n = 4096
    Af = torch.randint(-8, 8, (n, n), device=dev, generator=g).float()
    Bf = torch.randint(-8, 8, (n, n), device=dev, generator=g).float()
    Cf = torch.empty((n, n), dtype=torch.float32, device=dev)
    pa, pb, pc = wrap(eng, Af, T.F, [n, n]), wrap(eng, Bf, T.F, [n, n]), wrap(eng, Cf, T.F, [n, n])
    ms = timeit(P.prepare_op("matmult", [pa, pb], [pc]), 2, warm=1)
    res["exact_float_4096"] = {"ms": ms, "tflops": 2.0 * n ** 3 / ms / 1e9, "kernel": eng.last_kernel(),
                               "verified": bool(torch.equal(Cf, torch.matmul(Af.double(), Bf.double()).float()))}
    Ad = (torch.randint(-64, 64, (n, n), device=dev, generator=g).double() / 64)
    Bd = (torch.randint(-64, 64, (n, n), device=dev, generator=g).double() / 64)
    bad = -1.7976931348623157e308
    Ad[17, 33] = bad                      # one BAD element in a: row h=17 of c is BAD (Primitive.pd:227-241)
    Cd = torch.empty((n, n), dtype=torch.float64, device=dev)
    pa, pb, pc = wrap(eng, Ad, T.D, [n, n]).set_badflag(True), wrap(eng, Bd, T.D, [n, n]), wrap(eng, Cd, T.D, [n, n])
    ms = timeit(P.prepare_op("matmult", [pa, pb], [pc]), 2, warm=1)
    A0 = Ad.clone()
    A0[17, 33] = 0
    want = torch.matmul(A0, Bd)
    want[17, :] = bad
    res["exact_bad_double_4096"] = {"ms": ms, "tflops": 2.0 * n ** 3 / ms / 1e9, "kernel": eng.last_kernel(),
                                    "verified": bool(torch.equal(Cd, want))}
    res["ms"], res["tflops"] = res["dmma_8192"]["ms"], res["dmma_8192"]["tflops"]
    res["verified"] = all(v["verified"] for v in res.values() if isinstance(v, dict))
    return res


def next_rows(eng, dev, peak, reps=20):
    """Two of §8's "next" rows under the driver's clock: the single-pass scan of ONE long row (cumusumover of 2^28
    floats) and minmaximum of a flat 2^28-float ndarray (the body of `minmax`), both verified against torch."""
    g = torch.Generator(device=dev).manual_seed(6)
    n = 2 ** 28
    x = torch.randint(-8, 9, (n,), device=dev, generator=g).float()
    px = wrap(eng, x, T.F, [n])
    out = P.PDL.empty(T.F, [n], eng)
    c0 = eng.launch_count()
    P.run_op("cumusumover", [px], [out])
    launches = eng.launch_count() - c0
    ms = timeit(lambda: P.run_op("cumusumover", [px], [out]), reps)
    got = torch.as_tensor(type("C", (), {"__cuda_array_interface__": {"shape": (n,), "typestr": "<f4",
                          "data": (out.store.ptr, False), "version": 3}})(), device=dev)
    ok = bool(torch.equal(got, torch.cumsum(x.double(), 0).float()))        # integer-valued: every partial sum is exact
    res = {"workload": "next rows: cumusumover float[2^28] (one row) and minmaximum float[2^28] (flat)",
           "cumusumover_1d": {"ms": ms, "gbs": 8 * n / ms / 1e6, "frac": 8 * n / ms / 1e6 / peak, "launches": int(launches),
                              "kernel": eng.last_kernel(), "verified": ok}}
    del got, out
    x[123456789] = 50.0
    x[987654] = -50.0
    outs = [P.PDL.empty(T.F, [], eng), P.PDL.empty(T.F, [], eng), P.PDL.empty(T.IND, [], eng), P.PDL.empty(T.IND, [], eng)]
    prep = P.prepare_op("minmaximum", [px], outs)
    ms = timeit(prep, reps)
    vals = [o.to_numpy().item() for o in prep()]
    res["minmaximum_flat"] = {"ms": ms, "gbs": 4 * n / ms / 1e6, "frac": 4 * n / ms / 1e6 / peak, "kernel": eng.last_kernel(),
                              "verified": vals == [-50.0, 50.0, 987654, 123456789]}
    res["verified"] = res["cumusumover_1d"]["verified"] and res["minmaximum_flat"]["verified"]
    return res
