"""GPU parity proper: CUDA path (through the C-ABI) vs the oracle on seeded inputs at sizes the
oracle finishes in seconds, covering the kernel variants the golden fixtures are too small
to reach (128-bit vector bodies and tails, grid-stride, every reduce mode, split rows)."""
import os

import numpy as np
import pytest

import pdl_b200 as P
from pdl_b200 import types as T, ufunc
from parity import ALL_TYPES, INT_TYPES, assert_same, both, rand_array
from replay import ulp_diff

pytestmark = pytest.mark.gpu

BIOPS_ALL = ["plus", "mult", "minus", "gt", "lt", "le", "ge", "eq", "ne", "spaceship"]
BIOPS_INT = ["or2", "and2", "xor"]


@pytest.fixture()
def engines(cuda_engine, oracle_engine):
    return [cuda_engine, oracle_engine]


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
def test_biops_contiguous_odd_size(engines, t):
    rng = np.random.default_rng(100 + t)
    n = 1_000_003  # not a multiple of any vector width: exercises the tail unit
    a, b = rand_array(rng, t, (n,)), rand_array(rng, t, (n,))
    for op in BIOPS_ALL + (BIOPS_INT if t in T.INTEGER else []):
        (ga, oa), (gb, ob) = both(engines, a, t), both(engines, b, t)
        assert_same(f"{op}-{T.NAMES[t]}", P.run_biop(op, ga, gb), P.run_biop(op, oa, ob))


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
def test_divide_modulo(engines, t):
    rng = np.random.default_rng(200 + t)
    a, b = rand_array(rng, t, (300_001,)), rand_array(rng, t, (300_001,), "pos")
    if t not in T.UNSIGNED:
        b = (b * rng.choice(np.array([-1, 1], dtype=b.dtype), size=b.shape)).astype(b.dtype)
        if t in T.INTEGER:
            b[b == -1] = 3  # INT_MIN / -1 kills the reference (SIGFPE): no reference answer
    for op in ("divide", "modulo"):
        (ga, oa), (gb, ob) = both(engines, a, t), both(engines, b, t)
        assert_same(f"{op}-{T.NAMES[t]}", P.run_biop(op, ga, gb), P.run_biop(op, oa, ob))


@pytest.mark.parametrize("t", INT_TYPES, ids=lambda t: T.NAMES[t])
def test_shifts(engines, t):
    rng = np.random.default_rng(300 + t)
    width = max(32, T.SIZE[t] * 8)
    a = rand_array(rng, t, (100_003,))
    cnt = rng.integers(0, min(width - 2, 62), size=a.shape, endpoint=True).astype(T.NP_DTYPE[t])
    for op in ("shiftleft", "shiftright"):
        (ga, oa), (gb, ob) = both(engines, a, t), both(engines, cnt, t)
        assert_same(f"{op}-{T.NAMES[t]}", P.run_biop(op, ga, gb), P.run_biop(op, oa, ob))


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
def test_bad_values_elementwise(engines, t):
    rng = np.random.default_rng(400 + t)
    a, b = rand_array(rng, t, (513, 257), "small"), rand_array(rng, t, (513, 257), "pos")
    bad = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
    a[rng.random(a.shape) < 0.05] = bad
    b[rng.random(b.shape) < 0.05] = bad
    for op in ("plus", "mult", "divide", "lt", "modulo"):
        (ga, oa), (gb, ob) = both(engines, a, t, True), both(engines, b, t, True)
        assert_same(f"bad-{op}-{T.NAMES[t]}", P.run_biop(op, ga, gb), P.run_biop(op, oa, ob))
    (ga, oa) = both(engines, a, t, True)
    assert_same(f"bad-abs-{T.NAMES[t]}", P.run_ufunc("_rabs", ga), P.run_ufunc("_rabs", oa))


@pytest.mark.parametrize("t", [T.F, T.D], ids=lambda t: T.NAMES[t])
def test_float_ieee_specials_bit_exact(engines, t):
    rng = np.random.default_rng(500 + t)
    n = 200_001
    a, b = rand_array(rng, t, (n,)), rand_array(rng, t, (n,))
    for arr in (a, b):
        k = rng.integers(0, n, size=2000)
        arr[k[:400]] = np.inf; arr[k[400:800]] = -np.inf; arr[k[800:1200]] = np.nan
        arr[k[1200:1600]] = 0.0; arr[k[1600:]] = -0.0
    a[::7] *= np.finfo(T.NP_DTYPE[t]).tiny  # denormal products / quotients
    for op in ("plus", "minus", "mult", "divide"):
        (ga, oa), (gb, ob) = both(engines, a, t), both(engines, b, t)
        assert_same(f"{op}-{T.NAMES[t]}", P.run_biop(op, ga, gb), P.run_biop(op, oa, ob))
    pos = np.abs(a)
    (ga, oa) = both(engines, pos, t)
    assert_same(f"sqrt-{T.NAMES[t]}", P.run_ufunc("sqrt", ga), P.run_ufunc("sqrt", oa))


@pytest.mark.parametrize("t", [T.F, T.D], ids=lambda t: T.NAMES[t])
def test_transcendentals_within_ulp(engines, t):
    # glibc libm vs CUDA libdevice: documented max error of the CUDA functions is <= 2 ulp (float)
    # / <= 2 ulp (double) for these; glibc is <= 1 ulp.  Tolerance: 4 ulp.
    rng = np.random.default_rng(600 + t)
    a = rand_array(rng, t, (100_001,), "small")
    pos = rand_array(rng, t, (100_001,), "pos")
    for op, src in (("sin", a), ("cos", a), ("exp", a), ("log", pos), ("log10", pos)):
        (ga, oa) = both(engines, src, t)
        assert_same(f"{op}-{T.NAMES[t]}", P.run_ufunc(op, ga), P.run_ufunc(op, oa), tol_ulp=4)
    (ga, oa), (gb, ob) = both(engines, pos, t), both(engines, a, t)
    assert_same(f"power-{T.NAMES[t]}", P.run_biop("power", ga, gb), P.run_biop("power", oa, ob), tol_ulp=4)
    assert_same(f"atan2-{T.NAMES[t]}", P.run_biop("atan2", gb, ga), P.run_biop("atan2", ob, oa), tol_ulp=4)


def test_views_strided_dummy_negative(engines):
    rng = np.random.default_rng(700)
    big1, big2 = rand_array(rng, T.D, (4096,)), rand_array(rng, T.D, (4096,))
    outs = []
    for e in engines:
        a = P.PDL.from_numpy(big1, T.D, e).slice("0:-1:2").dummy(1, 1)
        b = P.PDL.from_numpy(big2, T.D, e).slice("0:-1:2").dummy(0, 1)
        outs.append(ufunc.sumover(a * b))
    assert_same("cfg3-small", outs[0], outs[1], tol_ulp=0)
    m = rand_array(rng, T.L, (301, 203))
    outs = []
    for e in engines:
        p = P.PDL.from_numpy(m, T.L, e)
        outs.append(p.slice("-1:0,:") - p.slice(":,-1:0") + p.xchg(0, 1).slice("(5),:").dummy(1, 1))
    assert_same("neg-strides", outs[0], outs[1])
    v = rand_array(rng, T.F, (4, 1, 33, 1, 5))
    w = rand_array(rng, T.F, (1, 6, 1, 7, 5))
    outs = [P.PDL.from_numpy(v, T.F, e) + P.PDL.from_numpy(w, T.F, e) for e in engines]
    assert_same("5d-broadcast", outs[0], outs[1])
    outs = []
    for e in engines:  # unaligned base offset: vector path must fall back per operand
        p = P.PDL.from_numpy(big1, T.D, e)
        outs.append(p.slice("1:-2") + p.slice("2:-1"))
    assert_same("unaligned-offset", outs[0], outs[1])
    outs = []
    for e in engines:  # inplace on a strided view writes through to the parent
        p = P.PDL.from_numpy(big1, T.D, e)
        v2 = p.slice("1:-1:3")
        v2 += 2.5
        outs.append(p + 0)
    assert_same("inplace-through-view", outs[0], outs[1])


NAN_FREE = {"sumover", "prodover", "dsumover", "dprodover", "average", "daverage"}
REDUCERS = ["sumover", "prodover", "dsumover", "average", "daverage", "minimum", "maximum",
            "minimum_ind", "maximum_ind", "andover", "orover", "zcover", "xorover"]
# (n, rows): thread-per-row, warp-per-row, CTA-per-row, and few-long-rows (split into chunks)
SHAPES = [(7, 1500), (33, 700), (1000, 257), (9001, 67), (70_001, 3), (300_007, 1)]


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"n{s[0]}x{s[1]}")
def test_reductions_all_modes(engines, t, shape):
    n, rows = shape
    rng = np.random.default_rng(800 + t + n)
    a = rand_array(rng, t, (rows, n), "exact" if t in (T.F, T.D) else "mixed")
    red = REDUCERS + (["bandover", "borover", "bxorover"] if t in T.INTEGER else [])
    for op in red:
        src = a
        if op in ("prodover",) and t in (T.F, T.D):
            src = np.where(rng.random(a.shape) < 0.5, 1.0, -1.0).astype(a.dtype)  # exact in any order
        (ga, oa) = both(engines, src, t)
        assert_same(f"{op}-{T.NAMES[t]}-{n}x{rows}", getattr(ufunc, op)(ga), getattr(ufunc, op)(oa),
                    nan_equal=op in NAN_FREE)


@pytest.mark.parametrize("t", [T.B, T.S, T.L, T.LL, T.F, T.D], ids=lambda t: T.NAMES[t])
@pytest.mark.parametrize("shape", [(33, 700), (9001, 67), (300_007, 2)], ids=lambda s: f"n{s[0]}x{s[1]}")
def test_reductions_bad(engines, t, shape):
    n, rows = shape
    rng = np.random.default_rng(900 + t + n)
    a = rand_array(rng, t, (rows, n), "exact" if t in (T.F, T.D) else "small")
    bad = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
    a[rng.random(a.shape) < 0.01] = bad
    a[0, :] = bad          # an all-BAD row
    if t in (T.F, T.D):
        a[1, ::5] = np.nan; a[1, 3] = -0.0; a[1, 4] = 0.0
    for op in ("sumover", "average", "daverage", "minimum", "maximum", "minimum_ind", "maximum_ind", "prodover", "orover"):
        src = a
        if op == "prodover" and t in (T.F, T.D):
            src = np.where(np.isin(a, [bad]), bad, np.where(rng.random(a.shape) < 0.5, 1.0, -1.0)).astype(a.dtype)
        (ga, oa) = both(engines, src, t, True)
        assert_same(f"bad-{op}-{T.NAMES[t]}-{n}x{rows}", getattr(ufunc, op)(ga), getattr(ufunc, op)(oa),
                    nan_equal=op in NAN_FREE)


def test_reduction_views(engines):
    rng = np.random.default_rng(1000)
    m = rand_array(rng, T.F, (513, 1025), "exact")
    outs = []
    for e in engines:
        p = P.PDL.from_numpy(m, T.F, e)
        outs.append([ufunc.sumover(p.xchg(0, 1)), ufunc.maximum_ind(p.slice("-1:0:-3,:")),
                     ufunc.minimum(p.slice("5:900:7,2:-1:2").xchg(0, 1)), ufunc.average(p.dummy(0, 3).clump(2)),
                     ufunc.sum(p), ufunc.max(p), ufunc.min(p.slice("1:-1,:"))])
    for k, (g, w) in enumerate(zip(*outs)):
        assert_same(f"reduce-view-{k}", g, w)


def test_float_sum_tolerance(engines):
    """Order-dependent float sums: the device tree order vs the reference's sequential order.
    Stated tolerance: |got - want| <= 2 * n * eps * sum|x| (classic sequential-sum bound)."""
    rng = np.random.default_rng(1100)
    for t in (T.F, T.D):
        a = rand_array(rng, t, (31, 50_000))
        (ga, oa) = both(engines, a, t)
        for op in ("sumover", "average"):
            g, w = getattr(ufunc, op)(ga).to_numpy().astype(np.float64), getattr(ufunc, op)(oa).to_numpy().astype(np.float64)
            scale = np.abs(a.astype(np.float64)).sum(axis=1) / (a.shape[1] if op == "average" else 1)
            bound = 2 * a.shape[1] * np.finfo(T.NP_DTYPE[t]).eps * scale
            assert np.all(np.abs(g - w) <= bound), (op, T.NAMES[t], np.max(np.abs(g - w) / bound))


@pytest.mark.parametrize("t", [T.F, T.D], ids=lambda t: T.NAMES[t])
def test_prodover_stops_when_the_running_product_underflows(engines, t):
    """`tmp *= a; if (tmp == 0) break;` (Ufunc.pd:102-110) also stops when the RUNNING product underflows to zero,
    which depends on the sequential order: [tiny, tiny, inf] is 0 in the reference.  Factors are powers of two, so every
    product is exact in any order until it leaves the representable range — the device must agree bit for bit (sign of
    the zero included), for every cooperation width, with BAD values, for dprodover, and rows that never get near the
    subnormal range must not be touched by the sequential path (same answer either way)."""
    rng = np.random.default_rng(1700 + t)
    dt = T.NP_DTYPE[t]
    dt = np.dtype(dt).type
    tiny = dt(2.0) ** (-100 if t == T.F else -800)
    for op in ("prodover", "dprodover"):
        a = np.array([[tiny, tiny, np.inf, 3.0], [tiny, -tiny, 0.0, np.inf], [2.0, 0.5, 4.0, 0.25]], dtype=dt)
        (ga, oa) = both(engines, a, t)
        assert_same(f"{op}-underflow-small-{T.NAMES[t]}", getattr(ufunc, op)(ga), getattr(ufunc, op)(oa))
        for n, rows in ((5000, 40), (70_000, 6), (300, 500), (40, 3000)):
            k = rng.integers(-40, 12, size=(rows, n))                         # drifts down: most rows underflow somewhere
            k[::3] = rng.integers(-3, 4, size=k[::3].shape)                   # every third row stays in range
            a = (dt(2.0) ** k.astype(dt)) * rng.choice(np.array([-1.0, 1.0], dtype=dt), size=k.shape)
            a[rng.random(a.shape) < 0.001] = np.inf                           # poison after the underflow point
            for bad in (False, True):
                src = a.copy()
                if bad:
                    src[rng.random(src.shape) < 0.02] = np.array(T.DEFAULT_BAD[t]).astype(dt)
                (ga, oa) = both(engines, src, t, bad)
                assert_same(f"{op}-underflow-{T.NAMES[t]}-{n}x{rows}-bad{int(bad)}", getattr(ufunc, op)(ga), getattr(ufunc, op)(oa),
                            nan_equal=True)


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
def test_matmult_exact_kernel(engines, t):
    rng = np.random.default_rng(1200 + t)
    os.environ["PDLB200_MATMULT"] = "exact"
    try:
        for (h, tt, w) in [(65, 67, 66), (130, 300, 70), (1, 513, 5)]:
            a, b = rand_array(rng, t, (h, tt)), rand_array(rng, t, (tt, w))
            (ga, oa), (gb, ob) = both(engines, a, t), both(engines, b, t)
            assert_same(f"matmult-{T.NAMES[t]}-{h}x{tt}x{w}", P.matmult(ga, gb), P.matmult(oa, ob))
    finally:
        os.environ.pop("PDLB200_MATMULT", None)


def test_matmult_bad_and_batched(engines):
    rng = np.random.default_rng(1300)
    for t in (T.L, T.F, T.D):
        a, b = rand_array(rng, t, (4, 70, 90), "small"), rand_array(rng, t, (90, 33), "small")
        bad = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
        a[rng.random(a.shape) < 0.01] = bad
        b[rng.random(b.shape) < 0.01] = bad
        (ga, oa), (gb, ob) = both(engines, a, t, True), both(engines, b, t, True)
        assert_same(f"matmult-bad-{T.NAMES[t]}", P.matmult(ga, gb), P.matmult(oa, ob))
    a, b = rand_array(rng, T.D, (200, 96), "small"), rand_array(rng, T.D, (130, 200), "small")
    outs = [P.matmult(P.PDL.from_numpy(a, T.D, e).xchg(0, 1), P.PDL.from_numpy(b, T.D, e).xchg(0, 1)) for e in engines]
    assert_same("matmult-transposed-views", outs[0], outs[1])


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
def test_matmult_bad_states(engines, t):
    """BAD-mode matmult (Primitive.pd:224-244) on the flag-per-sub-tile kernel: outputs stop either because the running
    value itself equals c's badvalue at a t-tile start (8/16-bit sums wrap onto it all the time) or because a BAD
    a/b element sits in the tile; sparse BAD values (most threads never stop), dense ones (whole CTAs stop early),
    BAD values only in a, only in b, ragged sizes around the reference's tile edges, NaN as the badvalue."""
    rng = np.random.default_rng(1350 + t)
    dt = T.NP_DTYPE[t]
    bad = np.array(T.DEFAULT_BAD[t]).astype(dt)
    for (h, tt, w), dens in (((70, 200, 90), (0.0005, 0.0005)), ((64, 640, 64), (0.0, 0.0)), ((130, 97, 67), (0.02, 0.0)),
                             ((33, 129, 140), (0.0, 0.02)), ((96, 80, 96), (0.3, 0.3)), ((5, 1, 3), (0.2, 0.2))):
        a, b = rand_array(rng, t, (h, tt), "mixed"), rand_array(rng, t, (tt, w), "mixed")
        a[rng.random(a.shape) < dens[0]] = bad
        b[rng.random(b.shape) < dens[1]] = bad
        (ga, oa), (gb, ob) = both(engines, a, t, True), both(engines, b, t, True)
        assert_same(f"matmult-badstates-{T.NAMES[t]}-{h}x{tt}x{w}", P.matmult(ga, gb), P.matmult(oa, ob))
    if t in (T.F, T.D):
        a, b = rand_array(rng, t, (70, 150), "small"), rand_array(rng, t, (150, 66), "small")
        a[rng.random(a.shape) < 0.002] = np.nan            # BAD (the badvalue is NaN) ...
        b[3, 5] = np.inf; b[4, 5] = -np.inf                 # ... and NaNs born in the sum: the running value turns BAD
        res = []
        for e in engines:
            pa, pb = P.PDL.from_numpy(a, t, e), P.PDL.from_numpy(b, t, e)
            pa.set_badvalue(float("nan")); pa.badflag = True
            res.append(P.matmult(pa, pb))
        assert_same(f"matmult-badstates-nanbad-{T.NAMES[t]}", res[0], res[1], nan_equal=True)


def test_matmult_default_path_double(engines):
    """Default double path (tensor-core tiles when eligible): exactly representable inputs make
    every product and partial sum exact, so the result must be BIT-EXACT whatever the order;
    random inputs must be within K*eps*sum|a||b| of the reference's sequential sum."""
    rng = np.random.default_rng(1400)
    h, tt, w = 256, 512, 384
    a = (rng.integers(-64, 64, size=(h, tt)) / 64).astype(np.float64)
    b = (rng.integers(-64, 64, size=(tt, w)) / 64).astype(np.float64)
    (ga, oa), (gb, ob) = both(engines, a, T.D), both(engines, b, T.D)
    assert_same("matmult-default-exact-inputs", P.matmult(ga, gb), P.matmult(oa, ob))
    a, b = rng.uniform(-1, 1, size=(h, tt)), rng.uniform(-1, 1, size=(tt, w))
    (ga, oa), (gb, ob) = both(engines, a, T.D), both(engines, b, T.D)
    g, wnt = P.matmult(ga, gb).to_numpy(), P.matmult(oa, ob).to_numpy()
    bound = 2 * tt * np.finfo(np.float64).eps * (np.abs(a) @ np.abs(b))
    assert np.all(np.abs(g - wnt) <= bound)


@pytest.mark.parametrize("shape", [(128, 16, 128), (384, 1000, 256), (130, 17, 129), (257, 50, 131), (1024, 8, 1024), (128, 1, 128),
                                   (640, 2048, 384), (1300, 330, 1290)],
                         ids=lambda s: "x".join(map(str, s)))
def test_matmult_tma_tiles_edges_and_views(engines, shape):
    """The TMA-staged DMMA kernels (matmult_tma.cu; the stream-K persistent one whenever the tile grid does not fill
    whole waves — every shape here: tiles cut across up to ~10 CTAs, partial accumulators added in k order): full
    tiles, ragged edges in every dim (the hardware zero-fills
    the out-of-range part of a box), and operands that are windows into bigger ndarrays (row pitch != row length).
    Exactly representable inputs -> bit-exact whatever the k order inside a k-tile."""
    cuda = engines[0]
    h, tt, w = shape
    rng = np.random.default_rng(1450 + h + tt + w)
    abig = (rng.integers(-64, 64, size=(h + 6, tt + 10)) / 64).astype(np.float64)
    bbig = (rng.integers(-64, 64, size=(tt + 4, w + 8)) / 64).astype(np.float64)
    for views in (False, True):
        res = []
        for e in engines:
            pa, pb = P.PDL.from_numpy(abig, T.D, e), P.PDL.from_numpy(bbig, T.D, e)
            if views:   # even element offsets keep the 16-byte alignment TMA needs; odd ones must take the cp.async kernel
                pa, pb = pa.slice(f"2:{tt + 1},4:{h + 3}"), pb.slice(f"6:{w + 5},2:{tt + 1}")
            else:
                pa, pb = pa.slice(f"0:{tt - 1},0:{h - 1}").copy(), pb.slice(f"0:{w - 1},0:{tt - 1}").copy()
            res.append(P.matmult(pa, pb))
            if e is cuda and h * w >= 128 * 128 and tt % 2 == 0 and w % 2 == 0:   # row pitches are multiples of 16 bytes
                assert cuda.last_kernel() == "matmult_dmma_tma", cuda.last_kernel()
        assert_same(f"matmult-tma-{shape}-views{views}", res[0], res[1])
    res = []
    for e in engines:   # odd offset: not 16-byte aligned -> the cp.async kernel, same answer
        pa, pb = P.PDL.from_numpy(abig, T.D, e), P.PDL.from_numpy(bbig, T.D, e)
        res.append(P.matmult(pa.slice(f"1:{tt},1:{h}"), pb.slice(f"3:{w + 2},1:{tt}")))
    assert_same(f"matmult-unaligned-{shape}", res[0], res[1])
    # inexact inputs: within the summation-order bound of the reference, identical bits from run to run (the order in
    # which the stream-K parts are added does not depend on timing), and the same bits as the one-CTA-per-tile kernel
    a, b = rng.uniform(-1, 1, size=(h, tt)), rng.uniform(-1, 1, size=(tt, w))
    ga, gb = P.PDL.from_numpy(a, T.D, cuda), P.PDL.from_numpy(b, T.D, cuda)
    r1, r2 = P.matmult(ga, gb).to_numpy(), P.matmult(ga, gb).to_numpy()
    assert np.array_equal(r1, r2)
    assert np.all(np.abs(r1 - a @ b) <= 2 * max(tt, 2) * np.finfo(np.float64).eps * (np.abs(a) @ np.abs(b)) + 1e-300)


@pytest.mark.parametrize("t", [T.SB, T.B, T.S, T.US], ids=lambda t: T.NAMES[t])
def test_small_int_ind_two_phase(engines, t):
    """minimum_ind / maximum_ind of 8/16-bit rows: packed value reduction + first-index search.  Rows with a unique
    extreme at the very end / start / middle, rows that are one long tie, unaligned and strided rows, BAD rows."""
    rng = np.random.default_rng(1460 + t)
    dt = T.NP_DTYPE[t]
    info = np.iinfo(dt)
    for n, rows in ((40_003, 9), (70, 300), (33, 50), (5, 1000)):
        a = rng.integers(max(info.min, -50), min(info.max, 50), size=(rows, n), endpoint=True).astype(dt)
        a[0, :] = 7                                        # one long tie: index 0 wins
        a[1, :] = 3; a[1, n - 1] = 99; a[1, 0] = -99 if info.min < 0 else 0   # unique max at the end, unique min at the start
        a[2, :] = 5; a[2, n // 2] = 100                    # unique max in the middle
        bad = np.array(T.DEFAULT_BAD[t]).astype(dt)
        b = a.copy()
        b[rng.random(b.shape) < 0.05] = bad
        b[3, :] = bad                                      # all BAD -> BAD index
        b[4, : n - 1] = bad                                # only the last element is good
        for arr, flag in ((a, False), (b, True)):
            for op in ("minimum_ind", "maximum_ind"):
                (ga, oa) = both(engines, arr, t, flag)
                assert_same(f"{op}-{T.NAMES[t]}-{n}x{rows}-{flag}", getattr(ufunc, op)(ga), getattr(ufunc, op)(oa))
                if n > 40:
                    assert_same(f"{op}-{T.NAMES[t]}-unaligned", getattr(ufunc, op)(ga.slice("3:-2,:")), getattr(ufunc, op)(oa.slice("3:-2,:")))
                    assert_same(f"{op}-{T.NAMES[t]}-strided", getattr(ufunc, op)(ga.slice("1:-1:3,:")), getattr(ufunc, op)(oa.slice("1:-1:3,:")))
                    assert_same(f"{op}-{T.NAMES[t]}-column", getattr(ufunc, op)(ga.xchg(0, 1)), getattr(ufunc, op)(oa.xchg(0, 1)))


@pytest.mark.parametrize("t", [T.F, T.D], ids=lambda t: T.NAMES[t])
def test_divide_sqrt_fast_path_and_specials(engines, t):
    """divide / sqrt evaluate a 16-byte unit straight-line and fall back per unit: ordinary values, and every kind of
    special operand (zeros of both signs, infinities, NaNs, denormals, huge and tiny magnitudes) mixed into the SAME
    units, must all be bit-identical to the reference's IEEE results."""
    rng = np.random.default_rng(1470 + t)
    dt = T.NP_DTYPE[t]
    n = 400_003
    a = (rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, size=n)).astype(dt)
    b = (rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, size=n)).astype(dt)
    tiny, huge = np.finfo(dt).tiny, np.finfo(dt).max
    specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, tiny, -tiny, tiny / 4, huge, -huge, 1.0, 3.0,
                         float(2.0 ** 41), float(2.0 ** -41), float(2.0 ** 100), float(2.0 ** -100)], dtype=dt)
    idx = rng.choice(n, size=6000, replace=False)
    a[idx[:3000]] = rng.choice(specials, size=3000)
    b[idx[1500:4500]] = rng.choice(specials, size=3000)
    with np.errstate(all="ignore"):
        (ga, oa), (gb, ob) = both(engines, a, t), both(engines, b, t)
        assert_same(f"divide-specials-{T.NAMES[t]}", P.run_biop("divide", ga, gb), P.run_biop("divide", oa, ob))
        assert_same(f"sqrt-specials-{T.NAMES[t]}", P.run_ufunc("sqrt", ga), P.run_ufunc("sqrt", oa))
        bad = np.array(T.DEFAULT_BAD[t]).astype(dt)
        a2, b2 = a.copy(), b.copy()
        a2[rng.random(n) < 0.01] = bad
        b2[rng.random(n) < 0.01] = bad
        (ga, oa), (gb, ob) = both(engines, a2, t, True), both(engines, b2, t, True)
        assert_same(f"divide-specials-bad-{T.NAMES[t]}", P.run_biop("divide", ga, gb), P.run_biop("divide", oa, ob))
        assert_same(f"sqrt-specials-bad-{T.NAMES[t]}", P.run_ufunc("sqrt", ga), P.run_ufunc("sqrt", oa))


@pytest.mark.parametrize("t", [T.CF, T.CD], ids=lambda t: T.NAMES[t])
def test_complex_arithmetic(engines, t):
    """plus minus mult divide on complex float / double: the device restatement of gcc's inline multiply and of
    libgcc's __divsc3 / __divdc3 (scaled Smith) against the oracle, which is built by the reference's own compiler.
    Magnitudes spread over the whole exponent range so that every scaling branch and the Annex G recovery of
    infinities are reached; NaN parts must be NaN on both sides, everything else bit for bit."""
    rng = np.random.default_rng(1480 + t)
    rt = np.float32 if t == T.CF else np.float64
    n = 300_007
    emax = 36 if t == T.CF else 300

    def parts():
        v = rng.standard_normal(n) * 10.0 ** rng.integers(-emax, emax + 1, size=n)
        v[rng.random(n) < 0.02] = 0.0
        sp = rng.random(n) < 0.01
        v[sp] = rng.choice(np.array([np.inf, -np.inf, np.nan, -0.0, np.finfo(rt).tiny / 8, np.finfo(rt).max]), size=int(sp.sum()))
        return v.astype(rt)
    with np.errstate(all="ignore"):
        a = (parts() + 1j * parts()).astype(T.NP_DTYPE[t])
        b = (parts() + 1j * parts()).astype(T.NP_DTYPE[t])
    for op in ("plus", "minus", "mult", "divide"):
        (ga, oa), (gb, ob) = both(engines, a, t), both(engines, b, t)
        g, w = P.run_biop(op, ga, gb), P.run_biop(op, oa, ob)
        assert g.type == w.type and g.dims == w.dims
        assert ulp_diff(g.to_numpy().view(rt), w.to_numpy().view(rt)) == 0, f"{op}-{T.NAMES[t]}"
        # BAD values, broadcasting over a dummy dim, a strided view
        bad = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
        a2 = a[:3000].copy().reshape(30, 100)
        a2[rng.random(a2.shape) < 0.05] = bad
        (ga, oa), (gb, ob) = both(engines, a2, t, True), both(engines, b[:100], t)
        g, w = P.run_biop(op, ga.slice("-1:0:-1,:"), gb), P.run_biop(op, oa.slice("-1:0:-1,:"), ob)
        assert g.badflag == w.badflag and ulp_diff(g.to_numpy().view(rt), w.to_numpy().view(rt)) == 0, f"{op}-{T.NAMES[t]}-bad"


@pytest.mark.parametrize("t", [T.B, T.S, T.L, T.LL, T.F, T.D], ids=lambda t: T.NAMES[t])
def test_scans(engines, t):
    rng = np.random.default_rng(1500 + t)
    for (n, rows) in [(5, 3000), (1000, 37), (40_001, 2), (300_007, 1)]:   # thread-per-row, warp-per-row, chunked
        a = rand_array(rng, t, (rows, n), "exact" if t in (T.F, T.D) else "mixed")
        for bad in (False, True):
            src = a.copy()
            if bad:
                src[rng.random(src.shape) < 0.02] = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
            for op in ("cumusumover", "dcumusumover", "cumuprodover"):
                s2 = src
                if op == "cumuprodover":
                    if t in T.INTEGER:
                        s2 = src
                    else:
                        s2 = np.where(src == np.array(T.DEFAULT_BAD[t]).astype(src.dtype), src,
                                      np.where(rng.random(src.shape) < 0.5, 1.0, -1.0).astype(src.dtype))
                (ga, oa) = both(engines, s2, t, bad)
                assert_same(f"{op}-{T.NAMES[t]}-{n}x{rows}-bad{int(bad)}", getattr(ufunc, op)(ga), getattr(ufunc, op)(oa))
    m = rand_array(rng, t, (64, 300), "exact" if t in (T.F, T.D) else "mixed")
    outs = [ufunc.cumusumover(P.PDL.from_numpy(m, t, e).xchg(0, 1)) for e in engines]      # column layout
    assert_same(f"cumusumover-xchg-{T.NAMES[t]}", outs[0], outs[1])


@pytest.mark.parametrize("t", [T.ULL, T.LL, T.F, T.D], ids=lambda t: T.NAMES[t])
def test_ipow(engines, t):
    from pdl_b200 import ops
    rng = np.random.default_rng(1600 + t)
    a = rand_array(rng, t, (20_001,), "pos" if t in T.INTEGER else "small")
    if t in (T.F, T.D):
        a[a == 0] = 1.5
    e = rng.integers(-6 if t in (T.F, T.D) else 0, 13, size=a.shape).astype(np.int64)
    (ga, oa), (ge, oe) = both(engines, a, t), both(engines, e, T.LL)
    assert_same(f"ipow-{T.NAMES[t]}", ops.ipow(ga, ge), ops.ipow(oa, oe))       # same multiplication chain: bit-exact
    assert_same(f"ipow-{T.NAMES[t]}-scalar", ops.ipow(ga, 3), ops.ipow(oa, 3))


# ---- Bad.pd elementwise ops, axisvalues, inner ------------------------------

@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
def test_bad_producers_consumers(engines, t):
    from pdl_b200 import bad as B
    rng = np.random.default_rng(900 + t)
    shape = (257, 1031)   # PDL dims [1031, 257]: tile bodies, tails and row decode
    a = rand_array(rng, t, shape, "small")
    badv = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
    ab = a.copy(); ab[rng.random(shape) < 0.03] = badv
    mask = (rng.random(shape) < 0.2).astype(np.int32) * rng.integers(-3, 3, size=shape, endpoint=True).astype(np.int32)
    rowmask = (rng.random(shape[1]) < 0.5).astype(np.int32)
    other = rand_array(rng, t, shape, "pos")
    for flagged, arr in ((False, a), (True, ab)):
        tag = f"{T.NAMES[t]}-{'bad' if flagged else 'good'}"
        (ga, oa) = both(engines, arr, t, flagged)
        for op in ("isbad", "isgood", "isnan"):
            assert_same(f"{op}-{tag}", getattr(B, op)(ga), getattr(B, op)(oa))
        (gm, om) = both(engines, mask, T.L)
        assert_same(f"setbadif-{tag}", B.setbadif(ga, gm), B.setbadif(oa, om))
        (gr, orr) = both(engines, rowmask, T.L)
        assert_same(f"setbadif-row-{tag}", B.setbadif(ga, gr), B.setbadif(oa, orr))
        assert_same(f"setvaltobad-{tag}", B.setvaltobad(ga, 3), B.setvaltobad(oa, 3))
        assert_same(f"setbadtoval-{tag}", B.setbadtoval(ga, 5), B.setbadtoval(oa, 5))
        (gb, ob) = both(engines, other, t)
        assert_same(f"badmask-{tag}", B.badmask(ga, gb), B.badmask(oa, ob))
        (gc, oc) = both(engines, ab, t, True)
        assert_same(f"copybad-{tag}", B.copybad(gb, gc), B.copybad(ob, oc))
        # strided / transposed views are read in place
        assert_same(f"setbadtoval-view-{tag}", B.setbadtoval(ga.slice("1:-1:3,:").xchg(0, 1), 1),
                    B.setbadtoval(oa.slice("1:-1:3,:").xchg(0, 1), 1))
        # in place
        gi, oi = ga.copy(), oa.copy()
        gi.badflag = oi.badflag = flagged
        B.setbadtoval(gi.inplace(), 2); B.setbadtoval(oi.inplace(), 2)
        assert_same(f"setbadtoval-inplace-{tag}", gi, oi)


@pytest.mark.parametrize("t", [T.F, T.D], ids=lambda t: T.NAMES[t])
def test_nonfinite_to_bad(engines, t):
    from pdl_b200 import bad as B
    rng = np.random.default_rng(950 + t)
    n = 300_001
    a = rand_array(rng, t, (n,))
    k = rng.integers(0, n, size=900)
    special = a.copy()
    special[k[:300]] = np.nan; special[k[300:600]] = np.inf; special[k[600:]] = -np.inf
    only_inf = a.copy(); only_inf[k[:10]] = np.inf
    for name, arr in (("finite", a), ("special", special), ("inf", only_inf)):
        for flagged in (False, True):
            (ga, oa) = both(engines, arr, t, flagged)
            for op in ("setnantobad", "setinftobad", "setnonfinitetobad", "setbadtonan", "isnan"):
                assert_same(f"{op}-{name}-{T.NAMES[t]}-{flagged}", getattr(B, op)(ga), getattr(B, op)(oa))


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
def test_axisvalues_constructors(engines, t):
    from pdl_b200 import basic
    for e_out in ([basic.sequence(t, 1000, 37, engine=e) for e in engines],
                  [basic.sequence(t, 100_003, engine=e) for e in engines],
                  [basic.zeroes(t, 513, 7, engine=e) for e in engines],
                  [basic.ones(t, 513, 7, engine=e) for e in engines]):
        assert_same(f"ctor-{T.NAMES[t]}", e_out[0], e_out[1])
    rng = np.random.default_rng(980 + t)
    a = rand_array(rng, t, (5, 33, 129), "small")
    (ga, oa) = both(engines, a, t)
    for f in (basic.xvals, basic.yvals, basic.zvals):
        assert_same(f"{f.__name__}-{T.NAMES[t]}", f(ga), f(oa))
        assert_same(f"{f.__name__}-view-{T.NAMES[t]}", f(ga.slice("1:-1:2,:,:").xchg(0, 2)), f(oa.slice("1:-1:2,:,:").xchg(0, 2)))


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
def test_inner(engines, t):
    rng = np.random.default_rng(990 + t)
    # exact-by-construction values: every product and partial sum is representable, so the 80-bit
    # sequential sum of the reference and the device's chunked sums must agree bit for bit
    flav = "exact" if t in (T.F, T.D) else "small"
    for shape in ((3, 70_001), (2000, 33), (7, 300), (1, 5)):
        a, b = rand_array(rng, t, shape, flav), rand_array(rng, t, shape, flav)
        if T.SIZE[t] == 1:
            a, b = a % 2, b % 2          # keep the row sums inside the 8-bit output type
        elif T.SIZE[t] == 2 and shape[1] > 1000:
            a, b = a % 2, b % 2
        (ga, oa), (gb, ob) = both(engines, a, t), both(engines, b, t)
        assert_same(f"inner-{T.NAMES[t]}-{shape}", P.inner(ga, gb), P.inner(oa, ob))
        # b broadcast along the rows, and BAD rows
        (gv, ov) = both(engines, b[0], t)
        assert_same(f"inner-bcast-{T.NAMES[t]}-{shape}", P.inner(ga, gv), P.inner(oa, ov))
        ab = a.copy(); ab[rng.random(shape) < 0.001] = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
        (gx, ox) = both(engines, ab, t, True)
        assert_same(f"inner-bad-{T.NAMES[t]}-{shape}", P.inner(gx, gb), P.inner(ox, ob))
    # the cfg3 shape: [N,1] x [1,M] dummies, fused mult -> sumover
    x, y = rand_array(rng, t, (300,), flav), rand_array(rng, t, (200,), flav)
    if T.SIZE[t] <= 2:
        x, y = x % 2, y % 2
    (gx, ox), (gy, oy) = both(engines, x, t), both(engines, y, t)
    fused = [P.inner(px.dummy(1, 1), py.dummy(0, 1)) for px, py in ((gx, gy), (ox, oy))]
    assert_same(f"inner-outer-{T.NAMES[t]}", fused[0], fused[1])
    if t in (T.L, T.LL, T.F, T.D):   # same result type as sumover there (no int+ widening)
        unfused = ufunc.sumover(gx.dummy(1, 1) * gy.dummy(0, 1))
        assert fused[0].to_numpy().tobytes() == unfused.to_numpy().tobytes()


def test_inner_double_random_within_tolerance(engines):
    """Inexact doubles: the reference sums in x87 80-bit, the device in a compensated double pair;
    both are within 1 ulp of the exact sum, so they differ by at most 1 ulp of the result
    (|result| bounded away from cancellation by using positive values)."""
    rng = np.random.default_rng(77)
    a = rng.random((4, 100_000)) + 0.5
    b = rng.random((4, 100_000)) + 0.5
    (ga, oa), (gb, ob) = both(engines, a, T.D), both(engines, b, T.D)
    assert_same("inner-double-random", P.inner(ga, gb), P.inner(oa, ob), tol_ulp=1)
    af, bf = a.astype(np.float32), b.astype(np.float32)
    (ga, oa), (gb, ob) = both(engines, af, T.F), both(engines, bf, T.F)
    assert_same("inner-float-random", P.inner(ga, gb), P.inner(oa, ob), tol_ulp=1)


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
def test_minmaximum(engines, t):
    rng = np.random.default_rng(1100 + t)
    for shape in ((3, 70_001), (2000, 33), (5, 4097), (1, 300_007)):
        a = rand_array(rng, t, shape, "small" if shape[1] > 1000 else "mixed")   # small: many ties -> first index must win
        (ga, oa) = both(engines, a, t)
        for k, (g, o) in enumerate(zip(ufunc.minmaximum(ga), ufunc.minmaximum(oa))):
            assert_same(f"minmaximum[{k}]-{T.NAMES[t]}-{shape}", g, o)
        ab = a.copy()
        ab[rng.random(shape) < 0.3] = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
        ab[0, :] = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])           # an all-BAD row
        (gb, ob) = both(engines, ab, t, True)
        for k, (g, o) in enumerate(zip(ufunc.minmaximum(gb), ufunc.minmaximum(ob))):
            assert_same(f"minmaximum-bad[{k}]-{T.NAMES[t]}-{shape}", g, o)
        # strided / transposed view
        for k, (g, o) in enumerate(zip(ufunc.minmaximum(ga.xchg(0, 1)), ufunc.minmaximum(oa.xchg(0, 1)))):
            assert_same(f"minmaximum-xchg[{k}]-{T.NAMES[t]}-{shape}", g, o)
    if t in (T.F, T.D):
        a = rand_array(rng, t, (64, 5000))
        a[rng.random(a.shape) < 0.2] = np.nan
        a[3, :] = np.nan                                   # all-NaN row: BAD outputs + badflag without any BAD input
        a[5, ::2] = 0.0; a[5, 1::2] = -0.0
        (ga, oa) = both(engines, a, t)
        for k, (g, o) in enumerate(zip(ufunc.minmaximum(ga), ufunc.minmaximum(oa))):
            assert_same(f"minmaximum-nan[{k}]-{T.NAMES[t]}", g, o)
        assert ufunc.minmaximum(ga)[0].badflag
        assert ufunc.minmax(ga) == ufunc.minmax(oa)
        # the data-dependent flag travels through a pinned slot (PDLB200_TRANS_DEFER_ANYBAD): the call itself does not
        # synchronise; the first question about the bad state settles it, for every output and every view
        outs = ufunc.minmaximum(ga)
        assert all(o.store._pend is not None for o in outs)
        view = outs[2].slice("0:3")
        assert view.badflag and all(o.store._pend is None for o in outs) and all(o.badflag for o in outs)
        good = np.ascontiguousarray(a[:3])
        good[np.isnan(good)] = 1.0
        outs = ufunc.minmaximum(P.PDL.from_numpy(good, t, engines[0]))
        assert outs[0].store._pend is not None and not outs[0].badflag and not outs[3].badflag
        outs = ufunc.minmaximum(ga)
        outs[0].badflag = False                            # an explicit setting wins over a flag still in flight
        assert outs[0].store._pend is None and not outs[0].badflag and outs[1].badflag


@pytest.mark.parametrize("t", [T.F, T.D], ids=lambda t: T.NAMES[t])
def test_magnover(engines, t):
    rng = np.random.default_rng(1200 + t)
    for shape in ((3, 70_001), (2000, 33), (1, 300_007)):
        a = rand_array(rng, t, shape)
        (ga, oa) = both(engines, a, t)
        assert_same(f"magnover-{T.NAMES[t]}-{shape}", ufunc.magnover(ga), ufunc.magnover(oa), tol_ulp=1)
        ab = a.copy(); ab[rng.random(shape) < 0.3] = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
        ab[0, :] = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
        (gb, ob) = both(engines, ab, t, True)
        assert_same(f"magnover-bad-{T.NAMES[t]}-{shape}", ufunc.magnover(gb), ufunc.magnover(ob), tol_ulp=1)
    # exactly representable: 3-4-5 style rows must be bit-exact
    a = np.array([[3, 4] * 500, [5, 12] * 500, [0, 0] * 500], dtype=T.NP_DTYPE[t])
    (ga, oa) = both(engines, a, t)
    assert_same(f"magnover-exact-{T.NAMES[t]}", ufunc.magnover(ga), ufunc.magnover(oa))


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
def test_outer(engines, t):
    rng = np.random.default_rng(1300 + t)
    a, b = rand_array(rng, t, (3, 5001), "small"), rand_array(rng, t, (777,), "small")
    (ga, oa), (gb, ob) = both(engines, a, t), both(engines, b, t)
    assert_same(f"outer-{T.NAMES[t]}", P.outer(ga, gb), P.outer(oa, ob))
    ab = a.copy(); ab[rng.random(a.shape) < 0.05] = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
    (gx, ox) = both(engines, ab, t, True)
    assert_same(f"outer-bad-{T.NAMES[t]}", P.outer(gx, gb), P.outer(ox, ob))
    # same numbers as the operator on dummy dims (what the reference's docs say outer is)
    (g1, _), (g2, _) = both(engines, a[0], t), both(engines, b, t)
    assert P.outer(g1, g2).to_numpy().tobytes() == (g1.dummy(1, 1) * g2.dummy(0, 1)).to_numpy().tobytes()


def test_inner_magnover_nonfinite_and_overflow(engines):
    """The reference accumulates in long double and carries inf / NaN through; the compensated device sum must
    not turn an infinite total into NaN (round-1 advisor finding)."""
    inf, nan = np.inf, np.nan
    rows = np.array([[1.0, inf, 2.0, 3.0], [1e308, 1e308, 1e308, 1.0], [inf, -inf, 1.0, 1.0],
                     [-inf, 5.0, 1.0, 1.0], [nan, 1.0, 2.0, 3.0], [1.0, 2.0, 3.0, 4.0]])
    ones = np.ones_like(rows)
    for t in (T.D, T.F):
        r = rows if t == T.D else np.where(np.abs(rows) == 1e308, 3e38, rows)
        (ga, oa), (gb, ob) = both(engines, r, t), both(engines, ones * 2, t)
        assert_same(f"inner-nonfinite-{T.NAMES[t]}", P.inner(ga, gb), P.inner(oa, ob), nan_equal=True)
        assert_same(f"magnover-nonfinite-{T.NAMES[t]}", ufunc.magnover(ga), ufunc.magnover(oa), nan_equal=True)
    big = np.tile(rows, (1, 5000))                         # long rows: the warp/CTA kernels and their merges
    (ga, oa), (gb, ob) = both(engines, big, T.D), both(engines, np.ones_like(big), T.D)
    assert_same("inner-nonfinite-long", P.inner(ga, gb), P.inner(oa, ob), nan_equal=True)
    assert_same("magnover-nonfinite-long", ufunc.magnover(ga), ufunc.magnover(oa), nan_equal=True)


def test_matmult_more_than_65535_batches(engines):
    """Batches of small matrices, (3,3,70000) x (3,3,70000): the broadcast positions exceed gridDim.z."""
    rng = np.random.default_rng(77)
    for t in (T.D, T.F, T.L):
        a = rand_array(rng, t, (70000, 3, 3), "small")
        b = rand_array(rng, t, (70000, 3, 3), "small")
        (ga, oa), (gb, ob) = both(engines, a, t), both(engines, b, t)
        assert_same(f"matmult-batch-{T.NAMES[t]}", P.matmult(ga, gb), P.matmult(oa, ob))
    a = rng.integers(-8, 8, size=(66000, 64, 16)).astype(np.float64)      # DMMA-eligible tiles, > 65535 of them
    b = rng.integers(-8, 8, size=(16, 64)).astype(np.float64)
    (ga, oa), (gb, ob) = both(engines, a, T.D), both(engines, b, T.D)
    assert_same("matmult-dmma-batch", P.matmult(ga, gb), P.matmult(oa, ob))


def test_flowing_product_fused_into_sumover(engines):
    cuda = engines[0]
    rng = np.random.default_rng(12)
    for t in (T.D, T.F):
        # double: products and sums exact (SURVEY.md 8(d) cfg3 values); float: small integers, so that the unfused
        # float accumulation is exact too (otherwise fused, which accumulates in double, is the MORE accurate one)
        big1, big2 = rng.integers(-1024, 1024, size=4000) / 256, rng.integers(-1024, 1024, size=1000) / 256
        if t == T.F:
            big1, big2 = np.round(big1 * 2), np.round(big2 * 2)
        res = []
        for e in engines:
            a = P.PDL.from_numpy(big1, t, e).slice("0:-1:2").dummy(1, 1)
            b = P.PDL.from_numpy(big2, t, e).slice("0:-1:2").dummy(0, 1)
            l0 = cuda.launch_count()
            fused = ufunc.sumover(a.flowing() * b)
            if e is cuda:
                assert cuda.last_kernel() == "inner" and cuda.launch_count() - l0 <= 2    # inner (+ its finish), no mult
            res.append((fused, ufunc.sumover(a * b)))
        assert_same(f"fused-{T.NAMES[t]}", res[0][0], res[1][0])
        assert_same(f"fused-vs-unfused-{T.NAMES[t]}", res[0][0], res[0][1])


@pytest.mark.parametrize("t", [T.SB, T.B, T.S, T.US], ids=lambda t: T.NAMES[t])
def test_small_int_divide_sqrt_exhaustive(engines, t):
    """8-bit divide runs on one approximate float division and 8/16-bit sqrt on the float square root:
    check EVERY operand pair / value against the oracle's integer division and double sqrt."""
    dt = T.NP_DTYPE[t]
    info = np.iinfo(dt)
    vals = np.arange(info.min, info.max + 1, dtype=np.int64)
    if T.SIZE[t] == 1:
        a = np.repeat(vals, vals.size).astype(dt)
        b = np.tile(vals, vals.size).astype(dt)
        keep = b != 0
        if t == T.SB:
            keep &= ~((a == -128) & (b == -1))          # INT_MIN / -1 kills the reference (SIGFPE)
        a, b = a[keep], b[keep]
        (ga, oa), (gb, ob) = both(engines, a, t), both(engines, b, t)
        assert_same(f"divide-exhaustive-{T.NAMES[t]}", P.run_biop("divide", ga, gb), P.run_biop("divide", oa, ob))
    v = vals[vals >= 0].astype(dt)
    (gv, ov) = both(engines, v, t)
    assert_same(f"sqrt-exhaustive-{T.NAMES[t]}", P.run_ufunc("sqrt", gv), P.run_ufunc("sqrt", ov))


@pytest.mark.parametrize("t", [T.L, T.UL] if hasattr(T, "UL") else [T.L], ids=lambda t: T.NAMES[t])
def test_int32_sqrt_float_root_plus_integer_checks(engines, t):
    """32-bit integer sqrt = float root corrected by two integer checks: every perfect square and its two neighbours up
    to the type's maximum, the maximum itself, random values — against the oracle's (T)sqrt((double)a)."""
    dt = T.NP_DTYPE[t]
    hi = np.iinfo(dt).max
    k = np.arange(0, int(np.sqrt(float(hi))) + 1, dtype=np.int64)
    v = np.concatenate([k * k - 1, k * k, k * k + 1, [hi, hi - 1, 0, 1, 2, 3]])
    v = v[(v >= 0) & (v <= hi)]
    rng = np.random.default_rng(1800 + t)
    v = np.concatenate([v, rng.integers(0, hi, size=200_000, dtype=np.int64)]).astype(dt)
    (gv, ov) = both(engines, v, t)
    assert_same(f"sqrt-int32-{T.NAMES[t]}", P.run_ufunc("sqrt", gv), P.run_ufunc("sqrt", ov))


def test_cabi_error_returns_on_device(cuda_engine):
    """Malformed descriptors come back as error codes + messages through the C-ABI, never as a crash or a
    silently wrong launch: missing `anybad` for the ops that need it, wrong fixed parameter types, a type
    outside the device matrix, an op id that does not exist."""
    import ctypes as C
    from pdl_b200 import _abi, trans
    lib, err = cuda_engine.lib, C.create_string_buffer(512)
    a = P.PDL.from_numpy(np.arange(8, dtype=np.float32), T.F, cuda_engine)
    o = P.PDL.empty(T.F, [8], cuda_engine)

    def desc(name, pdls, named=None, ttype=T.F):
        spec = trans.SPECS[name]
        bc = trans._broadcast(pdls, [len(p.realdims) for p in spec.pars], [False] * len(pdls), name)
        tr = trans._build_trans(spec, ttype, pdls, bc, named or {}, False)
        return tr

    # setnantobad / minmaximum without anybad
    tr = desc("setnantobad", [a, o]); tr.anybad = None
    assert lib.pdlb200_readdata(C.byref(tr), err, 512) == _abi.EINVAL and b"anybad" in err.value
    s = [P.PDL.empty(T.F, [], cuda_engine), P.PDL.empty(T.F, [], cuda_engine),
         P.PDL.empty(T.IND, [], cuda_engine), P.PDL.empty(T.IND, [], cuda_engine)]
    tr = desc("minmaximum", [a] + s, {"ind": [8], "rinc": [1]}); tr.anybad = None
    assert lib.pdlb200_readdata(C.byref(tr), err, 512) == _abi.EINVAL and b"anybad" in err.value
    # setbadif with a float mask, isbad with a float output
    tr = desc("setbadif", [a, a, o])
    assert lib.pdlb200_readdata(C.byref(tr), err, 512) == _abi.EINVAL and b"mask" in err.value
    tr = desc("isbad", [a, o])
    assert lib.pdlb200_readdata(C.byref(tr), err, 512) == _abi.EINVAL and b"int" in err.value
    # magnover of an integer type, an unknown op, a type outside the matrix
    li = P.PDL.from_numpy(np.arange(8, dtype=np.int32), T.L, cuda_engine)
    tr = desc("magnover", [li, P.PDL.empty(T.L, [], cuda_engine)], {"ind": [8], "rinc": [1]}, ttype=T.L)
    assert lib.pdlb200_readdata(C.byref(tr), err, 512) == _abi.EUNSUPPORTED
    tr = desc("plus", [a, a, o]); tr.op = 49
    assert lib.pdlb200_readdata(C.byref(tr), err, 512) == _abi.EINVAL and b"unknown op" in err.value
    tr = desc("plus", [a, a, o]); tr.datatype = 11            # long double: x87 80-bit, no device representation
    assert lib.pdlb200_readdata(C.byref(tr), err, 512) == _abi.EUNSUPPORTED and b"no device representation" in err.value
    tr = desc("sqrt", [a, o]); tr.datatype = 12               # complex float is on the device path for + - * / only
    assert lib.pdlb200_readdata(C.byref(tr), err, 512) == _abi.EUNSUPPORTED and b"no device representation" in err.value
    # and the engine turns them into PDLError for the host mirror
    with pytest.raises(P.PDLError):
        cuda_engine.readdata(tr)
