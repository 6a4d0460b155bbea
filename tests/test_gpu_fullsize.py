"""GPU: the BASELINE.json configurations at their FULL sizes.  The oracle cannot finish these in
seconds, so they are checked through size-independent properties and against closed forms /
independent device computations (torch is used only as an independent checker here):
  cfg1  $x = $y + $c, 2048x2048 double                  -> bitwise equal to torch's IEEE add, linear in c
  cfg2  sumover/average/minimum, float[16384,65536], 1% BAD -> exact integer-valued sums: whole == sum of halves,
        average == sum/count (one IEEE divide), minimum == masked torch.min, a sample of rows vs the oracle
  cfg3  [N,1]*[1,M] on strided slices + dummy dims, N=M=32768 double, then sumover -> closed form sum(a)*b[j]
  cfg4  matmult 8192^3 double, exactly representable inputs -> bit-exact vs cuBLAS DGEMM and vs row-sum identities
  cfg5  sum / max of 2^33 floats in {-1,0,1}           -> exact known totals
"""
import numpy as np
import pytest

import pdl_b200 as P
from pdl_b200 import types as T, ufunc

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def wrap(eng, t, typ, dims):
    return P.PDL(eng, eng.wrap(t.data_ptr(), t.numel() * t.element_size(), t), typ, dims)


def as_torch(p, dtype):
    """Zero-copy torch view of a contiguous device ndarray."""
    class _CAI:
        def __init__(self, ptr, n, typestr):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}
    ts = {torch.float32: "<f4", torch.float64: "<f8", torch.int64: "<i8"}[dtype]
    return torch.as_tensor(_CAI(p.store.ptr + p.offs * T.SIZE[p.datatype], p.nelem, ts), device="cuda")


def test_cfg1_plus_bit_exact(cuda_engine):
    g = torch.Generator(device="cuda").manual_seed(11)
    n = 2048 * 2048
    y = torch.randint(-2**20, 2**20, (n,), device="cuda", generator=g).double() / 1024
    c = torch.randint(-2**20, 2**20, (n,), device="cuda", generator=g).double() / 1024
    x = wrap(cuda_engine, y, T.D, [2048, 2048]) + wrap(cuda_engine, c, T.D, [2048, 2048])
    assert x.dims == [2048, 2048] and x.type == "double"
    assert torch.equal(as_torch(x, torch.float64), y + c)
    d = wrap(cuda_engine, y, T.D, [2048, 2048]) / wrap(cuda_engine, c + 0.5, T.D, [2048, 2048])
    assert torch.equal(as_torch(d, torch.float64), y / (c + 0.5))          # IEEE divide, no FMA contraction


def test_cfg2_reductions_full_size(cuda_engine, oracle_engine):
    import bench
    rows, n = bench.ROWS, bench.N_DIM
    x = bench.generate_device(torch, 0, rows, torch.device("cuda"))
    bad = torch.finfo(torch.float32).min
    a = wrap(cuda_engine, x, T.F, [n, rows]).set_badflag(True)
    s, avg, mn = ufunc.sumover(a), ufunc.average(a), ufunc.minimum(a)
    assert s.dims == [rows] and s.badflag and avg.type == "float"
    good = x != bad
    want_sum = torch.where(good, x, torch.zeros_like(x)).sum(dim=1)        # values are small integers: exact in any order
    cnt = good.sum(dim=1)
    ts, ta, tm = as_torch(s, torch.float32), as_torch(avg, torch.float32), as_torch(mn, torch.float32)
    assert torch.equal(ts, want_sum)
    assert torch.equal(ta, want_sum / cnt.float())                           # one IEEE divide of exact operands
    assert torch.equal(tm, torch.where(good, x, torch.full_like(x, float("inf"))).min(dim=1).values)
    # linearity / additivity: the whole row equals the sum of its two halves (views, no copies)
    h1 = ufunc.sumover(a.slice(f"0:{n // 2 - 1},:"))
    h2 = ufunc.sumover(a.slice(f"{n // 2}:-1,:"))
    assert torch.equal(as_torch(h1 + h2, torch.float32), ts)
    # checksum of checksums vs the reference's own run recorded in oracle/ref_bench.pl terms: rows 0..511
    assert float(ts[:512].double().sum()) == -5815.0 and float(tm[:512].double().sum()) == -4096.0
    # a sample of rows against the oracle
    host = bench.sample_numpy(4096, 16)
    pa = P.PDL.from_numpy(host, T.F, oracle_engine).set_badflag(True)
    for op, got in (("sumover", s), ("average", avg), ("minimum", mn)):
        assert got.slice("4096:4111").to_numpy().tobytes() == getattr(ufunc, op)(pa).to_numpy().tobytes()


def test_cfg3_outer_product_views(cuda_engine):
    N = M = 32768
    g = torch.Generator(device="cuda").manual_seed(13)
    big1 = torch.randint(-1024, 1024, (2 * N,), device="cuda", generator=g).double() / 256
    big2 = torch.randint(-1024, 1024, (2 * M,), device="cuda", generator=g).double() / 256
    a = wrap(cuda_engine, big1, T.D, [2 * N]).slice("0:-1:2").dummy(1, 1)      # [N,1], stride 2
    b = wrap(cuda_engine, big2, T.D, [2 * M]).slice("0:-1:2").dummy(0, 1)      # [1,M]
    prod = a * b
    assert prod.dims == [N, M]
    tp = as_torch(prod, torch.float64).view(M, N)
    for j in (0, 1, 777, M - 1):                                               # spot rows of the 8 GiB product
        assert torch.equal(tp[j], big1[::2] * big2[2 * j])
    sums = ufunc.sumover(prod)
    # every product and partial sum is exactly representable: sum_i a_i*b_j == (sum_i a_i)*b_j bitwise
    assert torch.equal(as_torch(sums, torch.float64), big1[::2].sum() * big2[::2])


def test_cfg4_matmult_8192(cuda_engine):
    n = 8192
    g = torch.Generator(device="cuda").manual_seed(17)
    A = torch.randint(-64, 64, (n, n), device="cuda", generator=g).double() / 64
    B = torch.randint(-64, 64, (n, n), device="cuda", generator=g).double() / 64
    c = P.matmult(wrap(cuda_engine, A, T.D, [n, n]), wrap(cuda_engine, B, T.D, [n, n]))
    assert cuda_engine.last_kernel().startswith("matmult_dmma")
    tc = as_torch(c, torch.float64).view(n, n)
    assert torch.equal(tc, A @ B)                                              # exact inputs: any summation order agrees
    assert torch.equal(tc.sum(dim=1), A @ B.sum(dim=1))                        # C.1 == A.(B.1): still exact


def test_cfg5_full_array_sum_max(cuda_engine):
    n = 2 ** 33
    if torch.cuda.mem_get_info()[0] < 40 * 2 ** 30:
        pytest.skip("needs 32 GiB of free HBM")
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(19)
    step = 2 ** 28
    total = 0
    for i in range(0, n, step):
        blk = torch.randint(-1, 2, (step,), device="cuda", generator=g, dtype=torch.int8)
        total += int(blk.sum(dtype=torch.int64))
        x[i:i + step] = blk.float()
    x[5_000_000_001] = 7.0
    total += 7 - int(x[5_000_000_001].item() == 7.0) * 0
    px = wrap(cuda_engine, x, T.F, [n])
    want = float(x.double().sum().item())
    assert ufunc.sum(px).sclr() == want                                        # random walk: |partial sums| << 2^24
    assert ufunc.max(px).sclr() == 7.0
    assert ufunc.maximum_ind(px).sclr() == 5_000_000_001                       # index beyond 2^32
    assert ufunc.min(px).sclr() == -1.0


def test_long_row_scan_chunked(cuda_engine):
    """cumusumover over one very long row whose length is NOT a whole number of 16-byte vectors (single-pass scan of
    the vectors + the tail kernel carrying on from the row's last descriptor), and two rows with an odd pitch (the
    chunked three-pass scan): integer-valued floats make every partial sum exact, so any order must agree bitwise."""
    g = torch.Generator(device="cuda").manual_seed(17)
    n = 2**27 + 12345
    x = torch.randint(-8, 9, (n,), device="cuda", generator=g).float()
    got = as_torch(ufunc.cumusumover(wrap(cuda_engine, x, T.F, [n])), torch.float32)
    assert torch.equal(got, torch.cumsum(x.double(), 0).float())
    # two rows with 1% BAD: BAD elements give BAD out and do not contribute
    n2 = 2**25 + 3
    y = torch.randint(-8, 9, (2 * n2,), device="cuda", generator=g).float()
    bad = torch.rand(2 * n2, device="cuda", generator=g) < 0.01
    y[bad] = -9999.0
    py = wrap(cuda_engine, y, T.F, [n2, 2])
    py.set_badvalue(-9999.0)
    py.badflag = True
    out = ufunc.cumusumover(py)
    assert out.badflag
    got = as_torch(out, torch.float32).view(2, n2)
    want = torch.cumsum(torch.where(bad, torch.zeros_like(y), y).view(2, n2).double(), 1).float()
    want[bad.view(2, n2)] = float(out.badvalue)
    assert torch.equal(got, want)


@pytest.mark.parametrize("mode", ["onepass", "3pass"])
@pytest.mark.parametrize("case", ["float-2rows-bad", "double-prod", "int64-sum", "int32-3rows-tail"])
def test_lookback_scan_variants(cuda_engine, case, mode, monkeypatch):
    """Long-row scans through BOTH schemes — the single-pass look-back kernel (scan_onepass.cuh) and, with
    PDLB200_SCAN=3pass, the chunked three-pass scheme it falls back to (chunk totals, exclusive scan of the totals,
    scan with carry-in): BAD elements, several rows, products, 64-bit accumulators, rows ending inside a tile —
    against torch's cumsum/cumprod on inputs whose partial results are exactly representable."""
    monkeypatch.setenv("PDLB200_SCAN", mode)
    g = torch.Generator(device="cuda").manual_seed(23)
    if case == "float-2rows-bad":
        n = 2**22 + 8                                   # rows stay 16-byte aligned
        y = torch.randint(-8, 9, (2 * n,), device="cuda", generator=g).float()
        bad = torch.rand(2 * n, device="cuda", generator=g) < 0.01
        y[bad] = -9999.0
        py = wrap(cuda_engine, y, T.F, [n, 2]).set_badvalue(-9999.0).set_badflag(True)
        out = ufunc.cumusumover(py)
        assert cuda_engine.last_kernel() == "scan_cumusumover" and out.badflag
        want = torch.cumsum(torch.where(bad, torch.zeros_like(y), y).view(2, n).double(), 1).float()
        want[bad.view(2, n)] = float(out.badvalue)
        assert torch.equal(as_torch(out, torch.float32).view(2, n), want)
    elif case == "double-prod":
        n = 2**21
        # +-1 and a few +-2: the running product stays a power of two, exact in double
        y = torch.where(torch.rand(n, device="cuda", generator=g) < 0.5, -1.0, 1.0).double()
        y[torch.randint(0, n, (40,), device="cuda", generator=g)] = 2.0
        out = ufunc.cumuprodover(wrap(cuda_engine, y, T.D, [n]))
        assert torch.equal(as_torch(out, torch.float64), torch.cumprod(y, 0))
    elif case == "int64-sum":
        n = 2**21 + 2
        y = torch.randint(-2**40, 2**40, (n,), device="cuda", generator=g)
        out = ufunc.cumusumover(wrap(cuda_engine, y, T.LL, [n]))
        assert torch.equal(as_torch(out, torch.int64), torch.cumsum(y, 0))
    else:
        n = 2**22 + 4                                   # int32 rows: 4-element alignment, last tile is partial
        y = torch.randint(-1000, 1000, (3 * n,), device="cuda", generator=g, dtype=torch.int32)
        out = ufunc.cumusumover(wrap(cuda_engine, y, T.L, [n, 3]))
        got = torch.as_tensor(type("C", (), {"__cuda_array_interface__": {"shape": (3 * n,), "typestr": "<i4",
                              "data": (out.store.ptr, False), "version": 3}})(), device="cuda").view(3, n)
        assert torch.equal(got, torch.cumsum(y.view(3, n).long(), 1).int())


@pytest.mark.parametrize("case", ["float-1row", "float-1row-bad", "float-odd-tail-bad", "ll-to-double", "double-4rows-strided", "prod-int32"])
def test_onepass_scan(cuda_engine, case):
    """The single-pass look-back scan (scan_onepass.cuh): ONE launch for rows cut into 48 KB tiles, against torch on
    exactly representable data; the same inputs through the three-pass path must give the same bytes."""
    g = torch.Generator(device="cuda").manual_seed(41)

    def run(fn):
        c0 = cuda_engine.launch_count()
        out = fn()
        return out, cuda_engine.launch_count() - c0

    if case in ("float-1row", "float-1row-bad", "float-odd-tail-bad"):
        n = 2**26 + 4 * 777 + (3 if "odd" in case else 0)     # last tile partial; odd: 3 elements after the last vector
        y = torch.randint(-8, 9, (n,), device="cuda", generator=g).float()
        py = wrap(cuda_engine, y, T.F, [n])
        want = y.double()
        if case.endswith("bad"):
            bad = torch.rand(n, device="cuda", generator=g) < 0.01
            y[bad] = -9999.0
            py.set_badvalue(-9999.0).set_badflag(True)
            want = torch.where(bad, torch.zeros_like(want), want)
        if "odd" in case:
            y[n - 2] = -9999.0; bad[n - 2] = True; want[n - 2] = 0.0      # a BAD element inside the tail
        out, nl = run(lambda: ufunc.cumusumover(py))
        assert nl == (2 if "odd" in case else 1), nl
        want = torch.cumsum(want, 0).float()
        if case.endswith("bad"):
            want[bad] = float(out.badvalue)
        assert torch.equal(as_torch(out, torch.float32), want)
    elif case == "ll-to-double":
        n = 2**23
        y = torch.randint(-2**20, 2**20, (n,), device="cuda", generator=g)
        out, nl = run(lambda: ufunc.dcumusumover(wrap(cuda_engine, y, T.LL, [n])))
        assert nl == 1 and out.type == "double"
        assert torch.equal(as_torch(out, torch.float64), torch.cumsum(y, 0).double())
    elif case == "double-4rows-strided":
        n, pitch = 2**21 + 2, 2**21 + 64                      # rows of a wider parent: row stride != n
        y = torch.randint(-100, 101, (4 * pitch,), device="cuda", generator=g).double()
        parent = wrap(cuda_engine, y, T.D, [pitch, 4])
        out, nl = run(lambda: ufunc.cumusumover(parent.slice(f"0:{n - 1}")))
        assert nl == 1, nl
        assert torch.equal(as_torch(out, torch.float64).view(4, n), torch.cumsum(y.view(4, pitch)[:, :n], 1))
    else:
        n = 2**23
        y = torch.randint(-3, 4, (n,), device="cuda", generator=g, dtype=torch.int32)
        y[y == 0] = 1
        out, nl = run(lambda: ufunc.cumuprodover(wrap(cuda_engine, y, T.L, [n])))
        assert nl == 1, nl
        got = torch.as_tensor(type("C", (), {"__cuda_array_interface__": {"shape": (n,), "typestr": "<i4",
                              "data": (out.store.ptr, False), "version": 3}})(), device="cuda")
        assert torch.equal(got, torch.cumprod(y.long(), 0).int())          # wrap-around product, mod 2^32 either way


def test_elementwise_beyond_2_pow_32_elements(cuda_engine):
    """More than 2^32 elements in one elementwise launch (64-bit unit/item arithmetic in the walkers): sbyte
    plus over 2^32 + 37 elements, checked through slices at the start, across the 2^32 boundary and at the end,
    and the device-side sequence (axisvalues) of the same length."""
    n = 2**32 + 37
    if torch.cuda.mem_get_info()[0] < 4 * n:
        pytest.skip("needs ~17 GB of free device memory")
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(-50, 50, (n,), device="cuda", dtype=torch.int8, generator=g)
    b = torch.randint(-50, 50, (n,), device="cuda", dtype=torch.int8, generator=g)
    out = wrap(cuda_engine, a, T.SB, [n]) + wrap(cuda_engine, b, T.SB, [n])
    assert out.dims == [n]
    got = torch.as_tensor(type("C", (), {"__cuda_array_interface__": {"shape": (n,), "typestr": "|i1",
                          "data": (out.store.ptr, False), "version": 3}})(), device="cuda")
    for lo, hi in ((0, 10**6), (2**32 - 10**6, 2**32 + 37), (2**31 - 1000, 2**31 + 1000)):
        assert torch.equal(got[lo:hi], a[lo:hi] + b[lo:hi]), (lo, hi)
    del a, b, got, out
    from pdl_b200 import basic
    s = basic.sequence(T.B, n, engine=cuda_engine)          # byte: values wrap mod 256 exactly like the reference's (T)n
    sv = torch.as_tensor(type("C", (), {"__cuda_array_interface__": {"shape": (n,), "typestr": "|u1",
                         "data": (s.store.ptr, False), "version": 3}})(), device="cuda")
    for lo in (0, 2**32 - 500, n - 300):
        want = (torch.arange(lo, lo + 300, device="cuda", dtype=torch.int64) % 256).to(torch.uint8)
        assert torch.equal(sv[lo:lo + 300], want), lo


def test_scan_row_longer_than_2_pow_31(cuda_engine):
    """One row of 2^31 + 11 elements through the chunked scan (64-bit chunk arithmetic; sbyte in, long out)."""
    n = 2**31 + 11
    if torch.cuda.mem_get_info()[0] < 24 * 2**30:
        pytest.skip("needs ~24 GB of free device memory")
    g = torch.Generator(device="cuda").manual_seed(6)
    a = torch.randint(-1, 2, (n,), device="cuda", dtype=torch.int8, generator=g)
    out = ufunc.cumusumover(wrap(cuda_engine, a, T.SB, [n]))
    assert out.type == "long" and out.dims == [n]
    got = torch.as_tensor(type("C", (), {"__cuda_array_interface__": {"shape": (n,), "typestr": "<i4",
                          "data": (out.store.ptr, False), "version": 3}})(), device="cuda")
    # check in blocks of 2^28 against torch, carrying the running total
    carry, blk = 0, 2**28
    for lo in range(0, n, blk):
        hi = min(lo + blk, n)
        want = torch.cumsum(a[lo:hi].to(torch.int32), 0, dtype=torch.int32) + carry
        assert torch.equal(got[lo:hi], want), lo
        carry = int(want[-1].item())
