"""CPU: known-answer assertions the reference's OWN test files hold for this path, restated
one by one with their file:line and replayed through the host mirror on the C oracle.  (The same t/*.t files
run unmodified under the Perl shim on the GPU: tests/test_gpu_perl_shim.py.)  Together with the fixtures
recorded from the built reference this pins the oracle on both the reference's outputs and its own tests."""
import numpy as np
import pytest

import pdl_b200 as P
from pdl_b200 import types as T, ufunc, bad as B, basic


@pytest.fixture()
def e(oracle_engine):
    return oracle_engine


def d(e, x, t=T.D):
    return P.PDL.from_numpy(np.array(x, dtype=T.NP_DTYPE[t]), t, e)


def vals(p):
    return p.to_numpy().tolist()


def test_ufunc_t_nan_handling(e):
    nan = float("nan")
    x, y = d(e, [nan, 0, 1, 2]), d(e, [0, 1, 2, nan])
    assert ufunc.min(x).sclr() == 0 and ufunc.min(y).sclr() == 0            # t/ufunc.t:90-91
    assert ufunc.max(x).sclr() == 2 and ufunc.max(y).sclr() == 2            # t/ufunc.t:92-93
    assert ufunc.minmax(x) == (0.0, 2.0) and ufunc.minmax(y) == (0.0, 2.0)  # t/ufunc.t:94-95
    assert ufunc.minimum(y).sclr() == 0 and ufunc.minover(y).sclr() == 0    # t/ufunc.t:96-97


def test_ufunc_t_empty_and_magnover(e):
    empty = P.PDL.from_numpy(np.zeros((0,), dtype=np.float64), T.D, e)
    m = ufunc.maximum(empty)
    assert m.badflag and m.bad_mask().all()                                 # t/ufunc.t:100
    empty.badflag = True
    m = ufunc.maximum(empty)
    assert m.badflag and m.bad_mask().all()                                 # t/ufunc.t:103
    g = ufunc.magnover(empty)
    assert g.badflag and g.bad_mask().all()                                 # t/ufunc.t:104
    assert ufunc.magnover(basic.zeroes(T.D, 4, engine=e)).sclr() == 0       # t/ufunc.t:106
    assert abs(ufunc.magnover(basic.sequence(T.D, 4, engine=e)).sclr() - 3.741657) < 1e-6   # t/ufunc.t:107


def test_ufunc_t_bit_reductions(e):
    ll = lambda x: d(e, x, T.LL)
    assert ufunc.borover(ll([10, 0, -4])).sclr() == -2                      # t/ufunc.t:144
    assert ufunc.bandover(ll([-6, -1, -4])).sclr() == -8                    # t/ufunc.t:151 (~0 == -1)
    assert ufunc.borover(B.setvaltobad(ll([10, 0, -4]), 0)).sclr() == -2    # t/ufunc.t:156 (on longlong)
    assert ufunc.bxorover(ll([6, 0, -2])).sclr() == -8                      # t/ufunc.t:165
    assert ufunc.bxorover(B.setvaltobad(ll([6, 0, -2]), 0)).sclr() == -8    # t/ufunc.t:167
    assert ufunc.xorover(ll([-1, 0, 2])).sclr() == 0                        # t/ufunc.t:170
    assert ufunc.xorover(B.setvaltobad(ll([-1, 0, 2]), 0)).sclr() == 0      # t/ufunc.t:171
    assert ufunc.max(d(e, [65535], T.US)).sclr() == 65535                   # t/ufunc.t:186
    assert not ufunc.max(d(e, [65535], T.US)).badflag


def test_ufunc_t_averages_and_sums(e):
    assert ufunc.avg(P.PDL.from_numpy(np.zeros((0,), dtype=np.int32), T.L, e)).sclr() == 0          # t/ufunc.t:189
    assert np.isnan(ufunc.average(P.PDL.from_numpy(np.zeros((0,), dtype=np.float64), T.D, e)).sclr())  # t/ufunc.t:190
    X = d(e, [[5, 4, 3], [2, 3, 1.5]])
    np.testing.assert_allclose(vals(ufunc.average(X)), [4, 2.1666666], rtol=1e-6)    # t/ufunc.t:196
    assert vals(ufunc.sumover(X)) == [12, 6.5]                              # t/ufunc.t:197
    assert vals(ufunc.prodover(X)) == [60, 9]                               # t/ufunc.t:198
    assert ufunc.dsumover(basic.ones(T.B, 3000, engine=e)).sclr() == 3000   # t/ufunc.t:219


def test_ops_t_arithmetic_and_logic(e):
    pd_ = d(e, [5, 6])
    assert vals(pd_ - 1) == [4, 5] and vals(1 - pd_) == [-4, -5]            # t/ops.t:37-38
    assert vals(d(e, [0, 1, 2]) > d(e, 1.5)) == [0, 0, 1]                   # t/ops.t:67
    assert vals(d(e, [0, 1, 3], T.B) << 2) == [0, 4, 12]                    # t/ops.t:72
    assert vals(P.run_ufunc("sqrt", d(e, [16, 64, 9]))) == [4, 8, 3]        # t/ops.t:78
    assert vals(P.run_ufunc("not", d(e, [1, 0]))) == [0, 1]                 # t/ops.t:101
    assert vals(d(e, [12, 13, 14, 15, 16, 17]) % 3) == [0, 1, 2, 0, 1, 2]   # t/ops.t:103
    a, b = d(e, [1, 0, 1]), d(e, [1, 1, 0])
    r = a & b
    assert r.type == "longlong" and vals(r) == [1, 0, 0]                    # t/ops.t:120
    assert vals(a | b) == [1, 1, 1]                                         # t/ops.t:121
    np.testing.assert_allclose(vals(P.run_biop("atan2", d(e, [1, 1]), d(e, [1, 1]))), [np.arctan2(1, 1)] * 2)  # t/ops.t:126


def test_ops_t_modulus_tables(e):
    pa = np.arange(-7, 8)
    pb = np.array([[-3], [0], [3]])
    pc = np.array([[-1, 0, -2] * 5, [0] * 15, [2, 0, 1] * 5])
    for t in (T.S, T.L, T.IND, T.LL, T.F, T.D):                             # t/ops.t:212-218
        r = d(e, pa, t) % d(e, pb, t)
        assert r.type == T.NAMES[t] and vals(r) == pc.tolist(), T.NAMES[t]
    ua, ub = np.arange(15), np.array([[0], [3]])
    uc = np.array([[0] * 15, [0, 1, 2] * 5])
    for t in (T.B, T.US):                                                   # t/ops.t:227-228
        assert vals(d(e, ua, t) % d(e, ub, t)) == uc.tolist(), T.NAMES[t]
    assert vals(d(e, [255], T.B) % 1) == [0] and vals(d(e, [65535], T.US) % 1) == [0]   # t/ops.t:236-237
    ll = lambda v: d(e, [v], T.LL)
    assert (ll(10555000100001145) - ll(10555000100001144)).sclr() == 1      # t/ops.t:244
    assert (ll(9223372036854775807) - ll(9223372036854775806)).sclr() == 1  # t/ops.t:248
    assert (ll(9223372036854775807) + ll(-9223372036854775808)).sclr() == -1  # t/ops.t:249


def test_primitive_matmult_t_fiducials(e):
    IM = [[1, 2, 3, 3, 5], [2, 3, 4, 5, 6], [13, 13, 13, 13, 13], [1, 3, 1, 3, 1], [10, 10, 2, 2, 2]]
    want = [[97, 106, 63, 71, 69], [125, 140, 87, 100, 97], [351, 403, 299, 338, 351], [33, 43, 33, 42, 41],
            [78, 102, 102, 116, 142]]
    assert vals(P.matmult(d(e, IM), d(e, IM))) == want                      # t/primitive-matmult.t:21-27
    PA, PB = [[1, 2, 3, 0], [1, -1, 2, 7], [1, 0, 0, 1]], [[1, 1], [0, 2], [0, 2], [1, 1]]
    PC = [[1, 11], [8, 10], [2, 2]]
    assert vals(P.matmult(d(e, PA), d(e, PB))) == PC                        # t/primitive-matmult.t:47
    assert vals(P.matmult(d(e, [[1, 1, 1, 1]], T.F), d(e, PB))) == [[2, 6]]  # t/primitive-matmult.t:72
    with pytest.raises(P.PDLError, match="mismatch in matmult"):            # t/primitive-matmult.t:75-79
        P.matmult(d(e, PB), d(e, [[1, 1, 1, 1]], T.F))
    assert vals(P.matmult(d(e, PB), d(e, 2.0))) == (np.array(PB) * 2).tolist()   # t/primitive-matmult.t:81
    nan = float("nan")
    C = P.matmult(d(e, [[1, nan, 0], [0, 1, 0], [0, 0, 1]]), basic.sequence(T.D, 2, 3, engine=e))
    B.setnantobad(C.inplace())
    B.setbadtoval(C.inplace(), 6)
    assert vals(C) == [[6, 6], [2, 3], [4, 5]]                              # t/primitive-matmult.t:84-91
    A = d(e, [[1, -9, 0], [0, 1, 0], [0, 0, 1]]).set_badvalue(-9.0).set_badflag(True)
    C = P.matmult(A, basic.sequence(T.D, 2, 3, engine=e))
    B.setbadtoval(C.inplace(), 6)
    assert vals(C) == [[6, 6], [2, 3], [4, 5]]                              # t/primitive-matmult.t:93-99


def bad_at(p):
    """(values with BAD as None) for comparing against the reference's 'BAD' literals."""
    a, m = p.to_numpy().reshape(-1).tolist(), p.bad_mask().reshape(-1).tolist() if p.badflag else [False] * p.nelem
    return [None if b else v for v, b in zip(a, m)]


def test_bad_t_propagation_and_queries(e):
    BAD = 255
    x = d(e, [1, 2, 3], T.B).set_badflag(True)
    y = d(e, [1, BAD, 3], T.B).set_badflag(True)
    assert bad_at(x + y) == [2, None, 6]                                    # t/bad.t:111
    c = y.convert(T.F)
    assert c.type == "float" and bad_at(c) == [1.0, None, 3.0]              # t/bad.t:115
    assert ufunc.sum(c).sclr() == 4 and ufunc.sum(c).type == "float"        # t/bad.t:116
    x = d(e, [1, 2, BAD, BAD, 5, 6, BAD, 8, 9], T.B).set_badflag(True)
    assert B.isbad(x).type == "long" and vals(B.isbad(x)) == [0, 0, 1, 1, 0, 0, 1, 0, 0]    # t/bad.t:119
    assert vals(B.isgood(x)) == [1, 1, 0, 0, 1, 1, 0, 1, 1]                 # t/bad.t:120
    assert ufunc.nbadover(x.flat()).sclr() == 3 and ufunc.ngoodover(x.flat()).sclr() == 6   # t/bad.t:121-122
    nan = float("nan")
    assert vals(B.isnan(d(e, [1, 2, nan, nan, 5, 6, nan, 8, 9]))) == [0, 0, 1, 1, 0, 0, 1, 0, 0]   # t/bad.t:125
    x = d(e, [[BAD, BAD], [BAD, 0], [0, 0]], T.B).set_badflag(True)
    assert ufunc.nbadover(x).type == "indx" and vals(ufunc.nbadover(x)) == [2, 1, 0]        # t/bad.t:128
    assert vals(ufunc.ngoodover(x)) == [0, 1, 2]                            # t/bad.t:129
    assert bad_at(d(e, [1, 2, BAD, 4], T.B).set_badflag(True) << 2) == [4, 8, None, 16]     # t/bad.t:250


def test_bad_t_setbad_family(e):
    data = [42, 47, 98, 13, 22, 96, 74, 41, 79, 76, 96, 3, 32, 76, 25, 59, 5, 96, 32, 6]
    want = [42, 47, 98, 20, 22, 96, 74, 41, 79, 76, 96, 20, 32, 76, 25, 59, 20, 96, 32, 20]
    mask = [0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 1]
    x = d(e, data)
    y = B.setbadif(x, x < 20)
    assert y.badflag
    assert vals(B.setbadtoval(y, 20)) == want and y.badflag                 # t/bad.t:175-177
    B.setbadtoval(y.inplace(), 20)
    assert vals(y) == want and not y.badflag                                # t/bad.t:180-183
    y = B.setbadif(x, x < 20)
    assert vals(B.isbad(B.copybad(x, y))) == mask                           # t/bad.t:186-190
    x2 = d(e, data)
    B.copybad(x2.inplace(), y)
    assert vals(B.isbad(x2)) == mask                                        # t/bad.t:192-196
    nan, inf = float("nan"), float("inf")
    x = d(e, [0, 1, -9, 3, 4]).set_badvalue(-9.0).set_badflag(True)
    B.badmask(x.inplace(), 0)
    assert vals(x) == [0, 1, 0, 3, 4] and not x.badflag                     # t/bad.t:314-316
    x = basic.sequence(T.D, 10, engine=e) % 4
    B.setvaltobad(x.inplace(), 1)
    assert bad_at(x) == [0, None, 2, 3, 0, None, 2, 3, 0, None]             # t/bad.t:319-321
    B.setbadtonan(x.inplace())
    got = vals(x)
    assert [v != v for v in got] == [False, True, False, False] * 2 + [False, True] and not x.badflag   # t/bad.t:323-324
    assert bad_at(B.setvaltobad(d(e, [1, 2, 3, 4], T.F), 2)) == [1, None, 3, 4]             # t/bad.t:327
    assert bad_at(B.setvaltobad(d(e, [1, 2, 3, 4], T.D), 2)) == [1, None, 3, 4]             # t/bad.t:328
    i2b = d(e, [0, inf, nan])
    B.setinftobad(i2b.inplace())
    r = bad_at(i2b)
    assert r[0] == 0 and r[1] is None and r[2] != r[2]                      # t/bad.t:330-332
    xc = d(e, [0, inf, 2, 3, 0, nan, 2, 3, 0, nan])
    B.setnonfinitetobad(xc.inplace())
    assert bad_at(xc) == [0, None, 2, 3, 0, None, 2, 3, 0, None]            # t/bad.t:334-336
    B.setnantobad(x.inplace())
    assert bad_at(x) == [0, None, 2, 3, 0, None, 2, 3, 0, None]             # t/bad.t:343
