"""CPU: the Python mirror of the reference's host logic, checked against facts recorded from the
real reference (SURVEY.md Appendix B and the cited source lines) with the oracle engine as backend."""
import numpy as np
import pytest

import pdl_b200 as P
from pdl_b200 import types as T, ufunc, ops, bad
from pdl_b200.trans import transtype_select, SPECS, par_type


def test_scalar_typing_matches_pdlperl_h():
    # lib/PDL/Core/pdlperl.h:185-205: IV -> smallest SIGNED type, NV -> double
    assert T.scalar_type(1) == T.SB and T.scalar_type(300) == T.S and T.scalar_type(100000) == T.L
    assert T.scalar_type(2**40) == T.IND and T.scalar_type(-2**40) == T.IND
    assert T.scalar_type(2**63 + 5) == T.ULL
    assert T.scalar_type(1.5) == T.D and T.scalar_type(float("nan")) == T.D


def test_type_selection_rules(oracle_engine):
    e = oracle_engine
    f = P.PDL.from_numpy(np.array([1, 2], dtype=np.float32), T.F, e)
    b = P.PDL.from_numpy(np.array([1, 2], dtype=np.uint8), T.B, e)
    lng = P.PDL.from_numpy(np.array([2, 3], dtype=np.int32), T.L, e)
    assert (f + 1.5).type == "double" and (f + 1).type == "float"          # Appendix B
    assert (b + 1).type == "byte" and (b + 300).type == "short"
    assert (lng ** 2).type == "double"                                      # power: GenericTypes [C.., F, LD, D]
    assert (f << 1).type == "longlong"                                      # bit ops: last integer type
    assert ufunc.sumover(b).type == "long" and ufunc.average(b).type == "long" and ufunc.minimum(b).type == "byte"
    assert ufunc.dsumover(b).type == "double" and ufunc.maximum_ind(f).type == "indx"
    # an existing output decides the type first (pdlapi.c:1216-1231)
    out = P.PDL.empty(T.D, [2], e)
    assert transtype_select(SPECS["plus"], [f, f, out]) == T.D
    assert par_type(SPECS["sumover"].pars[1], T.B) == T.L and par_type(SPECS["sumover"].pars[1], T.D) == T.D


def test_views_are_metadata_only(oracle_engine):
    e = oracle_engine
    a = P.sequence(T.D, 10, 10, engine=e)
    s = a.slice("1:8:2,(3)")
    assert s.dims == [4] and s.dimincs == [2] and s.offs == 31 and s.store is a.store
    assert a.xchg(0, 1).dimincs == [10, 1]
    assert P.sequence(T.D, 5, engine=e).dummy(0, 3).dimincs == [0, 1]       # Appendix B: dimincs 0,1
    assert a.slice("-1:0:-3,:").to_numpy()[0].tolist() == [9, 6, 3, 0]
    assert a.slice(":,*2,(0)").dims == [10, 2]
    assert a.clump(-1).dims == [100] and a.flat().store is a.store           # physical parent: view
    assert a.xchg(0, 1).flat().store is not a.store                          # strided parent: copied on device
    with pytest.raises(P.PDLError):
        a.slice("11,:")


def test_broadcast_errors_and_empty(oracle_engine):
    e = oracle_engine
    with pytest.raises(P.PDLError, match="Mismatched implicit broadcast dimension 0: size 3 vs. 4"):
        P.sequence(T.D, 3, engine=e) + P.sequence(T.D, 4, engine=e)
    z = P.PDL.from_numpy(np.zeros((3, 0)), T.D, e)                           # dims [0,3]
    assert (z + 1).dims == [0, 3]
    assert ufunc.sumover(z).to_numpy().tolist() == [0, 0, 0]
    m = ufunc.maximum(z)
    assert m.badflag and np.all(m.to_numpy() == T.DEFAULT_BAD[T.D])          # Ufunc.pd:463-464, t/ufunc.t:99-104
    with pytest.raises(P.PDLError, match=r"Dim mismatch in matmult of \[3x2\] x \[2x2\]: 3 != 2"):
        P.matmult(P.sequence(T.D, 3, 2, engine=e), P.sequence(T.D, 2, 2, engine=e))


def test_inplace_and_outputs(oracle_engine):
    e = oracle_engine
    a = P.sequence(T.L, 6, engine=e)
    v = a.slice("1:-1:2")
    v += 10                                                                   # writes through to the parent
    assert a.to_numpy().tolist() == [0, 11, 2, 13, 4, 15]
    r = a.inplace().plus(1) if hasattr(a, "plus") else ops.plus(a.inplace(), 1)
    assert r is a and a.to_numpy().tolist() == [1, 12, 3, 14, 5, 16] and not a.is_inplace()
    out = P.PDL.empty(T.L, [6], e)
    assert ops.minus(a, 1, out, 1) is out and out.to_numpy().tolist() == [0, -11, -2, -13, -4, -15]  # swap: 1 - a
    with pytest.raises(P.PDLError, match="can't broadcast over output"):
        P.run_op("plus", [P.sequence(T.L, 6, 2, engine=e), a], [P.PDL.empty(T.L, [6], e)])


def test_prepared_op_is_one_call(oracle_engine):
    e = oracle_engine
    a, b = P.sequence(T.F, 5, 4, engine=e), P.sequence(T.F, 5, engine=e)
    out = P.PDL.empty(T.F, [5, 4], e)
    prep = P.prepare_op("mult", [a, b], [out])
    n0 = e.calls
    prep(); prep()
    assert e.calls == n0 + 2
    assert np.array_equal(out.to_numpy(), a.to_numpy() * b.to_numpy())
    with pytest.raises(P.PDLError, match="prepare_op needs inputs already"):
        P.prepare_op("plus", [a, P.sequence(T.D, 5, engine=e)], [None])


def test_badflag_propagation(oracle_engine):
    e = oracle_engine
    a = P.PDL.from_numpy(np.array([1, 2, 3], dtype=np.int32), T.L, e)
    b = P.PDL.from_numpy(np.array([1, T.DEFAULT_BAD[T.L], 3], dtype=np.int32), T.L, e).set_badflag(True)
    c = a + b
    assert c.badflag and c.bad_mask().tolist() == [False, True, False]       # pdlapi.c:806-808, t/bad.t:27-46
    assert not (a + a).badflag
    # the OTHER operand holds the badvalue bit pattern but has no badflag: biop tests the state flag (Ops.pd:144)
    d = P.PDL.from_numpy(np.array([T.DEFAULT_BAD[T.L], 5, 6], dtype=np.int32), T.L, e)
    r = d + b
    assert r.to_numpy()[0] == np.int32(T.DEFAULT_BAD[T.L]) + np.int32(1)      # computed (wraps), not forced BAD


def test_ops_are_also_methods(oracle_engine):
    """`$x->sumover`, `$x->setbadif($m)->isbad`, `$x->inner($y)`, `$x->minmax`, `$x->yvals`: the op surface as methods."""
    e = oracle_engine
    x = P.sequence(T.F, 5, 3, engine=e)
    assert x.sumover().to_numpy().tolist() == [10.0, 35.0, 60.0]
    assert x.minmax() == (0.0, 14.0)
    m = P.PDL.from_numpy(np.array([0, 1, 0, 0, 1], dtype=np.int32), T.L, e)
    assert x.setbadif(m).isbad().sumover().to_numpy().tolist() == [2, 2, 2]
    assert x.inner(x).to_numpy().tolist() == [30.0, 255.0, 730.0]
    assert x.yvals().to_numpy()[2].tolist() == [2.0] * 5
    assert x.setbadif(m).setbadtoval(-1).minimum().to_numpy().tolist() == [-1.0, -1.0, -1.0]


def test_prepared_ops_apply_output_state_rules(oracle_engine):
    """prepare_op + call must leave the same badflags as run_op: data-dependent for setnantobad / minmaximum,
    unconditional for setbadif / setbadtoval."""
    e = oracle_engine
    x = P.PDL.from_numpy(np.array([1.0, np.nan, 3.0], dtype=np.float32), T.F, e)
    out = P.PDL.empty(T.F, [3], e)
    P.prepare_op("setnantobad", [x], [out])()
    assert out.badflag and out.bad_mask().tolist() == [False, True, False]
    clean = P.PDL.from_numpy(np.array([1.0, 2.0, 3.0], dtype=np.float32), T.F, e)
    out2 = P.PDL.empty(T.F, [3], e)
    P.prepare_op("setnantobad", [clean], [out2])()
    assert not out2.badflag
    allnan = P.PDL.from_numpy(np.full((2, 3), np.nan, dtype=np.float32), T.F, e)
    outs = [P.PDL.empty(T.F, [2], e), P.PDL.empty(T.F, [2], e), P.PDL.empty(T.IND, [2], e), P.PDL.empty(T.IND, [2], e)]
    P.prepare_op("minmaximum", [allnan], outs)()
    assert all(o.badflag for o in outs)
    m = P.PDL.from_numpy(np.array([0, 1, 0], dtype=np.int32), T.L, e)
    out3 = P.PDL.empty(T.F, [3], e)
    P.prepare_op("setbadif", [clean, m], [out3])()
    assert out3.badflag
    b = clean.copy().set_badflag(True)
    out4 = P.PDL.empty(T.F, [3], e)
    P.prepare_op("setbadtoval", [b], [out4])()
    assert not out4.badflag


def test_flowing_defers_readdata_and_fuses_into_sumover(oracle_engine):
    """`$a->flowing * $b` defers the product's readdata (pdlapi.c:781-801); `->sumover` as its consumer runs the
    pair as ONE transformation (inner); anything else that needs the data runs the deferred readdata first."""
    e = oracle_engine
    rng = np.random.default_rng(8)
    a = P.PDL.from_numpy(rng.integers(-1024, 1024, size=64) / 256, T.D, e).slice("0:-1:2").dummy(1, 1)
    b = P.PDL.from_numpy(rng.integers(-1024, 1024, size=40) / 256, T.D, e).slice("0:-1:2").dummy(0, 1)
    c0 = e.calls
    unfused = ufunc.sumover(a * b)
    assert e.calls - c0 == 2
    c0 = e.calls
    fused = ufunc.sumover(a.flowing() * b)
    assert e.calls - c0 == 1                                   # one launch, no intermediate
    assert fused.dims == unfused.dims and fused.type == unfused.type
    assert fused.to_numpy().tobytes() == unfused.to_numpy().tobytes()
    # the deferred product is an ordinary ndarray as soon as anyone else looks at it
    prod = a.flowing() * b
    assert prod.is_pending() and prod.dims == [32, 20]
    c0 = e.calls
    assert prod.to_numpy().tobytes() == (a * b).to_numpy().tobytes()
    assert not prod.is_pending()
    # not fused where the fused form would change the answer: BAD values (sumover skips, inner poisons), small ints
    ab = P.PDL.from_numpy(np.array([1, 2, 3, -3.4028234663852886e38], dtype=np.float32), T.F, e).set_badflag(True)
    r = ufunc.sumover(ab.flowing() * ab)
    assert r.badflag and r.sclr() == 14.0
    ai = P.PDL.from_numpy(np.array([100, 100, 100], dtype=np.int16), T.S, e)
    r = ufunc.sumover(ai.flowing() * ai)
    assert r.type == "long" and r.sclr() == 30000
    # the flag applies to the NEXT operation only
    f = a.flowing()
    assert (f * b).is_pending() and not (a * b).is_pending()


def test_badflag_is_shared_by_views_and_parent(oracle_engine):
    """A view and its parent are the same data: flagging through one is seen through the other (the reference
    propagates the flag through the vaffine family, pdlapi.c:28-36; round-1 advisor finding)."""
    e = oracle_engine
    x = P.PDL.from_numpy(np.arange(12, dtype=np.float64).reshape(3, 4), T.D, e)
    v = x.slice("0:2,(1)")
    assert not x.badflag and not v.badflag
    bad.setvaltobad(v.inplace(), 5.0)
    assert v.badflag and x.badflag                       # the parent sees it
    s = ufunc.sumover(x)                                 # ... and BAD-mode reductions run: 4+6+7 (5 is BAD)
    assert s.badflag and s.to_numpy().tolist() == [6.0, 17.0, 38.0]
    x.badflag = False
    assert not v.badflag
    # scalars come from a cache of device ndarrays: flagging one use must not flag the next
    a = P.as_pdl(7, e); a.badflag = True
    assert not P.as_pdl(7, e).badflag


def test_deferred_output_badflag_ring(oracle_engine, monkeypatch):
    """FlagRing (engine.py): ops with a data-dependent output badflag (minmaximum, set*tobad) hand the C call a
    pinned slot instead of synchronising; the output Stores settle it at the first question about their bad state.
    Host logic only — the oracle engine stands in for the device and writes the slot synchronously."""
    import ctypes as C
    from pdl_b200 import engine as E

    class _Lib:
        def pdlb200_host_alloc(self, n):
            self.buf = C.create_string_buffer(n)
            return C.addressof(self.buf)

    class _Dev:
        lib, syncs = _Lib(), 0

        def sync(self):
            _Dev.syncs += 1

    ring = E.FlagRing(_Dev())
    monkeypatch.setattr(type(oracle_engine), "flag_ring", property(lambda self: ring), raising=False)
    a = np.random.default_rng(3).random((4, 50)).astype(np.float32)
    a[2, :] = np.nan                                         # a row without a usable element: outputs flagged BAD
    ga = P.PDL.from_numpy(a, T.F, oracle_engine)
    outs = ufunc.minmaximum(ga)
    assert [o.store._pend for o in outs] == [0, 0, 0, 0] and _Dev.syncs == 0
    outs2 = ufunc.minmaximum(P.PDL.from_numpy(np.nan_to_num(a, nan=1.0), T.F, oracle_engine))
    assert outs2[0].store._pend == 1 and _Dev.syncs == 0
    assert outs[3].slice("0:1").badflag and _Dev.syncs == 1   # one synchronise settles every flag in flight
    assert all(o.store._pend is None for o in outs + outs2) and all(o.badflag for o in outs)
    assert not any(o.badflag for o in outs2) and _Dev.syncs == 1
    outs = ufunc.minmaximum(ga)
    outs[0].badflag = False                                   # an explicit setting wins over the flag in flight
    assert outs[0].store._pend is None and not outs[0].badflag and outs[1].badflag
    b = bad.setnantobad(ga)
    assert b.store._pend is not None and b.badflag
    # the ring wraps: a slot still in flight is settled before it is handed out again
    keep = [ufunc.minmaximum(ga) for _ in range(E.FlagRing.SLOTS + 3)]
    assert all(o.badflag for o in keep[0]) and all(o.badflag for o in keep[-1])
