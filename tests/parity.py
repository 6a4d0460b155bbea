"""Seeded parity cases: the same call on the same inputs through the CUDA engine and the C
oracle, compared byte for byte (or within a stated ulp tolerance for transcendental ops and
order-dependent float sums)."""
from __future__ import annotations

import numpy as np

import pdl_b200 as P
from pdl_b200 import types as T
from replay import ulp_diff

ALL_TYPES = [T.SB, T.B, T.S, T.US, T.L, T.UL, T.IND, T.ULL, T.LL, T.F, T.D]
INT_TYPES = [t for t in ALL_TYPES if t in T.INTEGER]


def rand_array(rng, t, shape, flavour="mixed"):
    dt = T.NP_DTYPE[t]
    n = int(np.prod(shape)) if len(shape) else 1
    if t in T.INTEGER:
        info = np.iinfo(dt)
        if flavour == "small":
            lo, hi = (0, 9) if t in T.UNSIGNED else (-9, 9)
        elif flavour == "pos":
            lo, hi = 1, 9
        else:
            lo, hi = max(info.min, -2**31), min(info.max, 2**31 - 1)
        a = rng.integers(lo, hi, size=n, endpoint=True).astype(dt)
    else:
        if flavour == "small":
            a = (rng.integers(-9, 9, size=n, endpoint=True) / 2).astype(dt)
        elif flavour == "pos":
            a = (rng.integers(1, 1000, size=n, endpoint=True) / 8).astype(dt)
        elif flavour == "exact":   # sums of these are exact in any order
            a = rng.integers(-8, 8, size=n, endpoint=True).astype(dt)
        else:
            a = (rng.integers(-1000000, 1000000, size=n, endpoint=True) / 1024).astype(dt)
    return a.reshape(shape)


def both(engines, arr, t, badflag=False, badvalue=None):
    out = []
    for e in engines:
        p = P.PDL.from_numpy(arr, t, e)
        p.badflag = badflag
        if badvalue is not None:
            p.set_badvalue(badvalue)
        out.append(p)
    return out


def assert_same(name, got: P.PDL, want: P.PDL, tol_ulp=None, nan_equal=False):
    """nan_equal: NaN results must be NaN on both sides but may differ in sign/payload (float
    sums and products: the bits of a NaN depend on summation order and on x86-vs-GPU NaN
    generation, not on PDL semantics); everything else stays bit-exact."""
    assert got.type == want.type, (name, got.type, want.type)
    assert got.dims == want.dims, (name, got.dims, want.dims)
    assert got.badflag == want.badflag, (name, "badflag")
    g, w = got.to_numpy(), want.to_numpy()
    if nan_equal and g.dtype.kind == "f" and not tol_ulp:
        assert ulp_diff(g, w) == 0, (name, "differs beyond NaN payload")
        return
    if tol_ulp and g.dtype.kind == "f":
        d = ulp_diff(g, w)
        assert d <= tol_ulp, (name, f"{d} ulp > {tol_ulp}")
    else:
        if g.tobytes() != w.tobytes():
            bad = np.flatnonzero(g.reshape(-1).view(np.uint8 if g.itemsize == 1 else f"u{g.itemsize}") !=
                                 w.reshape(-1).view(np.uint8 if w.itemsize == 1 else f"u{w.itemsize}"))
            i = int(bad[0])
            raise AssertionError(f"{name}: {bad.size} of {g.size} elements differ; first at flat {i}: "
                                 f"got {g.reshape(-1)[i]!r} want {w.reshape(-1)[i]!r}")
