"""PDL::Primitive::matmult on the device path (lib/PDL/Primitive.pd:191-264)."""
from __future__ import annotations

from .core import PDL
from .engine import PDLError
from .trans import run_op, as_pdl


def matmult(x, y, c=None) -> PDL:
    """The Perl wrapper PDL::matmult restated (Primitive.pd:197-209): promote to >= 2 dims,
    scalar shortcut through `*`, dim check with the reference's message, then _matmult_int."""
    x = as_pdl(x)
    y = as_pdl(y, x.engine)
    while x.getndims() < 2:
        x = x.dummy(-1)
    while y.getndims() < 2:
        y = y.dummy(-1)
    if (x.dim(0) == 1 and x.dim(1) == 1) or (y.dim(0) == 1 and y.dim(1) == 1):
        r = x * y
        if c is not None and not c.isnull():
            c.assign(r)
            return c
        return r
    if y.dim(1) != x.dim(0):
        raise PDLError("Dim mismatch in matmult of [%dx%d] x [%dx%d]: %d != %d" %
                       (x.dim(0), x.dim(1), y.dim(0), y.dim(1), x.dim(0), y.dim(1)))
    return run_op("matmult", [x, y], [c])[0]


def inner(a, b, c=None) -> PDL:
    """PDL::inner(a(n); b(n); [o]c()) — lib/PDL/Primitive.pd:48-70: c = sum_n a*b in ONE launch; the fused
    form of `($a * $b)->sumover` (the product ndarray never exists)."""
    a = as_pdl(a)
    return run_op("inner", [a, as_pdl(b, a.engine)], [c])[0]


def outer(a, b, c=None) -> PDL:
    """PDL::outer(a(n); b(m); [o]c(n,m)) — lib/PDL/Primitive.pd:78-96."""
    a = as_pdl(a)
    return run_op("outer", [a, as_pdl(b, a.engine)], [c])[0]


__all__ = ["matmult", "inner", "outer"]
