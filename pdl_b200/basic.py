"""PDL::Basic constructors on the device (lib/PDL/Basic.pm:117-129,479-485; PDL::Core zeroes/ones):
sequence / xvals / yvals / zvals / axisvals are `axisvalues` (lib/PDL/Primitive.pd:1468-1474) on a
view of a fresh ndarray; zeroes / ones are a broadcast assgn of a scalar.  Nothing is built on the
host and uploaded: the producers of every benchmark script stay resident."""
from __future__ import annotations

from . import types as T
from .core import PDL
from .engine import Engine, default_engine
from .trans import run_op, as_pdl


def _fill(p: PDL, value) -> PDL:
    run_op("assgn", [as_pdl(type(value)(value), p.engine).convert(p.datatype)], [p])   # $p .= value
    return p


def zeroes(datatype: int, *dims, engine: Engine | None = None) -> PDL:
    """zeroes(type, dims): the reference zero-fills in pdl_allocdata (pdlapi.c:172-209); here one write-only launch."""
    return _fill(PDL.empty(datatype, list(dims), engine or default_engine()), 0)


def ones(datatype: int, *dims, engine: Engine | None = None) -> PDL:
    return _fill(PDL.empty(datatype, list(dims), engine or default_engine()), 1)


def axisvals2(dummy: PDL, nth: int, keep_type: bool) -> PDL:
    """Basic.pm:117-124, line for line: sub-float types become float unless the caller gave a type;
    fewer dims than nth -> all zero; otherwise axisvalues in place on an xchg'd view."""
    if not keep_type and dummy.datatype < T.F:
        dummy = dummy.convert(T.F)
    if dummy.getndims() <= nth:
        return _fill(dummy, 0)
    v = dummy if nth == 0 else dummy.xchg(0, nth)
    run_op("axisvalues", [v], [v])
    return dummy


def axisvals(x, nth: int = 0) -> PDL:
    """$x->axisvals($nth): a NEW ndarray shaped like $x (Basic.pm:106-115)."""
    x = as_pdl(x)
    return axisvals2(PDL.empty(x.datatype, x.dims, x.engine), nth, False)


def xvals(x) -> PDL: return axisvals(x, 0)
def yvals(x) -> PDL: return axisvals(x, 1)
def zvals(x) -> PDL: return axisvals(x, 2)


def sequence(datatype: int | None, *dims, engine: Engine | None = None) -> PDL:
    """sequence([type,] dims) — Basic.pm:479-485: axisvals2 on the flat view; without a type: double."""
    given = datatype is not None
    p = PDL.empty(datatype if given else T.D, list(dims), engine or default_engine())
    axisvals2(p.flat(), 0, given)
    return p


__all__ = ["zeroes", "ones", "sequence", "axisvals", "xvals", "yvals", "zvals"]
