"""PDL::Bad surface on the device path (lib/PDL/Bad.pd:343-416,584-905): the bad-value producers
and consumers that sit either side of the hot path.  Each is ONE elementwise
launch through pdlb200_readdata; the output's badflag follows the reference's $PDLSTATESET* rules
(pdl_b200/trans.py)."""
from __future__ import annotations

from .trans import run_op, as_pdl


def _unary(name):
    def f(a, b=None):
        a = as_pdl(a)
        if b is None and a.is_inplace():
            a._inplace = False
            b = a
        return run_op(name, [a], [b])[0]
    f.__name__ = name
    f.__doc__ = f"PDL::{name}(a(); [o]b()) — lib/PDL/Bad.pd"
    return f


isbad = _unary("isbad")
isgood = _unary("isgood")
isnan = _unary("isnan")
setnantobad = _unary("setnantobad")
setinftobad = _unary("setinftobad")
setnonfinitetobad = _unary("setnonfinitetobad")
setbadtonan = _unary("setbadtonan")


def _with_value(name):
    def f(a, value, b=None):
        a = as_pdl(a)
        if b is None and a.is_inplace():
            a._inplace = False
            b = a
        return run_op(name, [a], [b], param=float(value))[0]
    f.__name__ = name
    f.__doc__ = f"PDL::{name}(a(); [o]b(); double value) — lib/PDL/Bad.pd"
    return f


setvaltobad = _with_value("setvaltobad")
setbadtoval = _with_value("setbadtoval")


def setbadif(a, mask, b=None):
    """PDL::setbadif(a(); int mask(); [o]b()) — Bad.pd:584-637.  Not available in place."""
    a = as_pdl(a)
    return run_op("setbadif", [a, as_pdl(mask, a.engine)], [b])[0]


def _binary(name):
    def f(a, other, c=None):
        a = as_pdl(a)
        if c is None and a.is_inplace():
            a._inplace = False
            c = a
        return run_op(name, [a, as_pdl(other, a.engine)], [c])[0]
    f.__name__ = name
    f.__doc__ = f"PDL::{name} — lib/PDL/Bad.pd:842-905"
    return f


badmask = _binary("badmask")
copybad = _binary("copybad")

__all__ = ["isbad", "isgood", "isnan", "setbadif", "setvaltobad", "setnantobad", "setinftobad",
           "setnonfinitetobad", "setbadtonan", "setbadtoval", "badmask", "copybad"]
