"""PDL::Ops function surface (lib/PDL/Ops.pd): same names, argument order and `swap`
meaning as the reference's XS entry points `PDL::plus($a,$b,[$c],$swap)` etc."""
from __future__ import annotations

from .trans import run_biop, run_ufunc, run_op, as_pdl

_BINARY = ["plus", "mult", "minus", "divide", "gt", "lt", "le", "ge", "eq", "ne",
           "shiftleft", "shiftright", "or2", "and2", "xor", "power", "atan2", "modulo", "spaceship"]
_UNARY = ["bitnot", "sqrt", "sin", "cos", "exp", "log", "log10", "abs2"]


def _mk_bin(name):
    def f(a, b, c=None, swap=0):
        return run_biop(name, a, b, c, swap)
    f.__name__ = name
    f.__doc__ = f"PDL::{name}(a, b, [c], swap) — lib/PDL/Ops.pd biop/bifunc"
    return f


def _mk_un(name):
    def f(a, b=None):
        return run_ufunc(name, a, b)
    f.__name__ = name
    f.__doc__ = f"PDL::{name}(a, [b]) — lib/PDL/Ops.pd ufunc"
    return f


for _n in _BINARY:
    globals()[_n] = _mk_bin(_n)
for _n in _UNARY:
    globals()[_n] = _mk_un(_n)

xor2 = globals()["xor"]          # lib/PDL/Ops.pd:315-318
not_ = _mk_un("not")             # `not` is a Python keyword
abs_ = _mk_un("_rabs")           # PDL::abs -> _rabs for real types (Ops.pd:488)
_rabs = abs_


def assgn(a, b):
    """PDL::assgn(a, b): b .= a (lib/PDL/Ops.pd:382-397)."""
    return run_op("assgn", [as_pdl(a, getattr(b, "engine", None))], [b])[0]


def ipow(a, b, ans=None):
    """PDL::ipow(a, b, [ans]): a ** b for integer b by squaring (lib/PDL/Ops.pd:443-476)."""
    a = as_pdl(a)
    if ans is None and a.is_inplace():
        a._inplace = False
        ans = a
    return run_op("ipow", [a, as_pdl(b, a.engine)], [ans])[0]


__all__ = ["ipow"] + _BINARY + _UNARY + ["xor2", "not_", "abs_", "_rabs", "assgn"]
