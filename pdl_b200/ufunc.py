"""PDL::Ufunc surface on the device path (lib/PDL/Ufunc.pd): the `Xover` reductions along
dim 0 and the whole-array wrappers `sum max ...` = `$x->flat->Xover` (Ufunc.pd:618-663)."""
from __future__ import annotations

from .trans import run_op, as_pdl, fused_reduction

_REDUCERS = ["sumover", "prodover", "dsumover", "dprodover", "average", "daverage",
             "minimum", "maximum", "minimum_ind", "maximum_ind",
             "andover", "orover", "zcover", "xorover", "bandover", "borover", "bxorover",
             "nbadover", "ngoodover", "cumusumover", "cumuprodover", "dcumusumover", "dcumuprodover", "magnover"]


def _mk(name):
    def f(a, b=None):
        a = as_pdl(a)
        if b is None and a._pending is not None:      # deferred producer (->flowing): try the fused form first
            r = fused_reduction(name, a)
            if r is not None:
                return r
        return run_op(name, [a], [b])[0]
    f.__name__ = name
    f.__doc__ = f"PDL::{name}(a(n); [o]b()) — lib/PDL/Ufunc.pd"
    return f


for _n in _REDUCERS:
    globals()[_n] = _mk(_n)

# synonyms (Ufunc.pd:431,470)
avgover = globals()["average"]
davgover = globals()["daverage"]
minover = globals()["minimum"]
maxover = globals()["maximum"]
minover_ind = globals()["minimum_ind"]
maxover_ind = globals()["maximum_ind"]

_WHOLE = {"sum": "sumover", "prod": "prodover", "avg": "average", "dsum": "dsumover",
          "dprod": "dprodover", "davg": "daverage", "min": "minimum", "max": "maximum",
          "zcheck": "zcover", "and_": "andover", "or_": "orover", "band": "bandover",
          "bor": "borover", "xorall": "xorover", "bxor": "bxorover"}


def _mk_whole(name, over):
    def f(a):
        return globals()[over](as_pdl(a).flat())
    f.__name__ = name
    f.__doc__ = f"PDL::{name.rstrip('_')}: $x->flat->{over} (lib/PDL/Ufunc.pd:660)"
    return f


for _n, _o in _WHOLE.items():
    globals()[_n] = _mk_whole(_n, _o)

def minmaximum(a, cmin=None, cmax=None, cmin_ind=None, cmax_ind=None):
    """PDL::minmaximum(a(n); [o]cmin(); [o]cmax(); indx [o]cmin_ind(); indx [o]cmax_ind()) — Ufunc.pd:563-613."""
    return tuple(run_op("minmaximum", [as_pdl(a)], [cmin, cmax, cmin_ind, cmax_ind]))


minmaxover = minmaximum


def _mk_n_ind(name):
    def f(a, c=None, m_size=None):
        """PDL::%s(a(n); indx [o]c(m); m_size) — lib/PDL/Ufunc.pd:502-561: `c` may be the size (as in the
        reference's Perl wrapper), an existing output, or None with m_size."""
        a = as_pdl(a)
        if c is not None and not hasattr(c, "dims"):
            c, m_size = None, int(c)
        if m_size is None:
            m_size = c.dims[0]
        return run_op(name, [a], [c], sizes={"m": int(m_size)})[0]
    f.__name__ = name
    return f


minimum_n_ind = _mk_n_ind("minimum_n_ind")
maximum_n_ind = _mk_n_ind("maximum_n_ind")
min_n_ind, max_n_ind = minimum_n_ind, maximum_n_ind


def minmax(a):
    """PDL::minmax (Ufunc.pd:738): map $_->sclr, ($x->flat->minmaximum)[0,1]"""
    r = minmaximum(as_pdl(a).flat())
    return r[0].sclr(), r[1].sclr()


__all__ = ["minmaximum", "minmaxover", "minmax", "minimum_n_ind", "maximum_n_ind"] + _REDUCERS + list(_WHOLE) + ["avgover", "davgover", "minover", "maxover", "minover_ind", "maxover_ind"]
