"""pdl_b200 — B200-native implementation of PDL's broadcast-loop hot path
(PDL::Ops elementwise ops, PDL::Ufunc reductions, PDL::Primitive::matmult) behind the
reference's operator surface.  The compute lives in libpdlb200.so (hand-written
sm_100a CUDA behind the C-ABI in include/pdlb200.h); this package is the host-side
mirror of the reference's interface for that path.  There is no CPU fallback."""
from . import types
from .types import SB, B, S, US, L, UL, IND, ULL, LL, F, D
from .engine import CudaEngine, Engine, PDLError, default_engine, set_default_engine
from .core import PDL, pdl, null
from .trans import run_op, prepare_op, Prepared, run_biop, run_ufunc, as_pdl, convert_type, SPECS
from . import ops, ufunc, primitive, bad, basic
from .basic import zeroes, ones, sequence, xvals, yvals, zvals, axisvals   # device-side constructors (no host build + upload)
from .primitive import matmult, inner, outer

__all__ = ["PDL", "pdl", "zeroes", "ones", "sequence", "null", "PDLError", "CudaEngine", "Engine",
           "default_engine", "set_default_engine", "run_op", "prepare_op", "Prepared", "run_biop", "run_ufunc", "as_pdl",
           "convert_type", "SPECS", "ops", "ufunc", "primitive", "bad", "basic", "matmult", "inner", "outer", "types",
           "xvals", "yvals", "zvals", "axisvals", "SB", "B", "S", "US", "L", "UL", "IND", "ULL", "LL", "F", "D"]


def _attach_methods() -> None:
    """In PDL every op is also a method (`$x->sumover`, `$x->setbadif($m)`, `$x->xvals`, `$x->inner($y)`):
    give the mirror the same spelling.  Names that the class already defines (views, copy, ...) are kept."""
    table = {}
    for mod, names in ((ufunc, ufunc.__all__), (bad, bad.__all__), (primitive, primitive.__all__),
                       (basic, ("xvals", "yvals", "zvals", "axisvals")), (ops, getattr(ops, "__all__", ()))):
        for n in names:
            f = getattr(mod, n, None)
            if callable(f):
                table[n] = f
    for n, f in table.items():
        if not hasattr(PDL, n):
            setattr(PDL, n, (lambda fn: lambda self, *a, **k: fn(self, *a, **k))(f))


_attach_methods()
