"""Host-side mirror of the reference's ndarray and of what happens between the
operator surface and `readdata`:

    XS stub -> pdl_run_<op> -> type_coerce -> make_trans_mutual -> redodims -> readdata
    (lib/PDL/PP.pm:1862-1869,1998-2014; lib/PDL/Core/pdlapi.c:746-869,1182-1327;
     lib/PDL/Core/pdlbroadcast.c:275-523)

Everything here is metadata (types, dims, strides, bad flags); the data stays on
the device in a Store and is only touched by libpdlb200 kernels.  Views (slice,
dummy, xchg, mv ...) are affine: they share the parent's Store and differ only in
dims / dimincs / offs, exactly like the reference's vaffine ndarrays
(lib/PDL/Core/pdlapi.c:956-1019), and reach the kernels unmaterialised.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import _abi, types as T
from .engine import Engine, PDLError, Store, default_engine

__all__ = ["PDL", "pdl", "zeroes", "ones", "sequence", "null", "PDLError"]


def _default_incs(dims: Sequence[int]) -> list[int]:
    incs, acc = [], 1
    for d in dims:
        incs.append(acc)
        acc *= d
    return incs


class PDL:
    """Device-resident ndarray.  dims[0] is the fastest-varying dim, as in PDL."""

    __slots__ = ("engine", "_store", "datatype", "dims", "dimincs", "offs", "_bf",
                 "_badvalue", "_inplace", "_null", "_flowing", "_pending")

    def __init__(self, engine: Engine, store: Store | None, datatype: int, dims, dimincs=None,
                 offs: int = 0, badflag: bool = False, badvalue=None):
        self.engine = engine
        self._store = store
        self._pending = None        # (op name, inputs): a deferred transformation whose readdata has not run yet
        self._flowing = False
        self.datatype = datatype
        self.dims = [int(d) for d in dims]
        self.dimincs = [int(i) for i in (dimincs if dimincs is not None else _default_incs(self.dims))]
        self.offs = int(offs)
        self._bf = False
        self.badflag = bool(badflag)
        self._badvalue = badvalue
        self._inplace = False
        self._null = False

    # ---- construction ---------------------------------------------------------------
    @classmethod
    def from_numpy(cls, arr, datatype: int | None = None, engine: Engine | None = None) -> "PDL":
        """numpy array (C order, shape reversed w.r.t. PDL dims) -> device ndarray."""
        engine = engine or default_engine()
        arr = np.asarray(arr)
        if datatype is None:
            datatype = T.from_numpy_dtype(arr.dtype)
        arr = np.ascontiguousarray(arr.astype(T.NP_DTYPE[datatype], copy=False))
        dims = list(reversed(arr.shape))
        st = engine.alloc(arr.nbytes)
        engine.upload(st, arr.reshape(-1).view(np.uint8))
        return cls(engine, st, datatype, dims)

    @classmethod
    def empty(cls, datatype: int, dims, engine: Engine | None = None) -> "PDL":
        engine = engine or default_engine()
        n = 1
        for d in dims:
            n *= d
        st = engine.alloc(n * T.SIZE[datatype])
        return cls(engine, st, datatype, dims)

    @classmethod
    def null(cls, engine: Engine | None = None) -> "PDL":
        p = cls(engine or default_engine(), None, T.D, [0])
        p._null = True
        return p

    # ---- bad state: per ndarray in the reference, but propagated through the vaffine family by
    # pdl_propagate_badflag_dir (pdlapi.c:28-36, make_trans_mutual :806-808): a view and its parent are the same
    # data, so the state lives with the shared buffer.  `x.slice(...).inplace().setvaltobad(0)` flags x too.
    @property
    def badflag(self) -> bool:
        st = self._store
        return st.bad if st is not None else self._bf

    @badflag.setter
    def badflag(self, flag) -> None:
        self._bf = bool(flag)
        st = self._store
        if st is not None:
            st.bad = bool(flag)

    # ---- dataflow: deferred readdata (lib/PDL/Core/pdlapi.c:781-801,890) ---------------
    @property
    def store(self):
        """The device allocation.  An ndarray produced by an op on a `flowing` parent has none until
        something needs its data: then the deferred transformation runs (pdl__ensure_trans from
        make_physical, pdlapi.c:886-905)."""
        if self._pending is not None:
            self._materialize()
        return self._store

    @store.setter
    def store(self, st):
        self._store = st

    def flowing(self) -> "PDL":
        """`$x->flowing`: turn on dataflow for the NEXT operation (PDL_DATAFLOW_F; cleared by
        make_trans_mutual, pdlapi.c:781-785).  That operation's readdata is deferred until its result is
        needed — which lets a reduction that is the result's only consumer run fused with it."""
        v = self._view(self.dims, self.dimincs, self.offs)
        v._flowing = True
        return v

    def is_pending(self) -> bool:
        return self._pending is not None

    def _materialize(self) -> None:
        name, ins = self._pending
        self._pending = None
        n = 1
        for d in self.dims:
            n *= d
        self._store = self.engine.alloc(n * T.SIZE[self.datatype])
        self._store.bad = self._bf
        from .trans import run_op
        run_op(name, ins, [self])

    # ---- introspection ----------------------------------------------------------------
    @property
    def ndims(self) -> int:
        return len(self.dims)

    def getndims(self) -> int:
        return len(self.dims)

    def dim(self, i: int) -> int:
        return self.dims[i]

    @property
    def nelem(self) -> int:
        n = 1
        for d in self.dims:
            n *= d
        return n

    @property
    def type(self) -> str:
        return T.NAMES[self.datatype]

    def isnull(self) -> bool:
        return self._null

    def is_contiguous(self) -> bool:
        return self.dimincs == _default_incs(self.dims) or self.nelem <= 1

    @property
    def badvalue(self):
        return self._badvalue if self._badvalue is not None else T.DEFAULT_BAD[self.datatype]

    def set_badvalue(self, v) -> "PDL":
        self._badvalue = v
        return self

    def badvalue_bits(self) -> int:
        return T.value_bits(self.datatype, self.badvalue)

    def badvalue_isnan(self) -> bool:
        v = self.badvalue
        if self.datatype in (T.CF, T.CD):
            v = complex(v)
            return math.isnan(v.real) or math.isnan(v.imag)
        return self.datatype in (T.F, T.D) and isinstance(v, float) and math.isnan(v)

    def set_badflag(self, flag: bool = True) -> "PDL":
        self.badflag = bool(flag)
        return self

    def inplace(self) -> "PDL":
        self._inplace = True
        return self

    def is_inplace(self) -> bool:
        return self._inplace

    # ---- host access (the lazy-sync edge of the device store) ---------------------------
    def to_numpy(self) -> np.ndarray:
        """Download.  A non-contiguous view is first made physical ON THE DEVICE (assgn)."""
        src = self if self.is_contiguous() else self.copy()
        nbytes = src.nelem * T.SIZE[src.datatype]
        raw = src.engine.download(_SubStore(src), nbytes) if src.offs else src.engine.download(src.store, nbytes)
        arr = raw.view(T.NP_DTYPE[src.datatype])
        return arr.reshape(list(reversed(src.dims))) if src.dims else arr.reshape(())

    def unpdl(self):
        return self.to_numpy().tolist()

    def sclr(self):
        if self.nelem != 1:
            raise PDLError("multielement ndarray in 'sclr' call")
        return self.to_numpy().reshape(-1)[0].item()

    def at(self, *idx):
        sub = self
        spec = ",".join(f"({i})" for i in idx)
        return sub.slice(spec).sclr()

    def bad_mask(self) -> np.ndarray:
        """Host-side boolean mask of BAD elements (test/diagnostic helper)."""
        a = self.to_numpy()
        if not self.badflag:
            return np.zeros(a.shape, dtype=bool)
        if self.badvalue_isnan():
            return np.isnan(a)
        return a == np.array(self.badvalue).astype(a.dtype)

    def __repr__(self) -> str:
        return f"PDL({self.type}, dims={self.dims}, badflag={int(self.badflag)})"

    # ---- affine views (lib/PDL/Slices.pd: slice / dummy / xchg / mv; metadata only) ------
    def _view(self, dims, dimincs, offs) -> "PDL":
        v = PDL(self.engine, self.store, self.datatype, dims, dimincs, offs, self.badflag, self._badvalue)
        return v

    def slice(self, spec: str) -> "PDL":
        """PDL::slice string syntax: ':' | 'n' | '(n)' | 'a:b' | 'a:b:s' | '*n' per dim."""
        parts = [s.strip() for s in spec.split(",")] if spec.strip() != "" else []
        dims, incs, offs = [], [], self.offs
        src = 0
        for part in parts:
            m = re.fullmatch(r"\*(\d*)", part)
            if m:
                dims.append(int(m.group(1) or 1))
                incs.append(0)
                continue
            if src >= self.ndims:
                if part in ("", ":", "0", "(0)", "0:0", "-1", "(-1)"):
                    if not (part.startswith("(")):
                        dims.append(1); incs.append(0)
                    src += 1
                    continue
                raise PDLError(f"slice: too many dims in slice '{spec}'")
            n, inc = self.dims[src], self.dimincs[src]
            squeeze = False
            if part.startswith("(") and part.endswith(")"):
                squeeze, part = True, part[1:-1]
            if part in ("", ":"):
                start, end, step = 0, n - 1, 1
                if n == 0:
                    dims.append(0); incs.append(inc); src += 1
                    continue
            else:
                f = part.split(":")
                if len(f) == 1:
                    start = end = int(f[0]); step = 1
                else:
                    start = int(f[0]) if f[0] != "" else 0
                    end = int(f[1]) if f[1] != "" else -1
                    step = int(f[2]) if len(f) > 2 and f[2] != "" else None
                if start < 0: start += n
                if end < 0: end += n
                if step is None:
                    step = 1 if end >= start else -1
                if not (0 <= start < n and 0 <= end < n) or step == 0:
                    raise PDLError(f"slice: '{part}' out of bounds for dim of size {n}")
            cnt = (end - start) // step + 1
            if cnt < 0:
                cnt = 0
            offs += start * inc
            if not squeeze:
                dims.append(cnt); incs.append(inc * step)
            src += 1
        for d in range(src, self.ndims):
            dims.append(self.dims[d]); incs.append(self.dimincs[d])
        return self._view(dims, incs, offs)

    def dummy(self, pos: int, size: int = 1) -> "PDL":
        nd = self.ndims
        if pos < 0:
            pos = nd + 1 + pos
        dims, incs = list(self.dims), list(self.dimincs)
        while len(dims) < pos:  # dummy beyond the end pads with size-1 dims
            dims.append(1); incs.append(0)
        dims.insert(pos, size); incs.insert(pos, 0)
        return self._view(dims, incs, self.offs)

    def xchg(self, a: int, b: int) -> "PDL":
        nd = self.ndims
        a, b = a % nd, b % nd
        dims, incs = list(self.dims), list(self.dimincs)
        dims[a], dims[b] = dims[b], dims[a]
        incs[a], incs[b] = incs[b], incs[a]
        return self._view(dims, incs, self.offs)

    def mv(self, a: int, b: int) -> "PDL":
        nd = self.ndims
        a, b = a % nd, b % nd
        dims, incs = list(self.dims), list(self.dimincs)
        d, i = dims.pop(a), incs.pop(a)
        dims.insert(b, d); incs.insert(b, i)
        return self._view(dims, incs, self.offs)

    def reorder(self, *order) -> "PDL":
        return self._view([self.dims[o] for o in order], [self.dimincs[o] for o in order], self.offs)

    def transpose(self) -> "PDL":
        if self.ndims == 0:
            return self.dummy(0).dummy(0)
        if self.ndims == 1:
            return self.dummy(0)
        return self.xchg(0, 1)

    def reshape_view(self, dims) -> "PDL":
        """Reinterpret a CONTIGUOUS ndarray with new dims (no copy)."""
        if not self.is_contiguous():
            raise PDLError("reshape_view needs a physical (contiguous) ndarray")
        n = 1
        for d in dims:
            n *= d
        if n != self.nelem:
            raise PDLError("reshape_view: element count mismatch")
        return self._view(list(dims), None, self.offs)

    def clump(self, n: int) -> "PDL":
        """clump the first n dims (-1: all).  The reference's _clump_int COPIES
        (lib/PDL/Slices.pd:1373-1402); a physical parent is clumped as a view here
        (SURVEY.md §8(f)2), a strided one is made physical on the device first."""
        nd = self.ndims
        if n < 0:
            n = nd + 1 + n
        n = max(0, min(n, nd))
        src = self if self.is_contiguous() else self.copy()
        lead = 1
        for d in src.dims[:n]:
            lead *= d
        return src._view([lead] + src.dims[n:], None, src.offs)

    def flat(self) -> "PDL":
        return self if self.ndims == 1 else self.clump(-1)

    def _as_words(self) -> "PDL":
        """A complex ndarray seen as 64-bit words (cfloat: one per element; cdouble: a leading dim of 2), for ops that
        only move bits."""
        if self.datatype == T.CF:
            return PDL(self.engine, self.store, T.LL, self.dims, self.dimincs, self.offs)
        return PDL(self.engine, self.store, T.LL, [2] + self.dims, [1] + [2 * i for i in self.dimincs], 2 * self.offs)

    def copy(self) -> "PDL":
        out = PDL.empty(self.datatype, self.dims, self.engine)
        out._badvalue = self._badvalue
        from .trans import run_op
        if self.datatype in (T.CF, T.CD):        # complex: a bit copy through the 64-bit assgn kernel
            run_op("assgn", [self._as_words()], [out._as_words()])
            out.badflag = self.badflag
            return out
        run_op("assgn", [self], [out])
        return out

    sever = copy

    def convert(self, datatype: int) -> "PDL":
        if datatype == self.datatype:
            return self
        from .trans import convert_type
        return convert_type(self, datatype)

    # ---- operator surface: overloads exactly as PDL::Ops declares them -----------------
    def _bin(self, name, other, swap=0):
        from .trans import run_biop
        return run_biop(name, self, other, None, swap)

    def __add__(self, o): return self._bin("plus", o)
    def __radd__(self, o): return self._bin("plus", o, 1)
    def __sub__(self, o): return self._bin("minus", o)
    def __rsub__(self, o): return self._bin("minus", o, 1)
    def __mul__(self, o): return self._bin("mult", o)
    def __rmul__(self, o): return self._bin("mult", o, 1)
    def __truediv__(self, o): return self._bin("divide", o)
    def __rtruediv__(self, o): return self._bin("divide", o, 1)
    def __gt__(self, o): return self._bin("gt", o)
    def __lt__(self, o): return self._bin("lt", o)
    def __ge__(self, o): return self._bin("ge", o)
    def __le__(self, o): return self._bin("le", o)
    def __eq__(self, o): return self._bin("eq", o)  # type: ignore[override]
    def __ne__(self, o): return self._bin("ne", o)  # type: ignore[override]
    __hash__ = None  # type: ignore[assignment]
    def __lshift__(self, o): return self._bin("shiftleft", o)
    def __rlshift__(self, o): return self._bin("shiftleft", o, 1)
    def __rshift__(self, o): return self._bin("shiftright", o)
    def __rrshift__(self, o): return self._bin("shiftright", o, 1)
    def __or__(self, o): return self._bin("or2", o)
    def __ror__(self, o): return self._bin("or2", o, 1)
    def __and__(self, o): return self._bin("and2", o)
    def __rand__(self, o): return self._bin("and2", o, 1)
    def __xor__(self, o): return self._bin("xor", o)
    def __rxor__(self, o): return self._bin("xor", o, 1)
    def __pow__(self, o): return self._bin("power", o)
    def __rpow__(self, o): return self._bin("power", o, 1)
    def __mod__(self, o): return self._bin("modulo", o)
    def __rmod__(self, o): return self._bin("modulo", o, 1)
    def __matmul__(self, o):
        from .primitive import matmult
        return matmult(self, o)

    def _ibin(self, name, other):
        from .trans import run_biop
        run_biop(name, self, other, self, 0)
        return self

    def __iadd__(self, o): return self._ibin("plus", o)
    def __isub__(self, o): return self._ibin("minus", o)
    def __imul__(self, o): return self._ibin("mult", o)
    def __itruediv__(self, o): return self._ibin("divide", o)

    def __invert__(self):
        from .trans import run_ufunc
        return run_ufunc("bitnot", self)

    def __neg__(self):
        return self._bin("minus", 0, 1)  # PDL: '-' unary is 0 - $a (PDL::Core neg overload)

    def __abs__(self):
        from .trans import run_ufunc
        return run_ufunc("_rabs", self)

    def assign(self, other) -> "PDL":
        """`$self .= $other` (assgn, lib/PDL/Ops.pd:382-397)."""
        from .trans import run_op, as_pdl
        run_op("assgn", [as_pdl(other, self.engine)], [self])
        return self


class _SubStore:
    """Pointer view into a Store for downloads that start at an element offset."""

    def __init__(self, p: PDL):
        self.engine = p.engine
        self.ptr = p.store.ptr + p.offs * T.SIZE[p.datatype]
        self.nbytes = p.nelem * T.SIZE[p.datatype]
        self.handle = None


# ---- constructors in the reference's spelling ------------------------------------------

def pdl(data, datatype: int | None = None, engine: Engine | None = None) -> PDL:
    """pdl(...) — like the reference, untyped numeric data becomes double (PDL::Core::pdl)."""
    if isinstance(data, PDL):
        return data if datatype is None else data.convert(datatype)
    a = np.asarray(data)
    if datatype is None:
        datatype = T.D if a.dtype.kind in "fiub" and not isinstance(data, np.ndarray) else T.from_numpy_dtype(a.dtype)
    return PDL.from_numpy(a, datatype, engine)


def zeroes(datatype: int, *dims, engine: Engine | None = None) -> PDL:
    return PDL.from_numpy(np.zeros(list(reversed(dims)), dtype=T.NP_DTYPE[datatype]), datatype, engine)


def ones(datatype: int, *dims, engine: Engine | None = None) -> PDL:
    return PDL.from_numpy(np.ones(list(reversed(dims)), dtype=T.NP_DTYPE[datatype]), datatype, engine)


def sequence(datatype: int, *dims, engine: Engine | None = None) -> PDL:
    n = 1
    for d in dims:
        n *= d
    return PDL.from_numpy(np.arange(n).astype(T.NP_DTYPE[datatype]).reshape(list(reversed(dims))), datatype, engine)


def null(engine: Engine | None = None) -> PDL:
    return PDL.null(engine)
