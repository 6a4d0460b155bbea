"""pdl_run_<op> for the device path: everything the reference does between the XS stub
and readdata, restated on PDL metadata, ending in ONE pdlb200_readdata call.

Reference call stack being mirrored:
  type selection     pdl__transtype_select        lib/PDL/Core/pdlapi.c:1182-1236
  output typing      pdl__set_output_type_badvalue lib/PDL/Core/pdlapi.c:1240-1260
  input conversion   pdl__type_convert            lib/PDL/Core/pdlapi.c:1262-1311 (on device: OP_CONVERT)
  bvalflag/badflag   pdl_make_trans_mutual        lib/PDL/Core/pdlapi.c:760-770,806-808
  named dims         pdl_dim_checks               lib/PDL/Core/pdlbroadcast.c:211-273
  broadcast dims     pdl_initbroadcaststruct      lib/PDL/Core/pdlbroadcast.c:340-488
  output creation    pdl_broadcast_create_parameter lib/PDL/Core/pdlbroadcast.c:490-523
  real-dim incs      pdl_redodims_default         lib/PDL/Core/pdlapi.c:858-864
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _abi, types as T
from .core import PDL, _default_incs
from .engine import Engine, PDLError, default_engine


@dataclass
class Par:
    name: str
    realdims: tuple = ()        # named dims, e.g. ('n',) or ('t','h')
    out: bool = False
    typed: int | None = None    # forced type (double, indx, long for int+)
    tplus: bool = False         # `int+`: max(typed, trans type)


@dataclass
class OpSpec:
    name: str
    pars: list
    gentypes: tuple             # GenericTypes, in declaration order (last = fallback)
    kind: str                   # biop | bifunc | ufunc | reduce | matmult | convert
    inplace: tuple = ()         # parameter names that may run in place
    fixed: tuple = ()           # named dims of fixed size, e.g. (("r", 4),)
    opid: int = field(init=False)

    def __post_init__(self):
        self.opid = _abi.OPS[self.name]


_A, _R, _I = T.ALL, T.REAL, T.INTEGER
# Ops.pd:8-9,321,332: [@$C, @$F] with D last so that non-float input falls back to double
_CF = T.COMPLEX + T.FLOATING
_F = T.FLOATING
_AF = (T.F, T.LD) + T.COMPLEX + (T.D,)

# what the op's Code does to the output state after make_trans_mutual propagated the input flags
_STATE_SETBAD = ("setbadif", "setvaltobad")                       # $PDLSTATESETBAD(b), unconditionally
_STATE_SETGOOD = ("setbadtonan", "setbadtoval", "badmask")        # $PDLSTATESETGOOD(b)
_STATE_IFFLAG = ("setnantobad", "setinftobad", "setnonfinitetobad",  # if (flag) $PDLSTATESETBAD(b)
                 "minmaximum",                                       # a row without a usable element (Ufunc.pd:578-583)
                 "minimum_n_ind", "maximum_n_ind")                   # a slot that could not be filled (Ufunc.pd:531-533)
_STATE_GOOD_UNLESS_FLAG = ("minimum_n_ind", "maximum_n_ind")         # $PDLSTATESETGOOD(c) first (Ufunc.pd:521)


def _bi(name, gentypes, kind="biop"):
    return OpSpec(name, [Par("a"), Par("b"), Par("c", out=True)], gentypes, kind, inplace=("a",))


def _un(name, gentypes):
    return OpSpec(name, [Par("a"), Par("b", out=True)], gentypes, "ufunc", inplace=("a",))


def _rd(name, gentypes, typed=None, tplus=False):
    return OpSpec(name, [Par("a", ("n",)), Par("b", out=True, typed=typed, tplus=tplus)], gentypes, "reduce")


SPECS = {s.name: s for s in [
    # biop, lib/PDL/Ops.pd:288-313
    _bi("plus", _A), _bi("mult", _A), _bi("minus", _A), _bi("divide", _A),
    _bi("gt", _R), _bi("lt", _R), _bi("le", _R), _bi("ge", _R), _bi("eq", _A), _bi("ne", _A),
    _bi("shiftleft", _I), _bi("shiftright", _I), _bi("or2", _I), _bi("and2", _I), _bi("xor", _I),
    # bifunc, lib/PDL/Ops.pd:321-324
    _bi("power", _CF, "bifunc"), _bi("atan2", _F, "bifunc"), _bi("modulo", _R, "bifunc"),
    _bi("spaceship", _R, "bifunc"),
    # ufunc etc., lib/PDL/Ops.pd:327-397,491-503
    _un("bitnot", _I), _un("sqrt", _A), _un("sin", _A), _un("cos", _A), _un("not", _R),
    _un("exp", _CF), _un("log", _CF), _un("log10", _A), _un("_rabs", _R),
    OpSpec("assgn", [Par("a"), Par("b", out=True)], _A, "ufunc"),
    OpSpec("abs2", [Par("a"), Par("b", out=True)], _A, "ufunc"),
    # reductions, lib/PDL/Ufunc.pd:88-118,143-187,413-500
    _rd("sumover", _A, T.L, True), _rd("prodover", _A, T.L, True),
    _rd("dsumover", _R, T.D), _rd("dprodover", _R, T.D),
    _rd("average", _R, T.L, True), _rd("daverage", _R, T.D),
    _rd("minimum", _R), _rd("maximum", _R),
    _rd("minimum_ind", _R, T.IND), _rd("maximum_ind", _R, T.IND),
    _rd("andover", _A), _rd("orover", _A), _rd("zcover", _A), _rd("xorover", _A),
    _rd("bandover", _I), _rd("borover", _I), _rd("bxorover", _I),
    # minmaximum, lib/PDL/Ufunc.pd:563-613 (the body of minmax, :738); magnover, :1235-1256 (GenericTypes "F last")
    OpSpec("minmaximum", [Par("a", ("n",)), Par("cmin", out=True), Par("cmax", out=True),
                          Par("cmin_ind", out=True, typed=T.IND), Par("cmax_ind", out=True, typed=T.IND)], _R, "reduce"),
    OpSpec("magnover", [Par("a", ("n",)), Par("b", out=True)], (T.D, T.LD) + T.COMPLEX + (T.F,), "reduce"),
    # minimum_n_ind / maximum_n_ind, lib/PDL/Ufunc.pd:502-561: a(n); indx [o]c(m), OtherPars m_size => m
    OpSpec("minimum_n_ind", [Par("a", ("n",)), Par("c", ("m",), out=True, typed=T.IND)], _R, "nind"),
    OpSpec("maximum_n_ind", [Par("a", ("n",)), Par("c", ("m",), out=True, typed=T.IND)], _R, "nind"),
    # lib/PDL/Bad.pd:418-480
    _rd("nbadover", _A, T.IND), _rd("ngoodover", _A, T.IND),
    # scans, lib/PDL/Ufunc.pd:120-141 : a(n); [o]b(n)
    OpSpec("cumusumover", [Par("a", ("n",)), Par("b", ("n",), out=True, typed=T.L, tplus=True)], _A, "scan"),
    OpSpec("cumuprodover", [Par("a", ("n",)), Par("b", ("n",), out=True, typed=T.L, tplus=True)], _A, "scan"),
    OpSpec("dcumusumover", [Par("a", ("n",)), Par("b", ("n",), out=True, typed=T.D)], _R, "scan"),
    OpSpec("dcumuprodover", [Par("a", ("n",)), Par("b", ("n",), out=True, typed=T.D)], _R, "scan"),
    # ipow, lib/PDL/Ops.pd:443-476: GenericTypes [P Q, non-integer types with D last]
    OpSpec("ipow", [Par("a"), Par("b", typed=T.LL), Par("ans", out=True)],
           (T.ULL, T.LL, T.F, T.LD, T.CF, T.CD, T.CLD, T.D), "ufunc", inplace=("a", "ans")),
    # lib/PDL/Bad.pd:343-416,584-905.  $AF = F LD CF CD CLD D ("so defaults to D if non-float given", Bad.pd:6-7)
    OpSpec("isbad", [Par("a"), Par("b", out=True, typed=T.L)], _A, "badop"),
    OpSpec("isgood", [Par("a"), Par("b", out=True, typed=T.L)], _A, "badop"),
    OpSpec("isnan", [Par("a"), Par("b", out=True, typed=T.L)], _A, "badop"),
    OpSpec("setbadif", [Par("a"), Par("mask", typed=T.L), Par("b", out=True)], _A, "badop"),
    OpSpec("setvaltobad", [Par("a"), Par("b", out=True)], _A, "badop", inplace=("a",)),
    OpSpec("setnantobad", [Par("a"), Par("b", out=True)], _AF, "badop", inplace=("a",)),
    OpSpec("setinftobad", [Par("a"), Par("b", out=True)], _AF, "badop", inplace=("a",)),
    OpSpec("setnonfinitetobad", [Par("a"), Par("b", out=True)], _AF, "badop", inplace=("a",)),
    OpSpec("setbadtonan", [Par("a"), Par("b", out=True)], _AF, "badop", inplace=("a",)),
    OpSpec("setbadtoval", [Par("a"), Par("b", out=True)], _A, "badop", inplace=("a",)),
    OpSpec("badmask", [Par("a"), Par("b"), Par("c", out=True)], _R, "badop", inplace=("a",)),
    OpSpec("copybad", [Par("a"), Par("mask"), Par("b", out=True)], _A, "badop", inplace=("a",)),
    # axisvalues, lib/PDL/Primitive.pd:1468-1474; inner, :48-70
    OpSpec("axisvalues", [Par("i", ("n",)), Par("a", ("n",), out=True)], _A, "axis", inplace=("i",)),
    OpSpec("inner", [Par("a", ("n",)), Par("b", ("n",)), Par("c", out=True)], _A, "inner"),
    OpSpec("outer", [Par("a", ("n",)), Par("b", ("m",)), Par("c", ("n", "m"), out=True)], _A, "outer"),
    # per-rank partial records of a sharded whole-array reduction (include/pdlb200.h PART_*): a(n); longlong [o]rec(r=4)
    *[OpSpec(n, [Par("a", ("n",)), Par("rec", ("r",), out=True, typed=T.LL)], g, "part", fixed=(("r", 4),))
      for n, g in (("part_sum", _A), ("part_dsum", _R), ("part_min", _R), ("part_max", _R))],
    # matmult, lib/PDL/Primitive.pd:191-195
    OpSpec("matmult", [Par("a", ("t", "h")), Par("b", ("w", "t")), Par("c", ("w", "h"), out=True)], _A, "matmult"),
]}


_SCALARS: dict = {}


# ---- scalars -> ndarrays (pdl_SvPDLV + pdl_scalar, lib/PDL/Core/pdlcore.c:64-78) -----------

def as_pdl(x, engine: Engine | None = None) -> PDL:
    if isinstance(x, PDL):
        return x
    engine = engine or default_engine()
    if isinstance(x, (bool, int, float, np.integer, np.floating)):
        # `$y * 2`: the 0-dim ndarray of a Perl scalar (pdl_scalar, pdlcore.c:64-78).  Uploading 8 bytes costs a
        # host-device round trip, so the ndarrays of recently used scalars are kept (they are only ever read).
        key = (id(engine), type(x) is float or isinstance(x, np.floating), x)
        p = _SCALARS.get(key) if x == x else None
        if p is None:
            t = T.scalar_type(x)
            p = PDL.from_numpy(np.array(x, dtype=T.NP_DTYPE[t]), t, engine)
            if x == x:
                if len(_SCALARS) >= 512:
                    _SCALARS.clear()
                _SCALARS[key] = p
        # a fresh ndarray object over the same bytes, behind its OWN Store record: the bad state lives with the Store, and
        # a caller flagging its scalar must not flag everybody else's
        from .engine import Store
        st = p.store
        return PDL(engine, Store(engine, None, st.ptr, st.nbytes, keep=st), p.datatype, p.dims, p.dimincs, p.offs)
    if isinstance(x, (list, tuple, np.ndarray)):
        from .core import pdl
        return pdl(x, engine=engine)
    raise PDLError(f"Error - tried to use an unknown data structure as a PDL: {type(x).__name__}")


# ---- type selection ---------------------------------------------------------------------------

def transtype_select(spec: OpSpec, pdls: list) -> int:
    """pdl__transtype_select (pdlapi.c:1182-1236): outputs that already exist decide first,
    then the highest input type that is in GenericTypes; types above the list's last entry
    (or none available) fall back to the LAST GenericTypes entry."""
    avail = set(spec.gentypes)
    last = spec.gentypes[-1]
    if spec.gentypes[0] == last:
        return last
    nparents = sum(1 for p in spec.pars if not p.out)
    retval, use_last = -1, False
    for i in range(len(spec.pars) - 1, -1, -1):
        par, p = spec.pars[i], pdls[i]
        if p is not None and not p.isnull() and par.typed is None:
            new = p.datatype
            ok = new in avail
            if not ok and new > last:
                use_last = True
            if ok and retval < new:
                retval = new
        if i == nparents and retval != -1:
            return retval
    if use_last or retval == -1 or retval not in avail:
        retval = last
    return retval


def par_type(par: Par, transtype: int) -> int:
    """PDL_TYPE_ADJUST_FROM_TRANS (pdlapi.c:1174-1181)."""
    if par.typed is not None:
        return max(par.typed, transtype) if par.tplus else par.typed
    return transtype


def convert_type(p: PDL, datatype: int) -> PDL:
    """converttype on the device (lib/PDL/Core/pdlconv.c:45-126): BAD maps to the TARGET
    type's default badvalue; the result has no per-ndarray badvalue."""
    if datatype in (T.CF, T.CD) and T.is_device_type(p.datatype) or (p.datatype in (T.CF, T.CD) and datatype in (T.CF, T.CD)):
        # real -> complex (imaginary part 0) and complex -> complex: host-side glue of the mirror (the Perl binding
        # leaves conversions that involve complex types to the reference's own converttypei)
        src = p.to_numpy()
        out = PDL.from_numpy(src.astype(T.NP_DTYPE[datatype]), datatype, p.engine)
        out.badflag = p.badflag
        return out
    if not T.is_device_type(datatype) or not T.is_device_type(p.datatype):
        raise PDLError(f"type {T.NAMES[datatype]} has no device representation")
    out = PDL.empty(datatype, p.dims, p.engine)
    out.badflag = p.badflag
    _launch(SPECS_CONVERT, p.datatype, [p, out], _broadcast([p, out], [0, 0], [False, False], "converttype"),
            {}, bval=p.badflag)
    return out


SPECS_CONVERT = OpSpec.__new__(OpSpec)
SPECS_CONVERT.name, SPECS_CONVERT.kind, SPECS_CONVERT.opid = "converttype", "convert", _abi.OPS["converttype"]
SPECS_CONVERT.pars = [Par("a"), Par("b", out=True)]
SPECS_CONVERT.gentypes, SPECS_CONVERT.inplace = T.ALL, ()


# ---- broadcast struct ---------------------------------------------------------------------------

@dataclass
class Broadcast:
    dims: list          # broadcast.dims (ndims >= 2: nobl=2, pdlapi.c:839)
    nimpl: int
    incs: list          # incs[d][p]


def _broadcast(pdls: list, realdims: list, creating: list, opname: str) -> Broadcast:
    """pdl_initbroadcaststruct + pdl_broadcast_dim_checks for implicit broadcast dims."""
    np_ = len(pdls)
    nimpl = 0
    for j, p in enumerate(pdls):
        if creating[j]:
            continue
        nimpl = max(nimpl, p.ndims - realdims[j])
    ndims = max(nimpl, 2)
    dims = [1] * ndims
    incs = [[0] * np_ for _ in range(ndims)]
    for nth in range(nimpl):
        for j, p in enumerate(pdls):
            if creating[j]:
                continue
            k = nth + realdims[j]
            if k >= p.ndims:
                continue
            cur = p.dims[k]
            if cur != 1:
                if dims[nth] != 1:
                    if dims[nth] != cur:
                        raise PDLError(
                            f"PDL: PDL::{opname}(...): Parameter '{j}':\n"
                            f"Mismatched implicit broadcast dimension {nth}: size {dims[nth]} vs. {cur}")
                else:
                    dims[nth] = cur
                incs[nth][j] = p.dimincs[k]
    return Broadcast(dims, nimpl, incs)


def _check_output_dims(p: PDL, realdims: int, bc: Broadcast, opname: str, pname: str) -> None:
    """An existing output cannot be broadcast over (pdlbroadcast.c:286-306)."""
    for nth in range(bc.nimpl):
        k = nth + realdims
        if k >= p.ndims:
            if bc.dims[nth] != 1:
                raise PDLError(f"PDL::{opname}: implicit dim {nth} size {bc.dims[nth]}, "
                               f"can't broadcast over output ndarray with size > 1")
            continue
        if p.dims[k] == 1 and bc.dims[nth] != 1:
            raise PDLError(f"PDL::{opname}: implicit dim {nth} size {bc.dims[nth]}, but dim has size 1")
        if p.dims[k] != 1 and p.dimincs[k] == 0:
            raise PDLError(f"PDL::{opname}: implicit dim {nth} size {bc.dims[nth]}, but dim is dummy")


# ---- the descriptor ---------------------------------------------------------------------------
_FLAG_DEFERRABLE = ("minmaximum", "setnantobad", "setinftobad", "setnonfinitetobad")   # ops honouring DEFER_ANYBAD


def _readdata_flagged(engine, tr, name):
    """One readdata of an op with a data-dependent output badflag.  -> the flag (0/1), or the FlagRing slot as
    a 1-tuple when the engine defers the read-back (no stream synchronise inside the call)."""
    ring = getattr(engine, "flag_ring", None) if name in _FLAG_DEFERRABLE else None
    if ring is None:
        if tr.tflags:                                  # a cached descriptor last used by a deferring engine
            import ctypes as _C
            tr.anybad, tr.tflags = _C.pointer(tr._anybad), 0
        engine.readdata(tr)
        return int(tr._anybad.value)
    slot, ptr = ring.take()
    tr.anybad, tr.tflags = ptr, _abi.TRANS_DEFER_ANYBAD
    engine.readdata(tr)
    return (slot,)


def _apply_flag(engine, flag, outs) -> None:
    """`if (flag) $PDLSTATESETBAD(out)` for every output — now, or when the deferred flag is first needed."""
    if isinstance(flag, tuple):
        engine.flag_ring.attach(flag[0], [o.store for o in outs])
    elif flag:
        for o in outs:
            o.badflag = True



class Prepared:
    """A transformation whose descriptor is already filled in: calling it is ONE C-ABI call.
    The cached-descriptor fast path for repeated shapes (SURVEY.md §7 "launch-latency floor"):
    type selection, broadcast merging and output creation were done once by prepare_op()."""

    __slots__ = ("engine", "trans", "outputs", "_keep", "_flagged")

    def __init__(self, engine, trans, outputs, keep, flagged=False):
        self.engine, self.trans, self.outputs, self._keep, self._flagged = engine, trans, outputs, keep, flagged

    def __call__(self):
        if self._flagged:                                  # `if (flag) $PDLSTATESETBAD(...)` of the op's Code
            _apply_flag(self.engine, _readdata_flagged(self.engine, self.trans, self._flagged), self.outputs)
        else:
            self.engine.readdata(self.trans)
        return self.outputs


def _launch(spec: OpSpec, transtype: int, pdls: list, bc: Broadcast, named: dict, bval: bool) -> int:
    """One readdata.  Returns the op's `flag` (ops with a data-dependent output badflag), else 0."""
    tr = _build_trans(spec, transtype, pdls, bc, named, bval)
    if spec.name in _STATE_IFFLAG:
        return _readdata_flagged(pdls[0].engine, tr, spec.name)
    pdls[0].engine.readdata(tr)
    return 0


def _build_trans(spec: OpSpec, transtype: int, pdls: list, bc: Broadcast, named: dict, bval: bool) -> _abi.Trans:
    if len(bc.dims) > _abi.MAXDIMS:
        raise PDLError(f"PDL::{spec.name}: more than {_abi.MAXDIMS} broadcast dims")
    tr = _abi.Trans()
    tr.op, tr.datatype, tr.bvalflag = spec.opid, transtype, int(bool(bval))
    tr.npdls, tr.ndims = len(pdls), len(bc.dims)
    for d, n in enumerate(bc.dims):
        tr.dims[d] = n
        for j in range(len(pdls)):
            tr.incs[d * len(pdls) + j] = bc.incs[d][j]
    for k, v in enumerate(named.get("ind", ())):
        tr.ind[k] = v
    for k, v in enumerate(named.get("rinc", ())):
        tr.rinc[k] = v
    tr.param = float(named.get("param", 0.0))
    if spec.name in _STATE_IFFLAG:
        import ctypes as _C
        tr._anybad = _C.c_int32(0)      # kept alive by the descriptor that points at it
        tr.anybad = _C.pointer(tr._anybad)
    for j, p in enumerate(pdls):
        par = tr.pdls[j]
        par.data = p.store.ptr if p.store is not None else None
        par.offs = p.offs
        par.type = p.datatype
        par.badval = p.badvalue_bits()
        par.flags = (_abi.PAR_BADFLAG if p.badflag else 0) | (_abi.PAR_BADNAN if p.badvalue_isnan() else 0)
    return tr


def _real_inc(p: PDL, j: int) -> int:
    # pdl_redodims_default (pdlapi.c:858-864)
    return 0 if (p.ndims <= j or p.dims[j] <= 1) else p.dimincs[j]


# ---- descriptor cache: the fast path for repeated shapes (SURVEY.md §7 "launch-latency floor") ---------------
# Everything run_op derives — type selection, broadcast merging, named dims, output typing, the flag rules — is a
# function of (op, per-parameter type/dims/strides/flags, OtherPars).  The filled descriptor is kept under that key;
# a later call with the same key only patches the data pointers, creates the outputs and makes the ONE C-ABI call.
class _Cached:
    __slots__ = ("tr", "pars", "nin", "outs", "bval", "force_bad", "name", "ifflag")

    def __init__(self, tr, nin, outs, bval, force_bad, name):
        self.tr, self.nin, self.outs, self.bval, self.force_bad, self.name = tr, nin, outs, bval, force_bad, name
        self.pars = [tr.pdls[j] for j in range(tr.npdls)]          # aliases into the descriptor's memory
        self.ifflag = name in _STATE_IFFLAG


_DESC_CACHE: dict = {}


def _cache_key(name, ins, outs, param, goff, sizes):
    key = [name, param, goff, tuple(sorted(sizes.items())) if sizes else None]
    for p in ins:
        bv = p._badvalue
        if bv is not None and bv != bv:            # NaN never compares equal: not a usable key
            return None
        key += (p.datatype, tuple(p.dims), tuple(p.dimincs), p.badflag, bv)
    for p in outs:
        if p is None:
            key.append(None)
            continue
        bv = p._badvalue
        if bv is not None and bv != bv:
            return None
        key += (-1, p.datatype, tuple(p.dims), tuple(p.dimincs), p.badflag, bv)
    return tuple(key)


def _replay(ent: _Cached, ins: list, outs: list, engine) -> list:
    pars, nin = ent.pars, ent.nin
    for j, x in enumerate(ins):
        par = pars[j]
        st = x.store
        par.data = st.ptr if st is not None else None
        par.offs = x.offs
    res = []
    for k, o in enumerate(outs):
        if o is None:
            t, dims, incs, nbytes = ent.outs[k]
            o = PDL(engine, engine.alloc(nbytes), t, dims, incs)
        par = pars[nin + k]
        st = o.store
        par.data = st.ptr if st is not None else None
        par.offs = o.offs
        if ent.bval or ent.force_bad:
            o.badflag = True
        res.append(o)
    name = ent.name
    flag = 0
    if ent.ifflag:
        flag = _readdata_flagged(engine, ent.tr, name)
    else:
        engine.readdata(ent.tr)
    if name in _STATE_GOOD_UNLESS_FLAG:
        for o in res:
            o.badflag = False
    if ent.ifflag and flag:
        _apply_flag(engine, flag, res)
    elif name in _STATE_SETBAD:
        for o in res:
            o.badflag = True
    elif name in _STATE_SETGOOD:
        for o in res:
            o.badflag = False
    return res


_NO_OUTS = {n: [None] * sum(1 for p in sp.pars if p.out) for n, sp in SPECS.items()}


def prepare_op(name: str, inputs: list, outputs: list | None = None, param: float = 0.0) -> Prepared:
    """Like run_op, but returns a Prepared instead of launching.  Only for calls that need no
    type conversion of inputs or outputs (a single readdata)."""
    return run_op(name, inputs, outputs, _prepare=True, param=param)


def run_op(name: str, inputs: list, outputs: list | None = None, _prepare: bool = False, param: float = 0.0,
           goff: int = 0, sizes: dict | None = None):
    """pdl_run_<name>(inputs..., outputs...).  `outputs` entries may be None (null ndarray:
    created with the broadcast dims).  Returns the output ndarrays."""
    # ---- fast path: every parameter is already an ndarray and this (op, shapes, flags) was seen before ----
    if not _prepare:
        fast = True
        for x in inputs:
            if x.__class__ is not PDL or x._null:
                fast = False
                break
        if fast:
            outs = [None if (o is None or o._null) else o for o in outputs] if outputs is not None else _NO_OUTS[name]
            key = _cache_key(name, inputs, outs, param, goff, sizes)
            ent = _DESC_CACHE.get(key) if key is not None else None
            if ent is not None:
                return _replay(ent, inputs, outs, inputs[0].engine)
    spec = SPECS[name]
    in_pars = [p for p in spec.pars if not p.out]
    out_pars = [p for p in spec.pars if p.out]
    if len(inputs) != len(in_pars):
        raise PDLError(f"PDL::{name}: expected {len(in_pars)} inputs")
    outputs = list(outputs) if outputs is not None else [None] * len(out_pars)
    engine = next((x.engine for x in list(inputs) + outputs if isinstance(x, PDL)), None) or default_engine()
    ins = [as_pdl(x, engine) for x in inputs]
    for x in ins:
        if x.isnull():
            raise PDLError(f"PDL::{name}: input parameter is null")
    outs = [None if (o is None or o.isnull()) else o for o in outputs]
    key = None if _prepare else _cache_key(name, ins, outs, param, goff, sizes)
    if key is not None:
        ent = _DESC_CACHE.get(key)
        if ent is not None:
            return _replay(ent, ins, outs, engine)
    ins_given = ins

    all_pdls = ins + outs
    transtype = transtype_select(spec, all_pdls)
    if not T.is_device_type_for(transtype, name):
        raise PDLError(f"PDL::{name}: type {T.NAMES[transtype]} has no device representation "
                       "(long double, and complex outside + - * /, are outside the device type matrix)")
    # inputs to the type the loop is instantiated for (converttypei -> device convert kernel)
    if _prepare and any(x.datatype != par_type(par, transtype) for x, par in zip(ins, in_pars)):
        raise PDLError(f"PDL::{name}: prepare_op needs inputs already in the operation's type")
    ins = [x if x.datatype == par_type(par, transtype) else convert_type(x, par_type(par, transtype))
           for x, par in zip(ins, in_pars)]
    bval = any(x.badflag for x in ins)

    # named dims (pdl_dim_checks): inputs define them; size-1 stretches; missing dims promote to 1
    ind: dict = dict(spec.fixed)
    ind.update(sizes or {})         # named dims given as OtherPars (`PDL_Indx m_size => m`)
    for x, par in zip(ins, in_pars):
        for j, dn in enumerate(par.realdims):
            sz = x.dims[j] if j < x.ndims else 1
            if dn not in ind or ind[dn] == 1:
                ind[dn] = sz
            elif sz != ind[dn] and sz != 1:
                raise PDLError(f"PDL::{name}: Parameter '{par.name}' index '{dn}' size {ind[dn]}, "
                               f"but ndarray dim has size {sz}")
    for o, par in zip(outs, out_pars):
        if o is None:
            continue
        for j, dn in enumerate(par.realdims):
            sz = o.dims[j] if j < o.ndims else 1
            if dn in ind and sz != ind[dn]:
                raise PDLError(f"PDL::{name}: Parameter '{par.name}' index '{dn}' size {ind[dn]}, "
                               f"but ndarray dim has size {sz}")

    realdims = [len(p.realdims) for p in spec.pars]
    creating = [False] * len(ins) + [o is None for o in outs]
    placeholder = [x for x in ins] + [o if o is not None else ins[0] for o in outs]
    bc = _broadcast(placeholder, realdims, creating, name)

    final_outs, temps = [], []
    for k, (o, par) in enumerate(zip(outs, out_pars)):
        want = par_type(par, transtype)
        if o is None:
            dims = [ind[dn] for dn in par.realdims] + bc.dims[:bc.nimpl]
            o = PDL.empty(want, dims, engine)
            target = o
        else:
            _check_output_dims(o, len(par.realdims), bc, name, par.name)
            # an existing output of another type: compute in the op type, convert back afterwards
            target = o if o.datatype == want else PDL.empty(want, o.dims, engine)
            if target is not o:
                temps.append((target, o))
        final_outs.append(o)
        j = len(ins) + k
        rd = len(par.realdims)
        for nth in range(bc.nimpl):
            kk = nth + rd
            bc.incs[nth][j] = target.dimincs[kk] if (kk < target.ndims and target.dims[kk] != 1) else 0
        placeholder[j] = target

    # pdl_make_trans_mutual: any BAD input -> every output carries the badflag (pdlapi.c:806-808)
    if bval:
        for o in placeholder[len(ins):]:
            o.badflag = True

    named = {}
    if spec.kind == "badop":
        named = {"param": param}
    elif spec.kind == "axis":
        i, a = placeholder
        named = {"ind": [ind["n"]], "rinc": [_real_inc(i, 0), _real_inc(a, 0)]}
    elif spec.kind == "outer":
        a, b, c = placeholder
        named = {"ind": [ind["n"], ind["m"]], "rinc": [_real_inc(a, 0), _real_inc(b, 0), _real_inc(c, 0), _real_inc(c, 1)]}
    elif spec.kind == "inner":
        a, b, _c = placeholder
        named = {"ind": [ind["n"]], "rinc": [_real_inc(a, 0), _real_inc(b, 0)]}
    elif spec.kind == "nind":
        a, c = placeholder
        if ind["m"] > ind["n"]:
            raise PDLError(f"Error in {name}:m_size > n_size")      # RedoDimsCode, Ufunc.pd:518
        named = {"ind": [ind["n"], ind["m"]], "rinc": [_real_inc(a, 0), _real_inc(c, 0)]}
    elif spec.kind == "part":
        a, rec = placeholder
        named = {"ind": [ind["n"], int(goff)], "rinc": [_real_inc(a, 0), _real_inc(rec, 0)]}
    if spec.kind == "reduce":
        a = placeholder[0]
        named = {"ind": [ind["n"]], "rinc": [_real_inc(a, 0)]}
        # minimum/maximum and the and/or..over family set the output badflag themselves when a row has
        # no good element (Ufunc.pd:463-464,177); without BAD inputs that can only be n == 0.
        if ind["n"] == 0 and name in ("minimum", "maximum", "minimum_ind", "maximum_ind"):
            for o in placeholder[len(ins):]:
                o.badflag = True
    elif spec.kind == "scan":
        a, b = placeholder
        named = {"ind": [ind["n"]], "rinc": [_real_inc(a, 0), _real_inc(b, 0)]}
    elif spec.kind == "matmult":
        a, b, c = placeholder
        named = {"ind": [ind["t"], ind["h"], ind["w"]],
                 "rinc": [_real_inc(a, 0), _real_inc(a, 1), _real_inc(b, 0), _real_inc(b, 1),
                          _real_inc(c, 0), _real_inc(c, 1)]}
    if _prepare:
        if temps:
            raise PDLError(f"PDL::{name}: prepare_op needs outputs already in the operation's type")
        if name in _STATE_SETBAD:
            for o in placeholder[len(ins):]:
                o.badflag = True
        elif name in _STATE_SETGOOD:
            for o in placeholder[len(ins):]:
                o.badflag = False
        return Prepared(engine, _build_trans(spec, transtype, placeholder, bc, named, bval), final_outs, placeholder,
                        flagged=name if name in _STATE_IFFLAG else False)
    tr = _build_trans(spec, transtype, placeholder, bc, named, bval)
    flag = 0
    if name in _STATE_IFFLAG:
        flag = _readdata_flagged(engine, tr, name)
    else:
        engine.readdata(tr)
    if key is not None and not temps and all(a is b for a, b in zip(ins, ins_given)):
        # nothing was converted: the descriptor is reusable for every later call with the same key
        if len(_DESC_CACHE) >= 4096:
            _DESC_CACHE.clear()
        force_bad = spec.kind == "reduce" and ind.get("n") == 0 and name in ("minimum", "maximum", "minimum_ind", "maximum_ind")
        meta = [(o.datatype, list(o.dims), list(o.dimincs), o.nelem * T.SIZE[o.datatype]) for o in placeholder[len(ins):]]
        _DESC_CACHE[key] = _Cached(tr, len(ins), meta, bval, force_bad, name)
    if name in _STATE_GOOD_UNLESS_FLAG:
        for o in placeholder[len(ins):]:
            o.badflag = False
    if name in _STATE_IFFLAG and flag:
        _apply_flag(engine, flag, placeholder[len(ins):])
    elif name in _STATE_SETBAD:
        for o in placeholder[len(ins):]:
            o.badflag = True
    elif name in _STATE_SETGOOD:
        for o in placeholder[len(ins):]:
            o.badflag = False

    for target, o in temps:
        conv = convert_type(target, o.datatype)
        run_op("assgn", [conv], [o])
        o.badflag = o.badflag or target.badflag
    return final_outs


_COLL = {}
for _k in ("sum", "avg", "min", "max", "min_ind", "max_ind"):
    _c = OpSpec.__new__(OpSpec)
    _c.name, _c.kind, _c.opid = "coll_" + _k, "coll", _abi.OPS["coll_" + _k]
    _c.pars, _c.gentypes, _c.inplace, _c.fixed = [Par("rec", ("r", "k"), typed=T.LL), Par("b", out=True)], T.ALL, (), (("r", 4),)
    _COLL[_k] = _c


def collapse_records(kind: str, recs: PDL, value_type: int, bval: bool) -> PDL:
    """COLL_<kind> (include/pdlb200.h): recs = longlong [4, k, rows...], the k per-rank records of every row;
    returns the merged [rows...] ndarray (type: `value_type`, or indx for the _ind kinds).  ONE launch."""
    spec = _COLL[kind]
    if recs.datatype not in (T.LL, T.IND) or recs.ndims < 2 or recs.dims[0] != 4:
        raise PDLError(f"PDL::{spec.name}: records must be longlong [4, k, ...]")
    out = PDL.empty(T.IND if kind.endswith("_ind") else value_type, recs.dims[2:], recs.engine)
    out.badflag = bool(bval)
    bc = _broadcast([recs, out], [2, 0], [False, False], spec.name)
    named = {"ind": [recs.dims[1]], "rinc": [_real_inc(recs, 0), _real_inc(recs, 1)]}
    _launch(spec, value_type, [recs, out], bc, named, bval)
    return out


# ---- the three generator shapes of Ops.pd ----------------------------------------------------

# ops whose deferred form a following reduction knows how to absorb
_DEFERRABLE = ("mult",)


def _deferred(name: str, a: PDL, b: PDL) -> PDL:
    """make_trans_mutual with PDL_ITRANS_DO_DATAFLOW_F (pdlapi.c:781-801): types, dims and flags of the child
    are settled now, readdata is NOT run — the child keeps the transformation as `_pending`."""
    spec = SPECS[name]
    transtype = transtype_select(spec, [a, b, None])
    ins = [x if x.datatype == transtype else convert_type(x, transtype) for x in (a, b)]
    bc = _broadcast(ins + [ins[0]], [0, 0, 0], [False, False, True], name)
    out = PDL(a.engine, None, transtype, bc.dims[:bc.nimpl])
    out.badflag = any(x.badflag for x in ins)
    out._pending = (name, ins)
    return out


def run_biop(name: str, a, b, c=None, swap: int = 0) -> PDL:
    """XS PDL::<name>(a,b,[c],swap): swap a/b, then PDL_XS_INPLACE (lib/PDL/Ops.pd:108-114,
    lib/PDL/Core/pdlperl.h:61-71)."""
    engine = next((x.engine for x in (a, b, c) if isinstance(x, PDL)), None)
    a, b = as_pdl(a, engine), as_pdl(b, engine)
    if swap:
        a, b = b, a
    if c is None and a.is_inplace():
        a._inplace = False
        c = a
    if c is None and name in _DEFERRABLE and (a._flowing or b._flowing) and T.is_device_type(a.datatype) and T.is_device_type(b.datatype):
        return _deferred(name, a, b)
    return run_op(name, [a, b], [c])[0]


def fused_reduction(name: str, x: PDL):
    """`($a->flowing * $b)->sumover`: the product's readdata was deferred and the reduction is its consumer, so
    the pair runs as ONE launch of the fused kernel (`inner`, lib/PDL/Primitive.pd:48-70 = sum_n a*b) and the
    intermediate ndarray is never written.  Only where the fused form has the unfused one's semantics: no BAD
    values (inner makes a row with a BAD element BAD, sumover skips it), float/double (for the small integer
    types sumover widens to `int+`, inner does not).  Returns None when the pair does not qualify."""
    if name != "sumover" or x._pending is None or x._pending[0] != "mult":
        return None
    a, b = x._pending[1]
    if a.badflag or b.badflag or x.datatype not in (T.F, T.D) or x.ndims < 1:
        return None
    # the reduced dim is dim 0 of the PRODUCT: give both operands that dim explicitly (size 1 = stretches)
    a1 = a if a.ndims >= 1 else a.dummy(0)
    b1 = b if b.ndims >= 1 else b.dummy(0)
    return run_op("inner", [a1, b1], [None])[0]


def run_ufunc(name: str, a, b=None) -> PDL:
    a = as_pdl(a)
    if b is None and a.is_inplace():
        a._inplace = False
        b = a
    return run_op(name, [a], [b])[0]
