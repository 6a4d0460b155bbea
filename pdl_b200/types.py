"""PDL's type table, as the device path sees it.

Mirrors lib/PDL/Types.pm:27-255 of the reference: the ORDER is the promotion order
(`pdl_datatypes` enum), default BAD values are type-min for signed/float and
type-max for unsigned (Types.pm `defbval`).  LD / CLD (x87 80-bit) and the complex
types are known by number so that type selection can be restated exactly, but
have no device representation.
"""
from __future__ import annotations

import struct

import numpy as np

SB, B, S, US, L, UL, IND, ULL, LL, F, D, LD, CF, CD, CLD = range(15)
NTYPES_DEVICE = 11

NAMES = ["sbyte", "byte", "short", "ushort", "long", "ulong", "indx", "ulonglong",
         "longlong", "float", "double", "ldouble", "cfloat", "cdouble", "cldouble"]
PPSYM = ["A", "B", "S", "U", "L", "K", "N", "P", "Q", "F", "D", "E", "G", "C", "H"]

NP_DTYPE = {
    SB: np.dtype(np.int8), B: np.dtype(np.uint8), S: np.dtype(np.int16), US: np.dtype(np.uint16),
    L: np.dtype(np.int32), UL: np.dtype(np.uint32), IND: np.dtype(np.int64), ULL: np.dtype(np.uint64),
    LL: np.dtype(np.int64), F: np.dtype(np.float32), D: np.dtype(np.float64),
    CF: np.dtype(np.complex64), CD: np.dtype(np.complex128),     # on the device path for + - * / only
}
SIZE = {t: d.itemsize for t, d in NP_DTYPE.items()}

INTEGER = (SB, B, S, US, L, UL, IND, ULL, LL)
UNSIGNED = (B, US, UL, ULL)
SIGNED_INT = (SB, S, L, IND, LL)       # PDL_TYPELIST_SIGNED order used by _pdl_whichdatatype_int
REAL = tuple(range(SB, LD + 1))        # ppdefs: default GenericTypes (PP.pm:1228)
ALL = tuple(range(SB, CLD + 1))        # ppdefs_all
FLOATING = (F, LD, D)                  # Ops.pd:8-9  $F, D last "so defaults to D"
COMPLEX = (CF, CD, CLD)

_FLT_MAX = float(np.finfo(np.float32).max)
_DBL_MAX = float(np.finfo(np.float64).max)
DEFAULT_BAD = {
    SB: -128, B: 255, S: -32768, US: 65535, L: -2**31, UL: 2**32 - 1,
    IND: -2**63, ULL: 2**64 - 1, LL: -2**63, F: -_FLT_MAX, D: -_DBL_MAX,
    CF: complex(-_FLT_MAX, -_FLT_MAX), CD: complex(-_DBL_MAX, -_DBL_MAX),      # Types.pm:209-232 defbval
}


def type_name(t: int) -> str:
    return NAMES[t]


def is_device_type(t: int) -> bool:
    return 0 <= t < NTYPES_DEVICE


COMPLEX_DEVICE_OPS = ("plus", "minus", "mult", "divide")


def is_device_type_for(t: int, opname: str) -> bool:
    """The device type matrix: the 11 real types for every op, complex float/double for + - * / only."""
    return is_device_type(t) or (t in (CF, CD) and opname in COMPLEX_DEVICE_OPS)


def from_numpy_dtype(dt) -> int:
    dt = np.dtype(dt)
    for t in (SB, B, S, US, L, UL, LL, ULL, F, D, CF, CD):  # int64 maps to longlong (PDL's `longlong`); indx on request
        if NP_DTYPE[t] == dt:
            return t
    if dt == np.dtype(bool):
        return B
    raise TypeError(f"numpy dtype {dt} has no PDL type on the device path")


def value_bits(t: int, value) -> int:
    """Bit pattern (low-order bytes, little endian) of `value` stored as type t."""
    dt = NP_DTYPE[t]
    if t in (CF, CD):
        # CF: both parts (re in the low 4 bytes); CD: the REAL part's bits, the imaginary part must be equal
        v = complex(value)
        if t == CD and v.imag != v.real and not (v.imag != v.imag and v.real != v.real):
            raise ValueError("a cdouble badvalue needs equal real and imaginary parts on the device path")
        raw = np.array([v]).astype(dt).tobytes()[:8]
        return int.from_bytes(raw, "little")
    with np.errstate(over="ignore"):
        arr = np.array([value]).astype(dt) if not isinstance(value, float) or t in (F, D) else np.array([int(value)]).astype(dt)
    raw = arr.tobytes()
    return int.from_bytes(raw.ljust(8, b"\0"), "little")


def bits_value(t: int, bits: int):
    dt = NP_DTYPE[t]
    raw = int(bits).to_bytes(8, "little")[: dt.itemsize]
    return np.frombuffer(raw, dtype=dt)[0]


def int_plus(t: int) -> int:
    """`int+` output type: max(long, t) in type order (PP/PdlParObj.pm:149-159, pdlapi.c:1174-1181)."""
    return max(L, t)


def whichdatatype_int(v: int) -> int:
    """Smallest signed type holding Perl IV v (lib/PDL/Core/pdlperl.h:193-198)."""
    for t in SIGNED_INT:
        info = np.iinfo(NP_DTYPE[t])
        if info.min <= v <= info.max:
            return t
    raise OverflowError(f"{v} cannot be converted by whichdatatype")


def whichdatatype_uint(v: int) -> int:
    """Smallest unsigned type holding Perl UV v (lib/PDL/Core/pdlperl.h:187-192)."""
    for t in UNSIGNED:
        if 0 <= v <= np.iinfo(NP_DTYPE[t]).max:
            return t
    raise OverflowError(f"{v} cannot be converted by whichdatatype")


def scalar_type(v) -> int:
    """Type PDL gives a bare Perl scalar: IV -> smallest signed int, UV beyond IV_MAX ->
    unsigned, NV -> double (lib/PDL/Core/pdlcore.c:64-78, pdlperl.h:185-205)."""
    if isinstance(v, (bool, np.bool_)):
        return whichdatatype_int(int(v))
    if isinstance(v, (int, np.integer)):
        v = int(v)
        if v > 2**63 - 1:
            return whichdatatype_uint(v)
        return whichdatatype_int(v)
    if isinstance(v, (float, np.floating)):
        return D
    raise TypeError(f"cannot make an ndarray from scalar {v!r}")


def pack_f32(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]
