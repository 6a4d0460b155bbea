"""ctypes image of include/pdlb200.h (the C-ABI of libpdlb200.so).

Only plain pointers and sizes cross this boundary: it is the same binding a
maintainer of the reference would write in XS, in ctypes.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

MAXDIMS = 16
MAXPDLS = 5
PAR_BADFLAG = 1
PAR_BADNAN = 2

OK, EINVAL, EUNSUPPORTED, ENODEVICE, ECUDA = range(5)


class Par(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("offs", C.c_int64),
        ("badval", C.c_uint64),
        ("type", C.c_int32),
        ("flags", C.c_int32),
    ]


class Trans(C.Structure):
    _fields_ = [
        ("op", C.c_int32),
        ("datatype", C.c_int32),
        ("bvalflag", C.c_int32),
        ("npdls", C.c_int32),
        ("ndims", C.c_int32),
        ("tflags", C.c_int32),
        ("dims", C.c_int64 * MAXDIMS),
        ("incs", C.c_int64 * (MAXDIMS * MAXPDLS)),
        ("ind", C.c_int64 * 4),
        ("rinc", C.c_int64 * 8),
        ("pdls", Par * MAXPDLS),
        ("stream", C.c_void_p),
        ("param", C.c_double),
        ("anybad", C.POINTER(C.c_int32)),
    ]


# op ids (enum in pdlb200.h)
OPS = {
    "plus": 0, "mult": 1, "minus": 2, "divide": 3,
    "gt": 4, "lt": 5, "le": 6, "ge": 7, "eq": 8, "ne": 9,
    "shiftleft": 10, "shiftright": 11, "or2": 12, "and2": 13, "xor": 14,
    "power": 15, "atan2": 16, "modulo": 17, "spaceship": 18,
    "bitnot": 19, "sqrt": 20, "sin": 21, "cos": 22, "not": 23,
    "exp": 24, "log": 25, "log10": 26, "_rabs": 27, "assgn": 28, "abs2": 29,
    "sumover": 30, "prodover": 31, "dsumover": 32, "dprodover": 33,
    "average": 34, "daverage": 35, "minimum": 36, "maximum": 37,
    "minimum_ind": 38, "maximum_ind": 39,
    "andover": 40, "orover": 41, "bandover": 42, "borover": 43,
    "zcover": 44, "xorover": 45, "bxorover": 46, "nbadover": 47, "ngoodover": 48,
    "cumusumover": 50, "cumuprodover": 51, "dcumusumover": 52, "dcumuprodover": 53,
    "matmult": 60, "converttype": 61, "ipow": 62,
    "isbad": 63, "isgood": 64, "isnan": 65, "setbadif": 66, "setvaltobad": 67,
    "setnantobad": 68, "setinftobad": 69, "setnonfinitetobad": 70, "setbadtonan": 71,
    "setbadtoval": 72, "badmask": 73, "copybad": 74, "axisvalues": 75, "inner": 76, "minmaximum": 77, "magnover": 78, "outer": 79,
    # sharded whole-array reductions: per-rank partial records and their rank-ordered merge
    "part_sum": 80, "part_dsum": 81, "part_min": 82, "part_max": 83,
    "coll_sum": 84, "coll_avg": 85, "coll_min": 86, "coll_max": 87, "coll_min_ind": 88, "coll_max_ind": 89,
    "minimum_n_ind": 90, "maximum_n_ind": 91,
}
ABI_VERSION = 4
TRANS_DEFER_ANYBAD = 1     # pdlb200_trans.tflags

# every symbol include/pdlb200.h declares (tests check the .so exports them all)
SYMBOLS = [
    "pdlb200_readdata", "pdlb200_elementwise", "pdlb200_reduce", "pdlb200_matmult",
    "pdlb200_buf_new", "pdlb200_buf_free", "pdlb200_buf_nbytes", "pdlb200_buf_devptr",
    "pdlb200_buf_upload", "pdlb200_buf_download", "pdlb200_buf_device_dirty",
    "pdlb200_abi_version", "pdlb200_build_id", "pdlb200_device_count", "pdlb200_set_device", "pdlb200_sm_count",
    "pdlb200_sync", "pdlb200_host_alloc", "pdlb200_host_free", "pdlb200_memcpy_h2d",
    "pdlb200_memcpy_d2h", "pdlb200_managed_alloc", "pdlb200_managed_free", "pdlb200_managed_trim", "pdlb200_ptr_kind",
    "pdlb200_prefetch", "pdlb200_launch_count", "pdlb200_last_kernel", "pdlb200_op_name",
    "pdlb200_type_size",
    "pdlb200_mbuf_new", "pdlb200_mbuf_adopt", "pdlb200_mbuf_retain", "pdlb200_mbuf_free", "pdlb200_mbuf_is",
    "pdlb200_mbuf_dev", "pdlb200_mbuf_host", "pdlb200_mbuf_state", "pdlb200_mbuf_stats", "pdlb200_mbuf_trim",
    "pdlb200_devop_register", "pdlb200_devop_is", "pdlb200_dev_alloc", "pdlb200_dev_free", "pdlb200_dev_trim",
    "pdlb200_host_alloc_wc",
    "pdlb200_peer_mailbox_new", "pdlb200_peer_mailbox_free", "pdlb200_ipc_export", "pdlb200_ipc_open", "pdlb200_ipc_close",
    "pdlb200_peer_exchange", "pdlb200_peer_gathered_offset",
]

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libpdlb200.so"

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """dlopen libpdlb200.so and type its entry points.  Fails loudly when the CUDA
    extension has not been built: there is no CPU fallback behind this package."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise LibraryMissing(
            f"{LIB_PATH} is missing: build it with `python -m pdl_b200.build` "
            "(pdl_b200 has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    # a prebuilt library must match the sources it ships with (it travels to the GPU box as a file): refuse a stale one
    try:
        from .build import source_id, CSRC
        want = source_id() if CSRC.is_dir() else None
    except Exception:  # a source-less install: nothing to compare against
        want = None
    if want is not None and os.environ.get("PDLB200_SKIP_BUILD_ID") != "1":
        lib.pdlb200_build_id.restype = C.c_char_p
        have = (lib.pdlb200_build_id() or b"").decode()
        if have != want:
            raise LibraryMissing(f"{LIB_PATH} was built from other sources (library {have}, sources {want}): "
                                 "rebuild it with `python -m pdl_b200.build`")
    errargs = [C.c_char_p, C.c_size_t]
    for name in ("pdlb200_readdata", "pdlb200_elementwise", "pdlb200_reduce", "pdlb200_matmult"):
        f = getattr(lib, name)
        f.argtypes = [C.POINTER(Trans)] + errargs
        f.restype = C.c_int
    lib.pdlb200_buf_new.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)] + errargs
    lib.pdlb200_buf_new.restype = C.c_int
    lib.pdlb200_buf_free.argtypes = [C.c_void_p]
    lib.pdlb200_buf_free.restype = None
    lib.pdlb200_buf_nbytes.argtypes = [C.c_void_p]
    lib.pdlb200_buf_nbytes.restype = C.c_size_t
    lib.pdlb200_buf_devptr.argtypes = [C.c_void_p, C.c_int]
    lib.pdlb200_buf_devptr.restype = C.c_void_p
    lib.pdlb200_buf_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p] + errargs
    lib.pdlb200_buf_upload.restype = C.c_int
    lib.pdlb200_buf_download.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p] + errargs
    lib.pdlb200_buf_download.restype = C.c_int
    lib.pdlb200_buf_device_dirty.argtypes = [C.c_void_p]
    lib.pdlb200_buf_device_dirty.restype = C.c_int
    lib.pdlb200_abi_version.restype = C.c_int
    lib.pdlb200_build_id.restype = C.c_char_p
    lib.pdlb200_device_count.restype = C.c_int
    lib.pdlb200_set_device.argtypes = [C.c_int] + errargs
    lib.pdlb200_set_device.restype = C.c_int
    lib.pdlb200_sm_count.restype = C.c_int
    lib.pdlb200_sync.argtypes = [C.c_void_p] + errargs
    lib.pdlb200_sync.restype = C.c_int
    lib.pdlb200_host_alloc.argtypes = [C.c_size_t]
    lib.pdlb200_host_alloc.restype = C.c_void_p
    lib.pdlb200_host_alloc_wc.argtypes = [C.c_size_t]
    lib.pdlb200_host_alloc_wc.restype = C.c_void_p
    lib.pdlb200_host_free.argtypes = [C.c_void_p]
    lib.pdlb200_host_free.restype = None
    for name in ("pdlb200_memcpy_h2d", "pdlb200_memcpy_d2h"):
        f = getattr(lib, name)
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p] + errargs
        f.restype = C.c_int
    lib.pdlb200_managed_alloc.argtypes = [C.c_size_t]
    lib.pdlb200_managed_alloc.restype = C.c_void_p
    lib.pdlb200_managed_free.argtypes = [C.c_void_p]
    lib.pdlb200_managed_free.restype = None
    lib.pdlb200_ptr_kind.argtypes = [C.c_void_p]
    lib.pdlb200_ptr_kind.restype = C.c_int
    lib.pdlb200_prefetch.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p] + errargs
    lib.pdlb200_prefetch.restype = C.c_int
    lib.pdlb200_launch_count.restype = C.c_uint64
    lib.pdlb200_last_kernel.restype = C.c_char_p
    lib.pdlb200_op_name.argtypes = [C.c_int]
    lib.pdlb200_op_name.restype = C.c_char_p
    lib.pdlb200_type_size.argtypes = [C.c_int]
    lib.pdlb200_type_size.restype = C.c_size_t
    lib.pdlb200_dev_alloc.argtypes = [C.c_size_t]
    lib.pdlb200_dev_alloc.restype = C.c_void_p
    lib.pdlb200_dev_free.argtypes = [C.c_void_p, C.c_size_t]
    lib.pdlb200_dev_free.restype = None
    lib.pdlb200_dev_trim.restype = None
    lib.pdlb200_peer_mailbox_new.argtypes = [C.c_int, C.c_size_t]
    lib.pdlb200_peer_mailbox_new.restype = C.c_void_p
    lib.pdlb200_peer_mailbox_free.argtypes = [C.c_void_p]
    lib.pdlb200_peer_mailbox_free.restype = None
    lib.pdlb200_ipc_export.argtypes = [C.c_void_p, C.c_char_p] + errargs
    lib.pdlb200_ipc_export.restype = C.c_int
    lib.pdlb200_ipc_open.argtypes = [C.c_char_p] + errargs
    lib.pdlb200_ipc_open.restype = C.c_void_p
    lib.pdlb200_ipc_close.argtypes = [C.c_void_p]
    lib.pdlb200_ipc_close.restype = None
    lib.pdlb200_peer_exchange.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p] + errargs
    lib.pdlb200_peer_exchange.restype = C.c_int
    lib.pdlb200_peer_gathered_offset.argtypes = [C.c_int, C.c_size_t, C.c_int64]
    lib.pdlb200_peer_gathered_offset.restype = C.c_int64
    lib.pdlb200_mbuf_new.argtypes = [C.c_size_t]
    lib.pdlb200_mbuf_new.restype = C.c_void_p
    lib.pdlb200_mbuf_adopt.argtypes = [C.c_void_p, C.c_size_t] + errargs
    lib.pdlb200_mbuf_adopt.restype = C.c_void_p
    for name in ("pdlb200_mbuf_retain", "pdlb200_mbuf_free"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = None
    lib.pdlb200_mbuf_is.argtypes = [C.c_void_p]
    lib.pdlb200_mbuf_is.restype = C.c_int
    lib.pdlb200_mbuf_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p] + errargs
    lib.pdlb200_mbuf_dev.restype = C.c_void_p
    lib.pdlb200_mbuf_host.argtypes = [C.c_void_p, C.c_int] + errargs
    lib.pdlb200_mbuf_host.restype = C.c_int
    lib.pdlb200_mbuf_state.argtypes = [C.c_void_p]
    lib.pdlb200_mbuf_state.restype = C.c_int
    lib.pdlb200_mbuf_stats.argtypes = [C.POINTER(C.c_uint64)]
    lib.pdlb200_mbuf_stats.restype = None
    lib.pdlb200_mbuf_trim.restype = None
    _lib = lib
    return lib
