"""Execution engine behind pdl_b200.PDL: where ndarray bytes live and who runs a
transformation's readdata.

The product has exactly ONE engine, `CudaEngine`, which talks to libpdlb200.so
through the C-ABI in include/pdlb200.h.  There is no CPU engine in this package
and no fallback: constructing a CudaEngine without the built library or without
a GPU raises.  (tests/ plug the C oracle in through the same small interface to
check the CUDA results; that lives under oracle/, never here.)
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from . import _abi


class PDLError(RuntimeError):
    """What the reference reports through barf / pdl_error (lib/PDL/Core/pdlutil.c:507-529)."""


class Store:
    """One device allocation (the reference's pdl.datasv / pdl.data, pdlapi.c:172-209).  `free` (set by the engine
    that allocated it) hands the block back when the last ndarray that views it goes away."""

    __slots__ = ("engine", "handle", "ptr", "nbytes", "_keep", "_free", "_bad", "_pend", "__weakref__")

    def __init__(self, engine, handle, ptr, nbytes, keep=None, free=None):
        self.engine, self.handle, self.ptr, self.nbytes, self._keep, self._free = engine, handle, ptr, nbytes, keep, free
        self._bad = False     # the bad state of the DATA: shared by every view of this buffer (pdl_propagate_badflag_dir)
        self._pend = None     # slot of a data-dependent flag still in flight (FlagRing), OR-ed in when first asked for

    @property
    def bad(self) -> bool:
        if self._pend is not None:
            self.engine.flag_ring.resolve()
        return self._bad

    @bad.setter
    def bad(self, v: bool) -> None:
        self._bad, self._pend = bool(v), None

    def __del__(self):
        f = self._free
        if f is not None:
            try:
                f(self.ptr, self.nbytes)
            except Exception:  # interpreter shutdown: the library may already be gone
                pass


class FlagRing:
    """Data-dependent output badflags without a host round trip per op (`if (flag) $PDLSTATESETBAD(...)`,
    lib/PDL/Bad.pd:695-707, Ufunc.pd:578-583): the op copies its flag word into one of 1024 pinned int32 slots
    (PDLB200_TRANS_DEFER_ANYBAD) and the output Stores remember the slot; the first question about their bad
    state synchronises the stream once and settles every flag in flight."""

    SLOTS = 1024

    def __init__(self, engine):
        import weakref
        self._ref = weakref.ref
        self.engine = engine
        self.page = engine.lib.pdlb200_host_alloc(4 * self.SLOTS)
        if not self.page:
            raise PDLError("pdl_b200: cannot allocate the pinned flag page")
        self.arr = (C.c_int32 * self.SLOTS).from_address(self.page)
        self.next, self.pending = 0, {}

    def take(self):
        """-> (slot, POINTER(c_int32)) for the next deferred call."""
        slot = self.next
        self.next = (slot + 1) % self.SLOTS
        if slot in self.pending:
            self.resolve()
        self.arr[slot] = 0
        return slot, C.cast(self.page + 4 * slot, C.POINTER(C.c_int32))

    def attach(self, slot, stores) -> None:
        for s in stores:
            s._pend = slot
        self.pending[slot] = [self._ref(s) for s in stores]

    def resolve(self) -> None:
        self.engine.sync()
        pending, self.pending = self.pending, {}
        for slot, refs in pending.items():
            v = self.arr[slot]
            for r in refs:
                s = r()
                if s is not None and s._pend == slot:
                    s._pend = None
                    if v:
                        s._bad = True


class Engine:
    """Interface a PDL object needs from its engine."""

    name = "abstract"

    def alloc(self, nbytes: int) -> Store:
        raise NotImplementedError

    def upload(self, store: Store, host: np.ndarray) -> None:
        raise NotImplementedError

    def download(self, store: Store, nbytes: int) -> np.ndarray:
        raise NotImplementedError

    def readdata(self, trans: _abi.Trans) -> None:
        raise NotImplementedError

    def sync(self) -> None:
        pass


class CudaEngine(Engine):
    """libpdlb200.so on one B200.  One process drives one GPU (torchrun sets LOCAL_RANK)."""

    name = "cuda"

    def __init__(self, device: int | None = None):
        self.lib = _abi.load()  # raises LibraryMissing if the extension is not built
        if self.lib.pdlb200_abi_version() != _abi.ABI_VERSION:
            raise PDLError("libpdlb200.so ABI version mismatch")
        n = self.lib.pdlb200_device_count()
        if n <= 0:
            raise PDLError("pdl_b200: no CUDA device is visible and there is no CPU fallback")
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) % n
        self.device = device
        self._err = C.create_string_buffer(512)
        self._check(self.lib.pdlb200_set_device(device, self._err, 512))
        self.stream = None  # legacy default stream: ordered with torch's default stream
        self._dev_alloc, self._dev_free = self.lib.pdlb200_dev_alloc, self.lib.pdlb200_dev_free
        self._flag_ring = None

    @property
    def flag_ring(self) -> FlagRing:
        if self._flag_ring is None:
            self._flag_ring = FlagRing(self)
        return self._flag_ring

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise PDLError(self._err.value.decode("utf-8", "replace"))

    def alloc(self, nbytes: int) -> Store:
        """pdl_allocdata for the device (pdlapi.c:172-209): one C call; recycled blocks come from the exact-size
        free list in front of the stream-ordered pool, and nothing is zero-filled."""
        ptr = self._dev_alloc(nbytes)
        if not ptr:
            raise PDLError(f"pdl_b200: cannot allocate {nbytes} bytes on the device")
        return Store(self, None, ptr, nbytes, None, self._dev_free)

    def wrap(self, ptr: int, nbytes: int, keep) -> Store:
        """Adopt foreign device memory (e.g. a torch tensor kept alive by `keep`)."""
        return Store(self, None, ptr, nbytes, keep)

    def upload(self, store: Store, host: np.ndarray) -> None:
        host = np.ascontiguousarray(host)
        if host.nbytes > store.nbytes:
            raise PDLError("upload larger than the allocation")
        if host.nbytes:
            self._check(self.lib.pdlb200_memcpy_h2d(store.ptr, host.ctypes.data, host.nbytes, self.stream, self._err, 512))
            self.sync()  # pageable source: make the call's completion explicit

    def upload_ptr(self, store: Store, host_ptr: int, nbytes: int, offset: int = 0) -> None:
        """Asynchronous H2D from (pinned) host memory into store[offset:offset+nbytes]."""
        self._check(self.lib.pdlb200_memcpy_h2d(store.ptr + offset, host_ptr, nbytes, self.stream, self._err, 512))

    def download(self, store: Store, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        if nbytes:
            self._check(self.lib.pdlb200_memcpy_d2h(out.ctypes.data, store.ptr, nbytes, self.stream, self._err, 512))
        self.sync()
        return out

    def download_ptr(self, store: Store, host_ptr: int, nbytes: int) -> None:
        self._check(self.lib.pdlb200_memcpy_d2h(host_ptr, store.ptr, nbytes, self.stream, self._err, 512))

    def readdata(self, trans: _abi.Trans) -> None:
        trans.stream = self.stream
        self._check(self.lib.pdlb200_readdata(C.byref(trans), self._err, 512))

    def sync(self) -> None:
        self._check(self.lib.pdlb200_sync(self.stream, self._err, 512))

    def launch_count(self) -> int:
        return int(self.lib.pdlb200_launch_count())

    def last_kernel(self) -> str:
        return (self.lib.pdlb200_last_kernel() or b"").decode()


_default: Engine | None = None


def default_engine() -> Engine:
    """The process-wide CudaEngine, created on first use.  Raises without GPU/library."""
    global _default
    if _default is None:
        _default = CudaEngine()
    return _default


def set_default_engine(e: Engine | None) -> None:
    global _default
    _default = e
