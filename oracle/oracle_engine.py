"""TEST INFRASTRUCTURE — an Engine that keeps ndarray bytes in host memory and runs
readdata through the C oracle (oracle/liboracle.so, built from oracle/pdl_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this.  It plugs into pdl_b200's Engine interface so that the SAME
host logic (type selection, broadcast merging, output creation) drives both the
CUDA path and the checker, and the two results can be compared byte for byte.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))

from pdl_b200 import _abi  # noqa: E402
from pdl_b200.engine import Engine, PDLError, Store  # noqa: E402

LIB = HERE / "liboracle.so"


def build_oracle(force: bool = False) -> Path:
    src = HERE / "pdl_oracle.c"
    hdr = HERE.parent / "include" / "pdlb200.h"
    if force or not LIB.exists() or LIB.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s", "liboracle.so"], check=True,
                       capture_output=True)
    return LIB


class OracleEngine(Engine):
    name = "oracle"

    def __init__(self):
        build_oracle()
        self.lib = C.CDLL(str(LIB))
        self.lib.pdl_oracle_readdata.argtypes = [C.POINTER(_abi.Trans), C.c_char_p, C.c_size_t]
        self.lib.pdl_oracle_readdata.restype = C.c_int
        self._err = C.create_string_buffer(512)
        self.calls = 0

    def alloc(self, nbytes: int) -> Store:
        buf = np.zeros(max(nbytes, 1) + 64, dtype=np.uint8)
        # 64-byte aligned start so alignment-dependent code paths see the same thing as on the device
        off = (-buf.ctypes.data) % 64
        view = buf[off:off + max(nbytes, 1)]
        return Store(self, None, view.ctypes.data, nbytes, keep=(buf, view))

    def upload(self, store: Store, host: np.ndarray) -> None:
        host = np.ascontiguousarray(host).reshape(-1).view(np.uint8)
        store._keep[1][: host.nbytes] = host

    def download(self, store, nbytes: int) -> np.ndarray:
        if nbytes == 0:
            return np.empty(0, dtype=np.uint8)
        return np.ctypeslib.as_array(C.cast(store.ptr, C.POINTER(C.c_uint8)), shape=(nbytes,)).copy()

    def readdata(self, trans: _abi.Trans) -> None:
        self.calls += 1
        rc = self.lib.pdl_oracle_readdata(C.byref(trans), self._err, 512)
        if rc != 0:
            raise PDLError(self._err.value.decode("utf-8", "replace"))
