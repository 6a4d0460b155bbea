"""CPU: the coherent host/device store's state machine (pdl_b200/csrc/store.cu) — dirty bits, mprotect'd host
mirror, fault-driven lazy download, write detection, recycling, aliasing owners — exercised WITHOUT a GPU.

The product library has no CPU path; this test compiles store.cu a second time with -DPDLB200_STORE_HOSTSIM into
tests/_build/libstoresim.so, where "device" memory is malloc'd host memory, and drives it through the same C-ABI
(include/pdlb200.h pdlb200_mbuf_*).  Host reads/writes below are real CPU loads/stores into PROT_NONE / PROT_READ
pages: they fault into the store's SIGSEGV handler exactly as the unmodified reference core would."""
import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "pdl_b200" / "csrc" / "store.cu"
LIB = ROOT / "tests" / "_build" / "libstoresim.so"


@pytest.fixture(scope="module")
def sim():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    LIB.parent.mkdir(exist_ok=True)
    if not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        subprocess.run([nvcc, "-DPDLB200_STORE_HOSTSIM", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC",
                        "-shared", "-o", str(LIB), str(SRC)], check=True, capture_output=True)
    lib = C.CDLL(str(LIB))
    lib.pdlb200_mbuf_new.argtypes = [C.c_size_t]
    lib.pdlb200_mbuf_new.restype = C.c_void_p
    lib.pdlb200_mbuf_adopt.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.pdlb200_mbuf_adopt.restype = C.c_void_p
    lib.pdlb200_mbuf_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_size_t]
    lib.pdlb200_mbuf_dev.restype = C.c_void_p
    lib.pdlb200_mbuf_host.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]
    lib.pdlb200_mbuf_free.argtypes = [C.c_void_p]
    lib.pdlb200_mbuf_free.restype = None
    lib.pdlb200_mbuf_retain.argtypes = [C.c_void_p]
    lib.pdlb200_mbuf_retain.restype = None
    lib.pdlb200_mbuf_is.argtypes = [C.c_void_p]
    lib.pdlb200_mbuf_state.argtypes = [C.c_void_p]
    lib.pdlb200_mbuf_stats.argtypes = [C.POINTER(C.c_uint64)]
    lib.pdlb200_mbuf_stats.restype = None
    return lib


def view(ptr, n, dtype=np.uint8):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,)).view(dtype)


def stats(lib):
    s = (C.c_uint64 * 8)()
    lib.pdlb200_mbuf_stats(s)
    return dict(zip(("new", "recycled", "uploads", "upload_bytes", "downloads", "download_bytes", "faults", "adopted"), s))


def test_device_result_read_by_host_faults_once(sim):
    n = 1 << 20
    h = sim.pdlb200_mbuf_new(n)
    assert h and sim.pdlb200_mbuf_is(h) and sim.pdlb200_mbuf_state(h) == 0          # mirror stale, device copy undefined
    d = sim.pdlb200_mbuf_dev(h, 1, 1, None, None, 0)                                # a "kernel" writes the whole buffer
    view(d, n)[:] = np.arange(n, dtype=np.uint64).astype(np.uint8)
    assert sim.pdlb200_mbuf_state(h) == 4                                           # device current, mirror stale
    s0 = stats(sim)
    got = view(h, n).copy()                                                          # host READ of PROT_NONE pages -> fault -> download
    s1 = stats(sim)
    assert np.array_equal(got, np.arange(n, dtype=np.uint64).astype(np.uint8))
    assert s1["faults"] == s0["faults"] + 1 and s1["downloads"] == s0["downloads"] + 1 and s1["download_bytes"] - s0["download_bytes"] == n
    assert sim.pdlb200_mbuf_state(h) == 5                                           # both copies current (mirror read-only)
    view(h, n)[:10].sum()                                                            # reading again costs nothing
    assert stats(sim)["downloads"] == s1["downloads"]
    sim.pdlb200_mbuf_dev(h, 0, 0, None, None, 0)                                    # a device op that only READS it: no upload
    assert stats(sim)["uploads"] == s1["uploads"]
    sim.pdlb200_mbuf_free(h)


def test_host_write_is_detected_and_reuploaded(sim):
    n = 3 * 4096 + 17                                                               # not a page multiple
    src = np.arange(n, dtype=np.uint8)
    h = sim.pdlb200_mbuf_adopt(src.ctypes.data, n, None, 0)
    assert sim.pdlb200_mbuf_state(h) == 4
    hv = view(h, n)
    hv[5] = 200                                                                      # host WRITE into a stale mirror: download, then writable
    assert sim.pdlb200_mbuf_state(h) == 2                                           # host modified, device copy stale
    assert hv[4] == 4 and hv[5] == 200 and hv[n - 1] == src[n - 1]
    u0 = stats(sim)["uploads"]
    d = sim.pdlb200_mbuf_dev(h, 0, 0, None, None, 0)                                # next device read uploads the host copy
    assert stats(sim)["uploads"] == u0 + 1 and view(d, n)[5] == 200
    assert sim.pdlb200_mbuf_state(h) == 5                                           # mirror back to read-only: the next write faults again
    hv[6] = 201
    assert sim.pdlb200_mbuf_state(h) == 2
    d = sim.pdlb200_mbuf_dev(h, 1, 0, None, None, 0)                                # read-modify-write on the device (inplace op)
    assert view(d, n)[6] == 201
    view(d, n)[7] = 77
    assert sim.pdlb200_mbuf_state(h) == 4
    assert view(h, n)[7] == 77 and view(h, n)[6] == 201
    sim.pdlb200_mbuf_free(h)


def test_explicit_choke_point_and_discard(sim):
    n = 65536
    h = sim.pdlb200_mbuf_new(n)
    d = sim.pdlb200_mbuf_dev(h, 1, 1, None, None, 0)
    view(d, n)[:] = 9
    f0 = stats(sim)["faults"]
    assert sim.pdlb200_mbuf_host(h, 0, None, 0) == 0                                # explicit "make current": no fault needed afterwards
    assert view(h, n)[123] == 9 and stats(sim)["faults"] == f0
    assert sim.pdlb200_mbuf_host(h, 1, None, 0) == 0                                # ... for writing: device copy stale
    view(h, n)[0] = 1
    assert sim.pdlb200_mbuf_state(h) == 2 and stats(sim)["faults"] == f0
    u0 = stats(sim)["uploads"]
    d = sim.pdlb200_mbuf_dev(h, 1, 1, None, None, 0)                                # a kernel that overwrites everything: no upload of the stale copy
    assert stats(sim)["uploads"] == u0 and sim.pdlb200_mbuf_state(h) == 4
    # plain host pointers are "always current"
    buf = np.zeros(16, dtype=np.uint8)
    assert sim.pdlb200_mbuf_host(buf.ctypes.data, 1, None, 0) == 0 and not sim.pdlb200_mbuf_is(buf.ctypes.data)
    assert sim.pdlb200_mbuf_state(buf.ctypes.data) == -1
    sim.pdlb200_mbuf_free(h)


def test_recycling_and_shared_owners(sim):
    n = 1 << 16
    h = sim.pdlb200_mbuf_new(n)
    view(sim.pdlb200_mbuf_dev(h, 1, 1, None, None, 0), n)[:] = 5
    assert view(h, n)[0] == 5                                                        # populate the mirror
    sim.pdlb200_mbuf_retain(h)                                                       # a clump()ed child shares the buffer
    sim.pdlb200_mbuf_free(h)
    assert sim.pdlb200_mbuf_is(h)                                                    # one owner left
    sim.pdlb200_mbuf_free(h)
    assert not sim.pdlb200_mbuf_is(h)
    r0 = stats(sim)["recycled"]
    h2 = sim.pdlb200_mbuf_new(n)                                                     # same size: comes from the free list ...
    assert stats(sim)["recycled"] == r0 + 1
    assert sim.pdlb200_mbuf_state(h2) == 0                                          # ... protected again, contents undefined
    d0 = stats(sim)["downloads"]
    view(sim.pdlb200_mbuf_dev(h2, 1, 1, None, None, 0), n)[:] = 6
    assert view(h2, n)[100] == 6 and stats(sim)["downloads"] == d0 + 1
    sim.pdlb200_mbuf_free(h2)


def test_fault_in_a_worker_thread(sim):
    """CPU loops of the reference may run in pthreads (autopthread): a fault there is handled the same way."""
    import threading
    n = 1 << 18
    h = sim.pdlb200_mbuf_new(n)
    view(sim.pdlb200_mbuf_dev(h, 1, 1, None, None, 0), n)[:] = 3
    out = []
    ts = [threading.Thread(target=lambda k=k: out.append(int(view(h, n)[k::4].astype(np.int64).sum()))) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert sum(out) == 3 * n
    sim.pdlb200_mbuf_free(h)
