"""CPU: the C-ABI library loads, exports every symbol include/pdlb200.h declares, and its
compute entry points FAIL LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from pdl_b200 import _abi

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_all_exported():
    hdr = (ROOT / "include" / "pdlb200.h").read_text()
    declared = set(re.findall(r"PDLB200_API[^;]*?\b(pdlb200_\w+)\s*\(", hdr))
    assert declared, "no PDLB200_API prototypes found"
    assert declared == set(_abi.SYMBOLS)
    lib = _abi.load()
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layout_matches_header():
    # sizes computed from the header's field list (LP64)
    assert C.sizeof(_abi.Par) == 8 + 8 + 8 + 4 + 4
    assert C.sizeof(_abi.Trans) == 6 * 4 + 16 * 8 + 80 * 8 + 4 * 8 + 8 * 8 + 5 * C.sizeof(_abi.Par) + 8 + 8 + 8


def test_field_offsets_match_a_c_compiler(tmp_path):
    """offsetof() of every descriptor field as gcc lays out include/pdlb200.h against the ctypes mirror (_abi.py):
    a renamed or re-ordered field shows up here, not as a wrong answer on the device."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    fields = {"pdlb200_trans": [n for n, _ in _abi.Trans._fields_], "pdlb200_par": [n for n, _ in _abi.Par._fields_]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "pdlb200.h"', 'int main(void) {']
    for st, names in fields.items():
        for n in names:
            lines.append(f'  printf("{st}.{n} %zu\\n", offsetof({st}, {n}));')
        lines.append(f'  printf("{st}.sizeof %zu\\n", sizeof({st}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "offsets.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "offsets"
    root = __import__("pathlib").Path(__file__).resolve().parent.parent
    subprocess.run(["gcc", "-I", str(root / "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for st, cls in (("pdlb200_trans", _abi.Trans), ("pdlb200_par", _abi.Par)):
        for n, _ in cls._fields_:
            assert int(got[f"{st}.{n}"]) == getattr(cls, n).offset, (st, n)
        assert int(got[f"{st}.sizeof"]) == C.sizeof(cls), st


def test_plumbing_without_gpu():
    lib = _abi.load()
    assert lib.pdlb200_abi_version() == _abi.ABI_VERSION == 4
    assert lib.pdlb200_type_size(10) == 8 and lib.pdlb200_type_size(0) == 1
    assert lib.pdlb200_op_name(30) == b"sumover"
    if lib.pdlb200_device_count() > 0:
        pytest.skip("a GPU is present; the no-device behaviour is checked on CPU boxes")
    err = C.create_string_buffer(256)
    tr = _abi.Trans()
    tr.op, tr.datatype, tr.npdls, tr.ndims = 0, 10, 3, 1
    rc = lib.pdlb200_readdata(C.byref(tr), err, 256)
    assert rc == _abi.ENODEVICE
    assert b"no CPU fallback" in err.value


def test_engine_refuses_without_gpu():
    import pdl_b200 as P
    lib = _abi.load()
    if lib.pdlb200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(P.PDLError):
        P.CudaEngine()


def test_loader_refuses_a_library_built_from_other_sources(monkeypatch):
    """The prebuilt .so travels to the GPU box as a file: the loader compares the id compiled into it
    (pdlb200_build_id) with the hash of the sources next to it and refuses a mismatch instead of testing a stale build."""
    from pdl_b200 import build
    lib = _abi.load()
    assert lib.pdlb200_build_id().decode() == build.source_id()
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(build, "source_id", lambda: "0" * 16)
    with pytest.raises(_abi.LibraryMissing, match="built from other sources"):
        _abi.load()
    monkeypatch.setenv("PDLB200_SKIP_BUILD_ID", "1")
    assert _abi.load() is not None
