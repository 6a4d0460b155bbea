"""GPU: the drop-in boundary itself.  PDL::B200 (perl/PDL-B200) swaps libpdlb200 into the
vtables of the UNMODIFIED reference built in oracle/_ref; then
  (a) perl/PDL-B200/t/parity.t runs every op twice in one process (device vs the reference's
      own CPU readdata) and compares bytes;
  (b) the reference's OWN test files for this path (copied into the git-ignored oracle/_ref/t
      by oracle/build_ref.sh) are run with the shim attached."""
import os
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
SHIM = ROOT / "perl" / "PDL-B200" / "blib"
pytestmark = pytest.mark.gpu


def _perl(args, timeout=900):
    inc = [f"-I{REF / 'blib' / 'lib'}", f"-I{REF / 'blib' / 'arch'}", f"-I{SHIM / 'lib'}", f"-I{SHIM / 'arch'}",
           f"-I{ROOT / 'oracle' / 'shim'}"]
    return subprocess.run(["perl", *inc, *args], capture_output=True, text=True, timeout=timeout, cwd=str(REF))


def _need():
    if shutil.which("perl") is None or not (REF / "blib").exists():
        pytest.skip("oracle/_ref (the built reference) is not present on this box")
    if not (SHIM / "arch" / "auto" / "PDL" / "B200" / "B200.so").exists():
        pytest.skip("perl/PDL-B200 shim not built (perl/PDL-B200/build.sh)")


def test_same_process_parity():
    _need()
    r = _perl([str(ROOT / "perl" / "PDL-B200" / "t" / "parity.t")])
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "not ok" not in r.stdout, tail


@pytest.mark.parametrize("tfile", ["ops.t", "ops-bitwise.t", "ufunc.t", "bad.t", "primitive-matmult.t", "thread.t", "slice.t",
                                   "core.t", "clump.t", "reduce.t"])
def test_reference_own_tests_with_shim(tfile):
    _need()
    if not (REF / "t" / tfile).exists():
        pytest.skip(f"oracle/_ref/t/{tfile} not present")
    r = _perl(["-MPDL::B200", str(REF / "t" / tfile)])
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    failed = [l for l in r.stdout.splitlines() if l.startswith("not ok") and "# TODO" not in l]
    assert not failed, failed[:10]
