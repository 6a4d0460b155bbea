"""CPU: the committed golden fixtures ARE what the real reference produces — regenerate them with
tests/golden/make_golden.pl against the reference built in oracle/_ref (PDL 2.106) and compare byte for byte.
Skipped where the reference build is absent (it is git-ignored; oracle/build_ref.sh creates it)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref" / "blib"
GOLDEN = ROOT / "tests" / "golden"


def test_fixtures_regenerate_identically(tmp_path):
    if not shutil.which("perl") or not (REF / "lib" / "PDL" / "LiteF.pm").exists():
        pytest.skip("oracle/_ref (the built reference) is not present")
    r = subprocess.run(["perl", f"-I{REF / 'lib'}", f"-I{REF / 'arch'}", str(GOLDEN / "make_golden.pl"), str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    committed = sorted(p.name for p in GOLDEN.glob("*.json"))
    produced = sorted(p.name for p in tmp_path.glob("*.json"))
    assert committed == produced, (committed, produced)
    for name in committed:
        assert (GOLDEN / name).read_bytes() == (tmp_path / name).read_bytes(), f"{name} differs from what the reference produces now"
