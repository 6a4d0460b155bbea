"""GPU: every fixture recorded from the real reference, replayed through the C-ABI on the
CUDA path.  Bit-exact except the transcendental ops, whose tolerance (in ulp) is stored in
the fixture (glibc libm vs CUDA libdevice)."""
import pytest

from replay import check_case, load_cases

FILES = ["biop.json", "bifunc.json", "ufunc.json", "coerce.json", "broadcast.json", "bad.json",
         "reduce.json", "matmult.json", "badops.json", "basic.json", "inner.json", "minmax.json", "outer.json", "edge.json", "round2.json", "complex.json"]


def _params():
    for f in FILES:
        for c in load_cases(f):
            yield pytest.param(c, id=f"{f[:-5]}:{c['name']}")


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(_params()))
def test_cuda_matches_reference(case, cuda_engine):
    check_case(case, cuda_engine)
