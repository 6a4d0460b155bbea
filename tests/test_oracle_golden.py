"""CPU: pins the C oracle (oracle/pdl_oracle.c) AND pdl_b200's host logic (type selection,
broadcast merging, output creation, bad-flag propagation) against fixtures recorded from the
real reference (tests/golden/make_golden.pl, PDL 2.106 built by oracle/build_ref.sh)."""
import pytest

from replay import check_case, load_cases

FILES = ["biop.json", "bifunc.json", "ufunc.json", "coerce.json", "broadcast.json", "bad.json",
         "reduce.json", "matmult.json", "badops.json", "basic.json", "inner.json", "minmax.json", "outer.json", "edge.json", "round2.json", "complex.json"]


def _params():
    for f in FILES:
        for c in load_cases(f):
            yield pytest.param(c, id=f"{f[:-5]}:{c['name']}")


@pytest.mark.parametrize("case", list(_params()))
def test_oracle_matches_reference(case, oracle_engine):
    check_case(case, oracle_engine)
