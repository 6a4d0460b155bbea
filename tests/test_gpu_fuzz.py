"""GPU: seeded random differential test — random op x type x shape x view chain x BAD pattern through the CUDA
path and the oracle, compared byte for byte (float sums/products/averages: exact-by-construction values, so the
comparison stays bitwise apart from NaN payloads).  Deterministic: the seeds are fixed."""
import numpy as np
import pytest

import pdl_b200 as P
from pdl_b200 import types as T, ufunc, bad as B
from parity import ALL_TYPES, assert_same, both, rand_array

pytestmark = pytest.mark.gpu

BIOPS = ["plus", "minus", "mult", "gt", "le", "eq", "ne", "spaceship"]
UNARY = ["_rabs", "not", "assgn", "abs2"]
REDUCE = ["sumover", "average", "minimum", "maximum", "minimum_ind", "maximum_ind", "orover", "andover",
          "zcover", "xorover", "nbadover", "ngoodover", "cumusumover"]
BADOPS = ["isbad", "isgood", "setbadtoval", "setvaltobad"]


def random_view(rng, p):
    """A random chain of slice / xchg / dummy views of a 2-D ndarray."""
    for _ in range(int(rng.integers(0, 3))):
        k = int(rng.integers(0, 4))
        if k == 0 and p.ndims >= 2:
            p = p.xchg(0, 1)
        elif k == 1:
            specs = []
            for d in p.dims:
                if d < 4 or rng.random() < 0.3:
                    specs.append(":")
                else:
                    lo = int(rng.integers(0, d // 2)); hi = int(rng.integers(d // 2, d)); st = int(rng.integers(1, 4))
                    specs.append(f"{hi}:{lo}:-{st}" if rng.random() < 0.3 else f"{lo}:{hi}:{st}")
            p = p.slice(",".join(specs))
        elif k == 2 and p.ndims < 4:
            p = p.dummy(int(rng.integers(0, p.ndims + 1)), int(rng.integers(1, 4)))
        elif k == 3 and p.ndims >= 2:
            p = p.mv(0, p.ndims - 1)
    return p


@pytest.mark.parametrize("seed", range(24))
def test_random_ops_cuda_vs_oracle(cuda_engine, oracle_engine, seed):
    engines = [cuda_engine, oracle_engine]
    rng = np.random.default_rng(9000 + seed)
    for it in range(12):
        t = ALL_TYPES[int(rng.integers(0, len(ALL_TYPES)))]
        shape = (int(rng.integers(1, 40)), int(rng.integers(1, 3000)))
        flav = "exact" if t in (T.F, T.D) else "small"
        a = rand_array(rng, t, shape, flav)
        flagged = rng.random() < 0.5
        if flagged:
            a[rng.random(shape) < 0.05] = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
        state = rng.bit_generator.state
        views = []
        for (pa,) in zip(both(engines, a, t, flagged)):
            rng.bit_generator.state = state           # the same view chain on both engines
            views.append(random_view(rng, pa))
        ga, oa = views
        kind = int(rng.integers(0, 5))
        tag = f"seed{seed}-it{it}-{T.NAMES[t]}-{ga.dims}"
        if kind == 0:
            op = BIOPS[int(rng.integers(0, len(BIOPS)))]
            b = rand_array(rng, t, (ga.dims[0],) if rng.random() < 0.5 else (), flav) if ga.ndims else rand_array(rng, t, (), flav)
            (gb, ob) = both(engines, np.asarray(b), t)
            assert_same(f"{op}-{tag}", P.run_biop(op, ga, gb), P.run_biop(op, oa, ob))
        elif kind == 1:
            op = UNARY[int(rng.integers(0, len(UNARY)))]
            assert_same(f"{op}-{tag}", P.run_ufunc(op, ga), P.run_ufunc(op, oa))
        elif kind == 2:
            op = REDUCE[int(rng.integers(0, len(REDUCE)))]
            assert_same(f"{op}-{tag}", getattr(ufunc, op)(ga), getattr(ufunc, op)(oa), nan_equal=True)
        elif kind == 3:
            op = BADOPS[int(rng.integers(0, len(BADOPS)))]
            args = (3,) if op in ("setbadtoval", "setvaltobad") else ()
            assert_same(f"{op}-{tag}", getattr(B, op)(ga, *args), getattr(B, op)(oa, *args))
        else:
            for k, (g, o) in enumerate(zip(ufunc.minmaximum(ga), ufunc.minmaximum(oa))):
                assert_same(f"minmaximum[{k}]-{tag}", g, o)
