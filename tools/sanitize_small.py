"""Small launches of the round-2 kernels for compute-sanitizer (memcheck / racecheck): the single-pass scan at its
smallest eligible size (with a BAD element and a tail), the stream-K matmult on cut tiles, the exact-order matmult in
BAD mode, minmaximum with the deferred flag, _n_ind, complex arithmetic.  Results are checked against numpy."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pdl_b200 as P  # noqa: E402
from pdl_b200 import types as T, ufunc  # noqa: E402

eng = P.CudaEngine(0)
rng = np.random.default_rng(9)

# one-pass scan: 296+ tiles of 36 KB, 3 elements after the last vector, BAD elements
n = 300 * 9216 + 3
x = rng.integers(-8, 9, size=n).astype(np.float32)
px = P.PDL.from_numpy(x, T.F, eng)
got = ufunc.cumusumover(px).to_numpy()
assert np.array_equal(got, np.cumsum(x.astype(np.float64)).astype(np.float32)), "scan"
xb = x.copy(); bad = rng.random(n) < 0.01; xb[bad] = -9999.0
pb = P.PDL.from_numpy(xb, T.F, eng).set_badvalue(-9999.0).set_badflag(True)
out = ufunc.cumusumover(pb)
want = np.cumsum(np.where(bad, 0, xb).astype(np.float64)).astype(np.float32); want[bad] = out.badvalue
assert np.array_equal(out.to_numpy(), want), "scan bad"
d = rng.integers(-100, 101, size=300 * 4608).astype(np.float64)
assert np.array_equal(ufunc.cumusumover(P.PDL.from_numpy(d, T.D, eng)).to_numpy(), np.cumsum(d)), "scan double"

# stream-K matmult: 6 tiles cut across ~47 CTAs; ragged edges
a = (rng.integers(-64, 64, size=(384, 1000)) / 64).astype(np.float64)
b = (rng.integers(-64, 64, size=(1000, 250)) / 64).astype(np.float64)
c = P.matmult(P.PDL.from_numpy(a, T.D, eng), P.PDL.from_numpy(b, T.D, eng))
assert eng.last_kernel() == "matmult_dmma_tma" and np.array_equal(c.to_numpy(), a @ b), "stream-K"

# exact-order matmult, BAD mode, every stop state
for t in (T.SB, T.S, T.L, T.F, T.D):
    dt = T.NP_DTYPE[t]
    a = rng.integers(-8, 9, size=(70, 200)).astype(dt); b = rng.integers(-8, 9, size=(200, 90)).astype(dt)
    a[rng.random(a.shape) < 0.002] = np.array(T.DEFAULT_BAD[t]).astype(dt)
    pa = P.PDL.from_numpy(a, t, eng).set_badflag(True)
    P.matmult(pa, P.PDL.from_numpy(b, t, eng)).to_numpy()

# minmaximum (deferred flag), _n_ind, complex
m = rng.random((64, 5000)).astype(np.float32); m[3, :] = np.nan
outs = ufunc.minmaximum(P.PDL.from_numpy(m, T.F, eng))
assert outs[0].badflag
flat = ufunc.minmaximum(P.PDL.from_numpy(rng.random(3_000_000).astype(np.float32), T.F, eng))
assert not flat[0].badflag
ufunc.maximum_n_ind(P.PDL.from_numpy(rng.integers(0, 50, size=(2000, 7)).astype(np.int32), T.L, eng), 5).to_numpy()
print("sanitize_small: ok")
