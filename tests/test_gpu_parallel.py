"""GPU: the sharded whole-array reductions — PART_* / COLL_* kernels against the oracle in one
process, and the full path (device partial records -> ncclAllGather -> device merge) under torchrun on every
pair of GPUs the box has."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import pdl_b200 as P
from pdl_b200 import types as T, ufunc, parallel
from parity import ALL_TYPES

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


class Solo:
    """world-size-1 communicator: the 'gather' of one rank's records is a copy."""
    rank, world, backend = 0, 1, "gloo"
    record_buffers = parallel.Comm.record_buffers

    def __init__(self):
        self._bufs = {}

    def all_gather_records(self, engine, lt, gt):
        engine.sync()
        gt.copy_(lt)


class SoloCuda(Solo):
    backend = "nccl"

    def all_gather_records(self, engine, lt, gt):
        gt.copy_(lt)


@pytest.mark.parametrize("t", ALL_TYPES, ids=lambda t: T.NAMES[t])
def test_partial_records_and_merge_match_oracle(cuda_engine, oracle_engine, t):
    rng = np.random.default_rng(900 + t)
    dt = T.NP_DTYPE[t]
    for shape, badflag in (((7, 70001), True), ((3, 70001), False), ((300, 257), True), ((1, 5), True)):
        full = rng.integers(0 if t in T.UNSIGNED else -9, 9, size=shape, endpoint=True).astype(dt)
        full[rng.random(shape) < 0.05] = np.array(T.DEFAULT_BAD[t]).astype(dt)
        if t in (T.F, T.D) and shape[1] > 100:
            full[0, 3] = np.nan
            full[-1, :] = np.nan                                    # a row whose good values are all NaN
        outs = []
        for e, comm in ((cuda_engine, SoloCuda()), (oracle_engine, Solo())):
            p = P.PDL.from_numpy(full, t, e).set_badflag(badflag)
            kinds = ("sum", "avg", "min", "max", "min_ind", "max_ind", "dsum", "davg")
            got = parallel.pcollapse(p, comm, kinds, offset=1_000_000_007, total=10 ** 10)
            lrec = comm.record_buffers(e, 4 * 4 * shape[0])[2]
            outs.append((got, lrec.to_numpy().copy(), p))
        (g, grec, gp), (o, orec, _op) = outs
        # records (numpy C order of [4, parts, rows]; parts in first-use order: sum, min, max, dsum): identical,
        # except the float SUM bits, which depend on the summation order -> those are compared as values below
        gr, orr = grec.reshape(shape[0], 4, 4), orec.reshape(shape[0], 4, 4)
        assert np.array_equal(gr[:, 1:3, :], orr[:, 1:3, :]), "min/max records differ"
        assert np.array_equal(gr[:, (0, 3), 1:], orr[:, (0, 3), 1:]), "sum records: count/state differ"
        if t in T.INTEGER:
            assert np.array_equal(gr[:, 0, 0], orr[:, 0, 0]), "integer sum records differ"
        for k, (a, b) in enumerate(zip(g, o)):
            assert a.type == b.type and a.dims == b.dims and a.badflag == b.badflag
            x, y = a.to_numpy(), b.to_numpy()
            if x.dtype.kind == "f" and k in (0, 1, 6, 7):
                assert np.array_equal(np.isnan(x), np.isnan(y)) and np.allclose(x[~np.isnan(x)], y[~np.isnan(y)], rtol=1e-6), (k, x, y)
            else:
                assert x.tobytes() == y.tobytes(), (T.NAMES[t], shape, k, x, y)
        # and against the ordinary reductions (global offset removed)
        want_ind = ufunc.maximum_ind(gp).to_numpy()
        got_ind = g[5].to_numpy()
        bad_ind = np.array(T.DEFAULT_BAD[T.IND])
        assert np.array_equal(np.where(want_ind == bad_ind, bad_ind, want_ind + 1_000_000_007), got_ind)
        assert g[3].to_numpy().tobytes() == ufunc.maximum(gp).to_numpy().tobytes()
        assert g[2].to_numpy().tobytes() == ufunc.minimum(gp).to_numpy().tobytes()


def test_whole_array_wrappers_solo(cuda_engine, oracle_engine):
    rng = np.random.default_rng(31)
    full = rng.integers(-1, 1, size=(3, 5, 100_003), endpoint=True).astype(np.float32)
    full[1, 2, 77_777] = 3.0
    p = P.PDL.from_numpy(full, T.F, cuda_engine)
    comm = SoloCuda()
    assert parallel.psum(p, comm, offset=0, total=p.nelem).sclr() == float(full.astype(np.float64).sum())
    assert parallel.pmax(p, comm, offset=0, total=p.nelem).sclr() == 3.0
    assert parallel.pmax_ind(p, comm, offset=0, total=p.nelem).sclr() == int(np.argmax(full.reshape(-1)))
    assert parallel.pavg(p, comm, offset=0, total=p.nelem).to_numpy().tobytes() == ufunc.avg(p).to_numpy().tobytes()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _visible_gpus():
    return P.default_engine().lib.pdlb200_device_count()


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_sharded_reductions_nccl(cuda_engine, nproc, exchange):
    """torchrun --nproc-per-node N on real GPUs, the records exchanged by ncclAllGather and by the peer-memory
    kernel (NVLink stores into IPC-mapped mailboxes); skipped (not failed) where the box has fewer GPUs."""
    if _visible_gpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs, {_visible_gpus()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(ROOT / "tests" / "nccl_worker.py"), exchange]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert r.stdout.count("sharded-reduction checks ok") == nproc
