"""CPU, world_size 2 over gloo: the N>1 host logic — partition rule, per-rank blocks, and the
collapse of the sharded dim (all-gather of partial records + rank-ordered finish) — with the
oracle engine standing in for the GPUs.  Results must equal the single-process reduction of the
whole ndarray bit for bit, on every rank."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import pdl_b200 as P
    from pdl_b200 import types as T, ufunc, parallel
    from oracle_engine import OracleEngine
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = OracleEngine()
        comm = parallel.Comm()
        rng = np.random.default_rng(77)                      # same data on both ranks
        res = {}
        for t in (T.F, T.D, T.L, T.B, T.LL):
            full = rng.integers(-8, 8, size=(37, 501), endpoint=True).astype(T.NP_DTYPE[t])
            if t == T.B:
                full = np.abs(full.astype(np.int16)).astype(np.uint8)
            bad = np.array(T.DEFAULT_BAD[t]).astype(T.NP_DTYPE[t])
            full[rng.random(full.shape) < 0.02] = bad
            if t in (T.F, T.D):
                full[3, 7] = np.nan
            whole = P.PDL.from_numpy(full, t, eng).set_badflag(True)
            mine = parallel.shard(whole, rank, world)        # rows (outermost dim) split by the reference rule
            # (1) no collective: reduction over the NON-sharded dim, block by block
            part = ufunc.sumover(mine).to_numpy()
            want = ufunc.sumover(whole).to_numpy()
            s0, c0 = parallel.split_dim(whole.dims[-1], world)[rank]
            assert np.array_equal(part.view(np.uint8), want[s0:s0 + c0].view(np.uint8))
            # (2) collective: whole-array reductions collapse the sharded dim
            for name, fn, ref in (("sum", parallel.psum, ufunc.sum), ("avg", parallel.pavg, ufunc.avg),
                                  ("min", parallel.pmin, ufunc.min), ("max", parallel.pmax, ufunc.max)):
                got, exp = fn(mine, comm), ref(whole)
                assert got.type == exp.type, (name, got.type, exp.type)
                g, e = got.to_numpy(), exp.to_numpy()
                assert g.tobytes() == e.tobytes() or (np.isnan(g) and np.isnan(e)), (name, T.NAMES[t], g, e)
                res[f"{name}-{T.NAMES[t]}"] = g.tobytes().hex()
            gi = parallel.pmax_ind(mine, comm).sclr()
            assert gi == ufunc.maximum_ind(whole.flat()).sclr(), ("max_ind", T.NAMES[t])
            # several reductions in ONE exchange, explicit offset/total (no host collective at all)
            blocks = parallel.split_dim(whole.dims[-1], world)
            off = sum(c for _, c in blocks[:rank]) * whole.dims[0]
            many = parallel.pcollapse(mine.flat(), comm, ("dsum", "davg", "min_ind", "sum"), offset=off, total=whole.nelem)
            for o, f in zip(many, (ufunc.dsumover, ufunc.daverage, ufunc.minimum_ind, ufunc.sumover)):
                e = f(whole.flat())
                assert o.type == e.type and o.badflag == e.badflag
                assert o.to_numpy().tobytes() == e.to_numpy().tobytes() or np.isnan(o.to_numpy()), (f.__name__, T.NAMES[t])
            # the REDUCED dim is the sharded one: one record per row, result has the row dims
            xw = whole.xchg(0, 1)
            rows = parallel.psumover(parallel.shard(xw, rank, world, dim=0), comm)
            assert rows.dims == [whole.dims[0]] and rows.to_numpy().tobytes() == ufunc.sumover(xw).to_numpy().tobytes()
            # minmax in one local pass (minmaximum) + merge; inner of two identically sharded ndarrays
            gmn, gmx = parallel.pminmax(mine, comm)
            emn, emx, _, _ = ufunc.minmaximum(whole.flat())
            assert gmn.to_numpy().tobytes() == emn.to_numpy().tobytes(), ("minmax-min", T.NAMES[t])
            assert gmx.to_numpy().tobytes() == emx.to_numpy().tobytes(), ("minmax-max", T.NAMES[t])
            other = np.abs(rng.integers(0, 3, size=full.shape)).astype(T.NP_DTYPE[t])
            clean = np.where(full == bad, 1, full).astype(T.NP_DTYPE[t])
            if t in (T.F, T.D):
                clean[3, 7] = 2
            if t == T.B:
                clean, other = clean % 2, other % 2          # keep the dot product inside the 8-bit result type
            wa, wb = P.PDL.from_numpy(clean, t, eng), P.PDL.from_numpy(other, t, eng)
            gin = parallel.pinner(parallel.shard(wa, rank, world), parallel.shard(wb, rank, world), comm)
            ein = P.inner(wa.flat(), wb.flat())
            assert gin.to_numpy().tobytes() == ein.to_numpy().tobytes(), ("inner", T.NAMES[t], gin.to_numpy(), ein.to_numpy())
            gbad = parallel.pinner(mine, parallel.shard(wb, rank, world), comm)     # BAD elements on some rank -> BAD
            assert gbad.badflag and gbad.bad_mask().all(), ("inner-bad", T.NAMES[t])
        # all-BAD and empty blocks
        allbad = P.PDL.from_numpy(np.full((4, 6), T.DEFAULT_BAD[T.F], dtype=np.float32), T.F, eng).set_badflag(True)
        r = parallel.psum(parallel.shard(allbad, rank, world), comm)
        assert r.badflag and r.to_numpy() == np.float32(T.DEFAULT_BAD[T.F])
        tiny = P.PDL.from_numpy(np.arange(3, dtype=np.int32).reshape(1, 3), T.L, eng)   # 1 row over 2 ranks
        assert parallel.psum(parallel.shard(tiny, rank, world), comm).sclr() == 3
        q.put((rank, "ok", res))
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), {}))
    finally:
        dist.destroy_process_group()


def test_split_dim_matches_reference_rule():
    from pdl_b200.parallel import split_dim
    assert split_dim(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]      # pdlbroadcast.h:39-60 example shape
    assert split_dim(8, 8) == [(i, 1) for i in range(8)]
    assert split_dim(3, 4) == [(0, 1), (1, 1), (2, 1), (3, 0)]
    assert sum(c for _, c in split_dim(65536, 8)) == 65536


def test_sharded_reductions_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, _ in out:
        assert status == "ok", f"rank {rank}: {status}"
    assert out[0][2] == out[1][2], "ranks disagree on the reduced bit patterns"
