"""Worker for tests/test_gpu_parallel.py: one rank of `torchrun --nproc-per-node N` on real GPUs.  Every rank
builds the SAME global ndarray, takes its block by the reference's split rule, runs the product's sharded
reductions (device partial records -> one ncclAllGather -> device merge) and compares the bits with (a) the
ordinary single-GPU reduction of the whole ndarray and (b) the C oracle.  Exit code 0 = all checks passed."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def main():
    import torch
    import torch.distributed as dist
    import pdl_b200 as P
    from pdl_b200 import types as T, ufunc, parallel
    from oracle_engine import OracleEngine

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng, ora = P.CudaEngine(local), OracleEngine()
    comm = parallel.Comm()
    use_peer = len(sys.argv) > 1 and sys.argv[1] == "peer"
    if use_peer:
        assert comm.enable_peer_exchange(eng), "the GPUs of this box cannot map each other's memory"
    rng = np.random.default_rng(2024)                      # same stream on every rank
    checked = 0
    for t in (T.F, T.D, T.L, T.B, T.US, T.LL):
        dt = T.NP_DTYPE[t]
        full = rng.integers(-8, 8, size=(41, 3001), endpoint=True).astype(dt)
        if t in T.UNSIGNED:
            full = np.abs(full.astype(np.int64)).astype(dt)
        bad = np.array(T.DEFAULT_BAD[t]).astype(dt)
        full[rng.random(full.shape) < 0.03] = bad
        if t in (T.F, T.D):
            full[5, 17] = np.nan
            full[40, 2999] = -0.0
        for badflag in (True, False):
            whole = P.PDL.from_numpy(full, t, eng).set_badflag(badflag)
            owhole = P.PDL.from_numpy(full, t, ora).set_badflag(badflag)
            mine = parallel.shard(whole, rank, world)
            outs = parallel.pcollapse(mine.flat(), comm, ("sum", "avg", "min", "max", "min_ind", "max_ind", "dsum", "davg"))
            refs = [ufunc.sumover, ufunc.average, ufunc.minimum, ufunc.maximum, ufunc.minimum_ind, ufunc.maximum_ind,
                    ufunc.dsumover, ufunc.daverage]
            for k, (o, f) in enumerate(zip(outs, refs)):
                for label, wp in (("single-gpu", whole), ("oracle", owhole)):
                    w = f(wp.flat())
                    assert o.type == w.type and o.badflag == w.badflag, (T.NAMES[t], k, label, o.type, w.type, o.badflag, w.badflag)
                    g, e = o.to_numpy(), w.to_numpy()
                    assert g.tobytes() == e.tobytes() or (g.dtype.kind == "f" and np.isnan(g) and np.isnan(e)), \
                        (T.NAMES[t], badflag, k, label, g, e)
                    checked += 1
            # a reduction over dim 0 when dim 0 is the sharded one: one record per row
            xw = whole.xchg(0, 1)                          # [41, 3001] -> dim 0 = 41 is split
            mine0 = parallel.shard(xw, rank, world, dim=0)
            got = parallel.psumover(mine0, comm)
            want = ufunc.sumover(xw)
            assert got.dims == want.dims and got.to_numpy().tobytes() == want.to_numpy().tobytes(), (T.NAMES[t], "psumover")
            checked += 1
    # all-BAD everywhere, and a rank with an empty block
    allbad = P.PDL.from_numpy(np.full((4, 6), T.DEFAULT_BAD[T.F], dtype=np.float32), T.F, eng).set_badflag(True)
    r = parallel.psum(parallel.shard(allbad, rank, world), comm)
    assert r.badflag and r.to_numpy() == np.float32(T.DEFAULT_BAD[T.F])
    tiny = P.PDL.from_numpy(np.arange(3, dtype=np.int32).reshape(1, 3), T.L, eng)     # 1 row over N ranks
    assert parallel.psum(parallel.shard(tiny, rank, world), comm).sclr() == 3
    assert parallel.pmax_ind(parallel.shard(tiny, rank, world), comm).sclr() == 2
    # identical bits on every rank
    sig = torch.tensor([checked], dtype=torch.int64, device="cuda")
    dist.all_reduce(sig)
    assert int(sig.item()) == checked * world
    dist.barrier()
    dist.destroy_process_group()
    if use_peer:
        assert comm._peer["epoch"] >= 10, "the peer-memory path was not the one that ran"   # (records beyond a mailbox go by NCCL)
    print(f"rank {rank}: {checked} sharded-reduction checks ok ({'peer-memory exchange' if use_peer else 'nccl all-gather'})")


if __name__ == "__main__":
    main()
