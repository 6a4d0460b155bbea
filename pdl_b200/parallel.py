"""Multi-GPU: one process per GPU, the ndarray partitioned along its OUTERMOST broadcast dim.

This generalises the reference's autopthread split (lib/PDL/Core/pdlbroadcast.c:469-484,
diagram lib/PDL/Core/pdlbroadcast.h:39-60): worker i of nthr gets a contiguous block of the
chosen dim, the first `dim % nthr` workers one element more.  Elementwise ops and reductions
over a non-sharded dim need NO collective — every rank runs the ordinary single-GPU call on
its block (bench.py does exactly that).  Only a reduction that collapses the sharded dim (the
whole-array wrappers sum / avg / min / max ..., lib/PDL/Ufunc.pd:618-663) exchanges data:
each rank reduces its block on the device to ONE partial record (PART_* ops), the records are
all-gathered over NCCL (32 bytes per rank over NVLink/NVSwitch; gloo on CPU for the tests) and every
rank merges them ON THE DEVICE (COLL_* ops) in rank order with the reference's own semantics (BAD -> skipped, all BAD -> BAD,
NaN loses to non-NaN, first index wins), so all ranks hold the same bit pattern.
"""
from __future__ import annotations

import numpy as np

from . import types as T
from .core import PDL, _default_incs
from .engine import PDLError, Store
from . import ufunc


def split_dim(n: int, nthr: int) -> list:
    """[(start, count)] per worker: pdl_initbroadcaststruct's mag_stride/mag_skip rule
    (pdlbroadcast.c:469-484): count = n // nthr, the first n % nthr workers get one more."""
    base, rem = divmod(n, nthr)
    out, start = [], 0
    for i in range(nthr):
        c = base + (1 if i < rem else 0)
        out.append((start, c))
        start += c
    return out


def shard(p: PDL, rank: int, world: int, dim: int = -1) -> PDL:
    """This rank's block of `p` along `dim` (default: outermost), as a view."""
    d = dim % p.ndims
    start, cnt = split_dim(p.dims[d], world)[rank]
    if cnt == 0:
        return p._view(p.dims[:d] + [0] + p.dims[d + 1:], p.dimincs, p.offs)
    spec = ",".join((f"{start}:{start + cnt - 1}" if k == d else ":") for k in range(p.ndims))
    return p.slice(spec)


class Comm:
    """Thin wrapper over a torch.distributed process group (nccl on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self._bufs = {}

    def all_gather_bytes(self, rec: np.ndarray) -> np.ndarray:
        """rec: uint8[k] on the host -> uint8[world, k], identical on every rank."""
        import torch
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = torch.from_numpy(rec.copy()).to(dev)
        out = torch.empty((self.world, rec.size), dtype=torch.uint8, device=dev)
        self.dist.all_gather_into_tensor(out, t, group=self.group) if dev == "cuda" else \
            self.dist.all_gather(list(out.unbind(0)), t, group=self.group)
        return out.cpu().numpy()

    # ---- device-side record exchange -----------------------------------------------------------
    def record_buffers(self, engine, nwords: int):
        """(local, gathered) longlong ndarrays of `nwords` and `world * nwords` elements that
        torch.distributed can address: device memory under nccl, host memory under gloo.  Cached per size;
        reuse is safe because every use is ordered on the engine's stream."""
        key = (id(engine), nwords)
        if key not in self._bufs:
            import torch
            dev = f"cuda:{engine.device}" if self.backend == "nccl" else "cpu"
            lt = torch.zeros(nwords, dtype=torch.int64, device=dev)
            gt = torch.zeros(self.world * nwords, dtype=torch.int64, device=dev)
            mk = (lambda t: engine.wrap(t.data_ptr(), t.numel() * 8, t)) if hasattr(engine, "wrap") else \
                (lambda t: Store(engine, None, t.data_ptr(), t.numel() * 8, t))
            self._bufs[key] = (lt, gt, PDL(engine, mk(lt), T.LL, [nwords]), PDL(engine, mk(gt), T.LL, [nwords, self.world]))
        return self._bufs[key]

    # ---- peer-memory exchange (NVLink / NVSwitch P2P): one kernel instead of a library collective ----------------
    PEER_CAP = 4096          # int64 words per rank a mailbox holds (1024 records)

    def enable_peer_exchange(self, engine) -> bool:
        """Give every rank a mailbox in its own HBM that all peers map through CUDA IPC (handles exchanged ONCE here,
        over the process group).  Afterwards `exchange()` is a single kernel per call (pdlb200_peer_exchange): stores
        over NVLink + epoch flags, no NCCL on the data path.  Returns False (and keeps the NCCL path) when the GPUs
        cannot map each other's memory."""
        import ctypes as C
        if self.backend != "nccl" or getattr(self, "_peer", None) is not None:
            return getattr(self, "_peer", None) is not None
        lib = engine.lib
        err = C.create_string_buffer(256)
        box = lib.pdlb200_peer_mailbox_new(self.world, self.PEER_CAP)
        handle = C.create_string_buffer(64)
        ok = bool(box) and lib.pdlb200_ipc_export(box, handle, err, 256) == 0
        handles = [None] * self.world
        self.dist.all_gather_object(handles, (ok, handle.raw), group=self.group)
        if not all(h[0] for h in handles):
            return False
        ptrs, good = [], True
        for r, (_, raw) in enumerate(handles):
            p = box if r == self.rank else lib.pdlb200_ipc_open(raw, err, 256)
            good = good and bool(p)
            ptrs.append(p or 0)
        flags = [None] * self.world
        self.dist.all_gather_object(flags, good, group=self.group)
        if not all(flags):
            return False
        import numpy as _np
        table = engine.alloc(8 * self.world)
        engine.upload(table, _np.array(ptrs, dtype=_np.uint64).view(_np.uint8))
        nbytes = 2 * (self.world * self.PEER_CAP + self.world) * 8
        self._peer = {"box": box, "ptrs": ptrs, "table": table, "epoch": 0,
                      "pdl": PDL(engine, engine.wrap(box, nbytes, None), T.LL, [nbytes // 8])}
        return True

    def exchange(self, engine, nwords: int):
        """(local, gathered): `local` = longlong [nwords] this rank fills; calling the returned `go()` moves the
        records of all ranks into `gathered` = longlong [nwords, world] — ONE kernel over peer memory when
        enable_peer_exchange() succeeded and the records fit a mailbox, else ONE all-gather of the process group."""
        lt, gt, lrec, grec = self.record_buffers(engine, nwords)
        peer = getattr(self, "_peer", None)
        if peer is None or nwords > self.PEER_CAP:
            return lrec, grec, (lambda: self.all_gather_records(engine, lt, gt))
        import ctypes as C
        lib, cap = engine.lib, self.PEER_CAP

        def go():
            peer["epoch"] += 1
            err = C.create_string_buffer(256)
            rc = lib.pdlb200_peer_exchange(lt.data_ptr(), nwords, cap, peer["table"].ptr, self.rank, self.world, peer["epoch"],
                                           engine.stream, err, 256)
            if rc != 0:
                raise PDLError(err.value.decode("utf-8", "replace"))
            off = lib.pdlb200_peer_gathered_offset(self.world, cap, peer["epoch"])
            go.gathered = peer["pdl"]._view([nwords, self.world], [1, cap], off)
        go.gathered = None
        return lrec, None, go

    def all_gather_records(self, engine, lt, gt) -> None:
        """ONE collective: every rank's `lt` into `gt` (rank-major), on the engine's stream."""
        if self.backend == "nccl":
            import contextlib
            import torch
            cm = torch.cuda.stream(torch.cuda.ExternalStream(engine.stream)) if getattr(engine, "stream", None) else contextlib.nullcontext()
            with cm:
                self.dist.all_gather_into_tensor(gt, lt, group=self.group)
        else:
            self.dist.all_gather(list(gt.view(self.world, -1).unbind(0)), lt, group=self.group)

    def block_offset(self, n_local: int):
        """(global index of this rank's element 0, total length) for blocks laid end to end in rank order.
        One small host collective; pass offset=/total= to the reductions to skip it."""
        sizes = self.all_gather_bytes(np.array([n_local], dtype=np.int64).view(np.uint8)).view(np.int64).reshape(-1)
        return int(sizes[:self.rank].sum()), int(sizes.sum())


# kind -> (PART op that makes its record, value type of the record)
def _part_of(kind: str, t: int):
    if kind in ("sum", "avg"):
        return "part_sum", T.int_plus(t)
    if kind in ("dsum", "davg"):
        return "part_dsum", T.D
    if kind in ("min", "min_ind"):
        return "part_min", t
    if kind in ("max", "max_ind"):
        return "part_max", t
    raise PDLError(f"unknown sharded reduction '{kind}'")


_COLL_OF = {"sum": "sum", "dsum": "sum", "avg": "avg", "davg": "avg", "min": "min", "max": "max",
            "min_ind": "min_ind", "max_ind": "max_ind"}


def pcollapse(local: PDL, comm: Comm, kinds, offset: int | None = None, total: int | None = None) -> list:
    """Reductions over dim 0 of an ndarray whose dim 0 is SHARDED across `comm` (blocks in rank order):
    `local` is this rank's block [n_local, rows...]; returns one [rows...] ndarray per entry of `kinds`
    ('sum' 'avg' 'dsum' 'davg' 'min' 'max' 'min_ind' 'max_ind'), identical bits on every rank.

    Everything stays on the device and on the engine's stream: one PART_* launch per distinct partial
    (sum+avg share one pass, max+max_ind too) writes 32-byte records, ONE all-gather moves the records of all
    kinds, one COLL_* launch per kind merges them in rank order with the reference's BAD / NaN / first-index
    rules (include/pdlb200.h).  No host synchronisation unless `offset` has to be discovered."""
    from .trans import run_op, collapse_records
    if local.ndims < 1:
        local = local.dummy(0)
    n_local, rows = local.dims[0], local.dims[1:]
    if offset is None:
        offset, total = comm.block_offset(n_local)
    nrows = 1
    for d in rows:
        nrows *= d
    parts = []                                   # distinct partial ops, in first-use order
    for k in kinds:
        pk = _part_of(k, local.datatype)
        if pk not in parts:
            parts.append(pk)
    nwords = 4 * len(parts) * nrows
    if hasattr(comm, "exchange"):
        lrec, grec, go = comm.exchange(local.engine, nwords)
    else:                                        # minimal communicators (tests): record_buffers + all_gather_records
        lt, gt, lrec, grec = comm.record_buffers(local.engine, nwords)
        go = (lambda: comm.all_gather_records(local.engine, lt, gt))
    lview = lrec.reshape_view([4, len(parts)] + rows)
    for j, (pname, _vt) in enumerate(parts):
        run_op(pname, [local], [lview.slice(f":,({j})")], goff=offset)
    go()
    if grec is None:                             # peer-memory path: the gathered records sit in this rank's mailbox
        g = go.gathered                          # [nwords, world], rank stride = mailbox capacity
        gview = g._view([4, len(parts)] + rows + [comm.world],
                        _default_incs([4, len(parts)] + rows) + [g.dimincs[1]], g.offs)
    else:
        gview = grec.reshape_view([4, len(parts)] + rows + [comm.world])
    outs = []
    for k in kinds:
        pk = _part_of(k, local.datatype)
        recs = gview.slice(f":,({parts.index(pk)})").mv(-1, 1)      # [4, world, rows...]
        out = collapse_records(_COLL_OF[k], recs, pk[1], local.badflag)
        if total == 0 and k in ("min", "max", "min_ind", "max_ind"):
            out.badflag = True                   # no element anywhere: BAD + badflag even in good mode (Ufunc.pd:463-464)
        outs.append(out)
    return outs


def _whole(local: PDL, comm: Comm, kind: str, offset=None, total=None) -> PDL:
    """Whole-array wrapper (`$x->flat->Xover`, lib/PDL/Ufunc.pd:618-663) of an ndarray sharded along its
    outermost dim: the local block is a contiguous run of the global flat ndarray."""
    return pcollapse(local.flat(), comm, (kind,), offset, total)[0]


def psum(local: PDL, comm: Comm, offset=None, total=None) -> PDL:
    """sum() of an ndarray sharded across comm (flat->sumover + collapse of the sharded dim)."""
    return _whole(local, comm, "sum", offset, total)


def pavg(local: PDL, comm: Comm, offset=None, total=None) -> PDL:
    return _whole(local, comm, "avg", offset, total)


def pmin(local: PDL, comm: Comm, offset=None, total=None) -> PDL:
    return _whole(local, comm, "min", offset, total)


def pmax(local: PDL, comm: Comm, offset=None, total=None) -> PDL:
    return _whole(local, comm, "max", offset, total)


def pmin_ind(local: PDL, comm: Comm, offset=None, total=None) -> PDL:
    return _whole(local, comm, "min_ind", offset, total)


def pmax_ind(local: PDL, comm: Comm, offset=None, total=None) -> PDL:
    return _whole(local, comm, "max_ind", offset, total)


def psumover(local: PDL, comm: Comm, offset=None, total=None) -> PDL:
    """sumover of an ndarray whose REDUCED dim (dim 0) is the sharded one."""
    return pcollapse(local, comm, ("sum",), offset, total)[0]


def pminmax(local: PDL, comm: Comm):
    """minmax() of a sharded ndarray (lib/PDL/Ufunc.pd:738 over :563-613): ONE local pass (minmaximum on the
    flat block) instead of separate min and max reductions, then 48-byte records all-gathered and merged in
    rank order.  BAD and NaN elements are skipped; no usable element anywhere -> (BAD, BAD).
    Returns two 0-dim ndarrays."""
    flat = local.flat()
    rec = np.zeros(6, dtype=np.uint64)           # have, min bits, min index, max bits, max index, pad
    sizes = comm.all_gather_bytes(np.array([local.nelem], dtype=np.int64).view(np.uint8)).view(np.int64).reshape(-1)
    offset = int(sizes[:comm.rank].sum())
    if flat.nelem:
        cmin, cmax, imin, imax = ufunc.minmaximum(flat)
        if not (cmin.badflag and cmin.bad_mask().any()):     # the flag alone may just be propagated from the input
            rec[:] = [1, T.value_bits(cmin.datatype, cmin.sclr()), int(imin.sclr()) + offset,
                      T.value_bits(cmax.datatype, cmax.sclr()), int(imax.sclr()) + offset, 0]
    recs = comm.all_gather_bytes(rec.view(np.uint8)).view(np.uint64).reshape(comm.world, 6)
    t = local.datatype
    best = None
    for r in recs:
        if not r[0]:
            continue
        mn, mx = T.bits_value(t, int(r[1])), T.bits_value(t, int(r[3]))
        if best is None:
            best = [mn, int(r[2]), mx, int(r[4])]
            continue
        if mn < best[0] or (mn == best[0] and r[2] < best[1]):
            best[0], best[1] = mn, int(r[2])
        if mx > best[2] or (mx == best[2] and r[4] < best[3]):
            best[2], best[3] = mx, int(r[4])
    outs = []
    for v in ((best[0], best[2]) if best else (T.DEFAULT_BAD[t], T.DEFAULT_BAD[t])):
        o = PDL.from_numpy(np.array(v, dtype=T.NP_DTYPE[t]), t, local.engine)
        o.badflag = best is None or local.badflag
        outs.append(o)
    return tuple(outs)


def pinner(a_local: PDL, b_local: PDL, comm: Comm) -> PDL:
    """inner() of two ndarrays sharded the same way over their (flattened) n (lib/PDL/Primitive.pd:48-70): the
    local dot product on the device, then the per-rank partials added in rank order; a BAD element on any rank
    makes the result BAD."""
    from .primitive import inner
    part = inner(a_local.flat(), b_local.flat())
    t = part.datatype
    bad = 1 if part.badflag and part.bad_mask().any() else 0
    rec = np.array([bad, T.value_bits(t, part.sclr()) if not bad else 0], dtype=np.uint64)
    recs = comm.all_gather_bytes(rec.view(np.uint8)).view(np.uint64).reshape(comm.world, 2)
    dt = T.NP_DTYPE[t]
    anybad = bool(recs[:, 0].any())
    with np.errstate(over="ignore"):
        tot = dt.type(0)
        for r in recs:
            if not r[0]:
                tot = dt.type(tot + T.bits_value(t, int(r[1])))
    out = PDL.from_numpy(np.array(T.DEFAULT_BAD[t] if anybad else tot, dtype=dt), t, a_local.engine)
    out.badflag = anybad or a_local.badflag or b_local.badflag
    return out


__all__ = ["split_dim", "shard", "Comm", "pcollapse", "psumover", "psum", "pavg", "pmin", "pmax", "pmin_ind", "pmax_ind", "pminmax", "pinner"]
