"""Multi-GPU: one process per GPU, the ndarray partitioned along its OUTERMOST broadcast dim.

This generalises the reference's autopthread split (lib/PDL/Core/pdlbroadcast.c:469-484,
diagram lib/PDL/Core/pdlbroadcast.h:39-60): worker i of nthr gets a contiguous block of the
chosen dim, the first `dim % nthr` workers one element more.  Elementwise ops and reductions
over a non-sharded dim need NO collective — every rank runs the ordinary single-GPU call on
its block (bench.py does exactly that).  Only a reduction that collapses the sharded dim (the
whole-array wrappers sum / avg / min / max ..., lib/PDL/Ufunc.pd:618-663) exchanges data:
each rank reduces its block on the device to ONE partial record, the records are all-gathered
over NCCL (32 bytes per rank over NVLink/NVSwitch; gloo on CPU for the tests) and every rank
finishes them in rank order with the reference's own semantics (BAD -> skipped, all BAD -> BAD,
NaN loses to non-NaN, first index wins), so all ranks hold the same bit pattern.
"""
from __future__ import annotations

import numpy as np

from . import types as T
from .core import PDL
from .engine import PDLError
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

    def all_gather_bytes(self, rec: np.ndarray) -> np.ndarray:
        """rec: uint8[k] on the host -> uint8[world, k], identical on every rank."""
        import torch
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = torch.from_numpy(rec.copy()).to(dev)
        out = torch.empty((self.world, rec.size), dtype=torch.uint8, device=dev)
        self.dist.all_gather_into_tensor(out, t, group=self.group) if dev == "cuda" else \
            self.dist.all_gather(list(out.unbind(0)), t, group=self.group)
        return out.cpu().numpy()


_REC = np.dtype([("val", "u8"), ("ngood", "i8"), ("idx", "i8"), ("state", "i8")])


def _partial(local: PDL, op: str, offset: int):
    """One 32-byte record for this rank's block (flattened): value bits, good count, GLOBAL index,
    state (0 no good element, 1 value is non-NaN, 2 value is NaN)."""
    flat = local.flat()
    rec = np.zeros((), dtype=_REC)
    if flat.nelem == 0:
        return rec
    ngood = int(ufunc.ngoodover(flat).sclr())
    rec["ngood"] = ngood
    if ngood == 0:
        return rec
    if op in ("sum", "avg"):
        v = ufunc.sumover(flat)
        rec["val"] = T.value_bits(v.datatype, v.sclr())
        rec["state"] = 1
    else:
        red, ind = (ufunc.minimum, ufunc.minimum_ind) if op in ("min", "min_ind") else (ufunc.maximum, ufunc.maximum_ind)
        v = red(flat)
        val = v.sclr()
        rec["val"] = T.value_bits(v.datatype, val)
        rec["idx"] = int(ind(flat).sclr()) + offset
        rec["state"] = 2 if (v.datatype in (T.F, T.D) and val != val) else 1
    return rec


def _finish(recs: np.ndarray, op: str, dtype_id: int):
    """Merge the per-rank records in rank order.  Returns (value, is_bad)."""
    dt = T.NP_DTYPE[dtype_id]
    live = [r for r in recs if r["ngood"] > 0]
    if not live:
        return None, True
    vals = [T.bits_value(dtype_id, int(r["val"])) for r in live]
    if op in ("sum", "avg"):
        with np.errstate(over="ignore"):
            tot = vals[0]
            for v in vals[1:]:
                tot = dt.type(tot + v)
        if op == "sum":
            return tot, False
        cnt = int(sum(int(r["ngood"]) for r in live))
        if dt.kind == "f":
            return dt.type(tot / dt.type(cnt)), False
        if dtype_id == T.ULL:
            return dt.type(int(tot) // cnt), False
        q = abs(int(tot)) // cnt
        return dt.type(q if int(tot) >= 0 else -q), False       # C division truncates toward zero
    ismax = op in ("max", "max_ind")
    best = None
    for r, v in zip(live, vals):
        if best is None:
            best = (r, v)
            continue
        br, bv = best
        if br["state"] == 1 and r["state"] == 1:
            better = (v > bv) if ismax else (v < bv)
            if better or (v == bv and r["idx"] < br["idx"]):
                best = (r, v)
        elif br["state"] == 2 and r["state"] == 1:
            best = (r, v)
        elif br["state"] == 2 and r["state"] == 2 and r["idx"] > br["idx"]:
            best = (r, v)       # every good value is NaN: the reference ends on the LAST one
    r, v = best
    if op.endswith("_ind"):
        return np.int64(r["idx"]), False
    return v, False


def _collapse(local: PDL, comm: Comm, op: str, out_type: int, offset: int | None = None) -> PDL:
    if offset is None:  # global flat index of this rank's first element: exclusive scan of block sizes
        sizes = comm.all_gather_bytes(np.array([local.nelem], dtype=np.int64).view(np.uint8)).view(np.int64).reshape(-1)
        offset = int(sizes[:comm.rank].sum())
    rec = _partial(local, op, offset)
    recs = comm.all_gather_bytes(np.frombuffer(rec.tobytes(), dtype=np.uint8)).view(_REC).reshape(-1)
    val, bad = _finish(recs, op, out_type if not op.endswith("_ind") else local.datatype)
    res_type = T.IND if op.endswith("_ind") else out_type
    if bad:
        val = T.DEFAULT_BAD[res_type]
    out = PDL.from_numpy(np.array(val, dtype=T.NP_DTYPE[res_type]), res_type, local.engine)
    out.badflag = bool(bad) or local.badflag
    return out


def psum(local: PDL, comm: Comm) -> PDL:
    """sum() of an ndarray sharded across comm (flat->sumover + collapse of the sharded dim)."""
    return _collapse(local, comm, "sum", T.int_plus(local.datatype))


def pavg(local: PDL, comm: Comm) -> PDL:
    return _collapse(local, comm, "avg", T.int_plus(local.datatype))


def pmin(local: PDL, comm: Comm) -> PDL:
    return _collapse(local, comm, "min", local.datatype)


def pmax(local: PDL, comm: Comm) -> PDL:
    return _collapse(local, comm, "max", local.datatype)


def pmin_ind(local: PDL, comm: Comm) -> PDL:
    return _collapse(local, comm, "min_ind", local.datatype)


def pmax_ind(local: PDL, comm: Comm) -> PDL:
    return _collapse(local, comm, "max_ind", local.datatype)


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


__all__ = ["split_dim", "shard", "Comm", "psum", "pavg", "pmin", "pmax", "pmin_ind", "pmax_ind", "pminmax", "pinner"]
