#!/usr/bin/env python
"""bench.py — throughput of the PDL broadcast-loop hot path on B200.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
  sumover / average / minimum along dim 0 of a [16384, 65536] float ndarray with 1% BAD.
One STEP = those three reductions over the whole 4 GiB ndarray (3 x 2^30 elements).
  value     : elements/sec, whole job, inputs resident in HBM (device data store).
  e2e       : same metric through the host-buffer path — pinned host ndarray -> H2D -> the
              three reductions -> D2H of the three result ndarrays, all inside the timed region.
  roofline  : slowest of the three kernels; achieved = algorithmic bytes / CUDA-event time.
  N > 1     : one process per GPU; the ndarray is partitioned along its outermost broadcast
              dim (rows), every rank reduces its own [16384, 65536] block, no data-path
              collective -> weak scaling.
  extra     : the other BASELINE.json configs under the same clock.  `cfg5` (every N): full-array sum
              and max of a float ndarray of 2^30 elements PER GPU (2^33 at N = 8) sharded along its only
              dim, through the product's sharded-reduction API (pdl_b200.parallel.pcollapse: device
              partial records -> ONE ncclAllGather -> device merge), input = SURVEY.md §8(d)'s
              ({-1,0,+1} counter hash + one planted maximum), verified against exact integer sums;
              `cfg5_strong`: the same with a fixed 2^30-element ndarray split N ways.  `cfg1`, `cfg3`,
              `cfg4`, `perl` (N = 1 only): tools/config_legs.py and perl/PDL-B200/bench_ops.pl.
  --impl reference : the UNMODIFIED reference (PDL built into oracle/_ref) on the host cores,
              all threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

REF_ROWS = 4096        # the ONE bounded CPU sample: 4096 of 65536 rows (256 MiB), in cpu_baseline and --impl reference
REF_WARMUP = 2         # at least this many untimed steps (page-in of the sample) in both places
N_DIM = 16384          # reduced dim (dim 0), contiguous
ROWS = 65536           # broadcast dim per GPU
WORKLOAD = "cfg2: sumover+average+minimum over dim 0 of float[16384,65536], 1% BAD"
OPS = ("sumover", "average", "minimum")
SEED = 0x5EED


# ---------------------------------------------------------------------------------------------
# synthetic input: counter hash of the flat index; identical in torch (device
# generation), numpy (oracle sample) and PDL (oracle/ref_bench.pl)
# ---------------------------------------------------------------------------------------------
def _hash_numpy(idx):
    import numpy as np
    z = idx.astype(np.uint64) + np.uint64(SEED)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def sample_numpy(row0: int, rows: int):
    """float32 [rows, N_DIM] block (numpy C order == PDL dims [N_DIM, rows]) + BAD as -FLT_MAX."""
    import numpy as np
    idx = (np.arange(rows * N_DIM, dtype=np.uint64) + np.uint64(row0 * N_DIM))
    z = _hash_numpy(idx)
    vals = ((z >> np.uint64(11)) % np.uint64(17)).astype(np.float32) - np.float32(8)
    bad = ((z >> np.uint64(40)) % np.uint64(100)) == 0
    vals[bad] = -np.finfo(np.float32).max
    return vals.reshape(rows, N_DIM)


def generate_device(torch, row0: int, rows: int, device):
    """Same block generated on the device in 2048-row chunks (int64 wrap-around arithmetic)."""
    out = torch.empty((rows, N_DIM), dtype=torch.float32, device=device)
    m1 = 0xBF58476D1CE4E5B9 - (1 << 64)
    m2 = 0x94D049BB133111EB - (1 << 64)

    def lsr(z, k):
        return (z >> k) & ((1 << (64 - k)) - 1)

    step = 2048
    for r in range(0, rows, step):
        rr = min(step, rows - r)
        idx = torch.arange(rr * N_DIM, dtype=torch.int64, device=device) + (row0 + r) * N_DIM
        z = idx + SEED
        z = (z ^ lsr(z, 30)) * m1
        z = (z ^ lsr(z, 27)) * m2
        z = z ^ lsr(z, 31)
        vals = (lsr(z, 11) % 17).to(torch.float32) - 8.0
        bad = (lsr(z, 40) % 100) == 0
        vals[bad] = -torch.finfo(torch.float32).max
        out[r:r + rr] = vals.view(rr, N_DIM)
    return out


def generate_cfg5(torch, g0: int, count: int, device, plant_at: int = -1, plant_value: float = 3.0):
    """SURVEY.md §8(d) cfg5 input for global flat indices [g0, g0 + count): x[i] in {-1, 0, +1} from the counter
    hash (partial sums stay far below 2^24, so float accumulation is exact in any order) and, if `plant_at`
    falls in the range, one planted maximum.  Returns (float32 tensor, exact int64 sum of it)."""
    out = torch.empty(count, dtype=torch.float32, device=device)
    m1 = 0xBF58476D1CE4E5B9 - (1 << 64)
    m2 = 0x94D049BB133111EB - (1 << 64)

    def lsr(z, k):
        return (z >> k) & ((1 << (64 - k)) - 1)

    total = torch.zeros((), dtype=torch.int64, device=device)
    step = 1 << 25
    for r in range(0, count, step):
        rr = min(step, count - r)
        z = torch.arange(rr, dtype=torch.int64, device=device) + (g0 + r + SEED)
        z = (z ^ lsr(z, 30)) * m1
        z = (z ^ lsr(z, 27)) * m2
        z = z ^ lsr(z, 31)
        v = lsr(z, 11) % 3 - 1
        total += v.sum()
        out[r:r + rr] = v.to(torch.float32)
    if g0 <= plant_at < g0 + count:
        total += int(plant_value) - int(out[plant_at - g0].item())
        out[plant_at - g0] = plant_value
    return out, total


def _bind_near_gpu(torch, local: int):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that host staging memory is
    allocated there.  Returns the node number, or None when the topology cannot be read."""
    try:
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = Path("/sys/bus/pci/devices") / dev
        node = int((base / "numa_node").read_text().strip())
        cpus = set()
        for part in (base / "local_cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
        return node if node >= 0 else None
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
def _clock_sampler(stop, samples):
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    dev = os.environ.get("LOCAL_RANK", "0")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", dev, f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            parts = [p.strip() for p in r.stdout.strip().split(",")]
            if len(parts) >= 6:
                samples.append(parts)
        except Exception:
            pass
        stop.wait(0.2)


def _clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
    sm = [int(s[0]) for s in samples if s[0].isdigit()]
    mx = [int(s[1]) for s in samples if s[1].isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for k, n in enumerate(names) if any(s[2 + k] == "Active" for s in samples)]
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons, "samples": len(samples)}


def _peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ref_cmd(rows, steps, warmup, threads):
    ref = ROOT / "oracle" / "_ref" / "blib"
    return ["perl", f"-I{ref / 'lib'}", f"-I{ref / 'arch'}", str(ROOT / "oracle" / "ref_bench.pl"),
            "--n", str(N_DIM), "--rows", str(rows), "--steps", str(steps), "--warmup", str(warmup),
            "--threads", str(threads)]


def _have_ref():
    return (ROOT / "oracle" / "_ref" / "blib" / "arch" / "auto" / "PDL" / "Ufunc" / "Ufunc.so").exists()


def run_reference_sample(rows, steps, warmup, threads):
    r = subprocess.run(_ref_cmd(rows, steps, warmup, threads), capture_output=True, text=True, timeout=1500)
    if r.returncode != 0:
        raise RuntimeError("reference run failed: " + r.stderr[-400:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def run_port_sample(rows, steps, warmup):
    """The C restatement (oracle/pdl_oracle.c), single thread, on the same sample."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import pdl_b200 as P
    from pdl_b200 import types as T, ufunc
    from oracle_engine import OracleEngine
    e = OracleEngine()
    a = P.PDL.from_numpy(sample_numpy(0, rows), T.F, e).set_badflag(True)
    tot = 0.0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for op in OPS:
            getattr(ufunc, op)(a)
        if it >= warmup:
            tot += time.perf_counter() - t0
    return {"ms_per_step": 1000 * tot / steps, "elements_per_sec": 3 * rows * N_DIM * steps / tot}


def run_perl_bench():
    """The reference-facing plugin call under real PDL: perl/PDL-B200/bench_ops.pl times `$x = $y + $c`,
    a 3-op chain and ->sumover/->average/->minimum through the UNCHANGED operator surface, first on the
    reference's CPU path, then with the PDL::B200 shim attached (same process, same ndarrays)."""
    ref = ROOT / "oracle" / "_ref" / "blib"
    shim = ROOT / "perl" / "PDL-B200" / "blib"
    if not (shim / "arch" / "auto" / "PDL" / "B200" / "B200.so").exists() or not _have_ref():
        return {"error": "PDL::B200 shim or oracle/_ref not built"}
    cmd = ["perl", f"-I{ref / 'lib'}", f"-I{ref / 'arch'}", f"-I{shim / 'lib'}", f"-I{shim / 'arch'}",
           str(ROOT / "perl" / "PDL-B200" / "bench_ops.pl"), "--reps", "50"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        return {"error": (r.stderr or r.stdout)[-400:]}
    out = json.loads(r.stdout.strip().splitlines()[-1])
    out["workload"] = "perl: the operator surface under real PDL 2.106, CPU path vs PDL::B200 shim (ms per op, wall clock)"
    return out


# ---------------------------------------------------------------------------------------------
def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    rows = REF_ROWS  # bounded sample: 4096 of 65536 rows = 256 MiB; reference time scales linearly in rows
    if _have_ref():
        res = run_reference_sample(rows, args.steps, max(args.warmup, REF_WARMUP), cores)
        kind, threads = "reference", int(res.get("autopthread_actual") or 0) or 1
        detail = f"PDL {res['pdl_version']} autopthread target {cores}, actual {res.get('autopthread_actual')}"
    else:
        res = run_port_sample(rows, args.steps, args.warmup)
        kind, threads, detail = "port", 1, "oracle/pdl_oracle.c (oracle/_ref not present)"
    v = res["elements_per_sec"]
    sample = f"{rows} of {ROWS} rows ({rows * N_DIM * 4 >> 20} MiB), 3 ops per step; {detail}"
    line = {
        "impl": "reference", "metric": "elements/sec", "value": v, "unit": "elements/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"] * (ROWS / rows),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "elements/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main_ours(args):
    import numpy as np
    import torch
    import pdl_b200 as P
    from pdl_b200 import types as T, ufunc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    eng = P.CudaEngine(local)
    P.set_default_engine(eng)

    # ---- resident input: this rank's block of rows [rank*ROWS, (rank+1)*ROWS) ----
    dev_in = generate_device(torch, rank * ROWS, ROWS, device)
    torch.cuda.synchronize()
    a = P.PDL(eng, eng.wrap(dev_in.data_ptr(), dev_in.numel() * 4, dev_in), T.F, [N_DIM, ROWS]).set_badflag(True)
    outs = {op: P.PDL.empty(T.F, [ROWS], eng) for op in OPS}
    for o in outs.values():
        o.badflag = True

    prepared = {op: P.prepare_op(op, [a], [outs[op]]) for op in OPS}   # descriptor cached: 1 C-ABI call per op

    def step():
        for op in OPS:
            prepared[op]()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region (device-resident): K steps, CUDA events on the launch stream ----
    stop, samples = threading.Event(), []
    th = threading.Thread(target=_clock_sampler, args=(stop, samples), daemon=True)
    th.start()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = eng.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)

    # per-kernel times (same region repeated per op so each op's launches are bracketed alone)
    per_op_ms = {}
    for op in OPS:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            prepared[op]()
        e1.record()
        torch.cuda.synchronize()
        per_op_ms[op] = e0.elapsed_time(e1) / args.steps
    # the timed region lasts ~20 ms, shorter than one nvidia-smi poll: keep the same kernels running
    # for ~1.5 s so the clock/throttle record covers sustained load as well as the timed steps
    t_end = time.perf_counter() + 1.5
    sustained_steps, s0 = 0, torch.cuda.Event(enable_timing=True)
    s1 = torch.cuda.Event(enable_timing=True)
    s0.record()
    while time.perf_counter() < t_end:
        for _ in range(20):
            step()
        sustained_steps += 20
        torch.cuda.synchronize()
    s1.record()
    torch.cuda.synchronize()
    sustained_ms = s0.elapsed_time(s1) / max(sustained_steps, 1)
    stop.set()
    th.join(timeout=2)

    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    elems_step = 3 * N_DIM * ROWS * world
    value = elems_step / (ms_step / 1e3)

    peak, peak_src = _peak_hbm()
    bytes_op = N_DIM * ROWS * 4 + ROWS * 4                      # 4 295 229 440 B per op
    per_op = {op: {"ms": per_op_ms[op], "gbs": bytes_op / per_op_ms[op] / 1e6,
                   "frac": bytes_op / per_op_ms[op] / 1e6 / peak,
                   "elements_per_sec": N_DIM * ROWS / (per_op_ms[op] / 1e3)} for op in OPS}
    dom = max(OPS, key=lambda o: per_op_ms[o])
    roofline = {"bound": "hbm", "kernel": f"reduce_{dom} (reduce_rows_kernel)", "achieved": per_op[dom]["gbs"],
                "peak": peak, "unit": "GB/s", "frac": per_op[dom]["frac"], "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_op}
    prof = ROOT / "profiles" / "traffic.json"
    if prof.exists():
        try:
            roofline["traffic"] = json.loads(prof.read_text()).get(f"reduce_{dom}")
        except Exception:
            pass

    # ---- verify a slice of the results against the oracle (checker only, outside timing) ----
    verified = None
    if rank == 0:
        sys.path.insert(0, str(ROOT / "oracle"))
        from oracle_engine import OracleEngine
        oe = OracleEngine()
        chk = 64
        host = sample_numpy(rank * ROWS, chk)
        pa = P.PDL.from_numpy(host, T.F, oe).set_badflag(True)
        verified = True
        for op in OPS:
            want = getattr(ufunc, op)(pa).to_numpy()
            got = outs[op].slice(f"0:{chk - 1}").to_numpy()
            verified = verified and got.tobytes() == want.tobytes()

    # ---- e2e: pinned host ndarray -> H2D -> 3 reductions -> D2H, per step ----
    e2e_steps = max(1, min(args.steps, 5))
    nbytes_in = N_DIM * ROWS * 4
    # staging memory near the GPU: run this rank on the CPUs of the GPU's NUMA node while the pinned buffer is
    # created (first touch places the pages there), and make it write-combined (the host only fills it, the GPU's
    # DMA reads it) — with 8 ranks pulling 4 GiB each, cross-socket traffic and snooping were what capped N = 8
    old_affinity = os.sched_getaffinity(0)
    numa = _bind_near_gpu(torch, local)
    host_ptr = eng.lib.pdlb200_host_alloc_wc(nbytes_in)
    staging = "write-combined pinned"
    if not host_ptr:
        host_ptr = eng.lib.pdlb200_host_alloc(nbytes_in)
        staging = "pinned"
    eng.download_ptr(a.store, host_ptr, nbytes_in)
    eng.sync()
    os.sched_setaffinity(0, old_affinity)      # only the allocation is placed; the CPU legs below use every core again
    host_out = torch.empty((3, ROWS), dtype=torch.float32, pin_memory=True)
    stage = torch.empty((ROWS, N_DIM), dtype=torch.float32, device=device)
    sa = P.PDL(eng, eng.wrap(stage.data_ptr(), nbytes_in, stage), T.F, [N_DIM, ROWS]).set_badflag(True)
    e2e_outs = {op: P.PDL.empty(T.F, [ROWS], eng) for op in OPS}
    copy_s, comp_s = torch.cuda.Stream(device), torch.cuda.Stream(device)
    SLABS = 16
    rows_slab = ROWS // SLABS
    slab_views = [(sa.slice(f":,{s * rows_slab}:{(s + 1) * rows_slab - 1}"),
                   {op: e2e_outs[op].slice(f"{s * rows_slab}:{(s + 1) * rows_slab - 1}") for op in OPS})
                  for s in range(SLABS)]
    evs = [torch.cuda.Event() for _ in range(SLABS)]

    def e2e_step():
        # upload slab s on the copy stream while slab s-1 is being reduced on the compute stream
        for s in range(SLABS):
            off = s * rows_slab * N_DIM * 4
            eng.stream = copy_s.cuda_stream
            eng.upload_ptr(sa.store, host_ptr + off, rows_slab * N_DIM * 4, off)
            evs[s].record(copy_s)
            comp_s.wait_event(evs[s])
            eng.stream = comp_s.cuda_stream
            view, o = slab_views[s]
            for op in OPS:
                P.run_op(op, [view], [o[op]])
        for k, op in enumerate(OPS):
            eng.download_ptr(e2e_outs[op].store, host_out.data_ptr() + k * ROWS * 4, ROWS * 4)
        comp_s.synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    eng.stream = None
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = elems_step * e2e_steps / float(t.item())
    e2e_ok = bool(np.array_equal(host_out[0].numpy().view(np.uint32), outs["sumover"].to_numpy().view(np.uint32)))

    # ---- cfg5 extra: full-array sum + max of a sharded 1-D float ndarray through the product's API ----
    from pdl_b200 import parallel

    class _Solo:                     # N = 1: same code path, the "gather" of one rank's records is a device copy
        rank, world, backend = 0, 1, "nccl"
        _bufs = {}
        record_buffers = parallel.Comm.record_buffers

        def all_gather_records(self, engine, lt, gt):
            gt.copy_(lt)

    comm = parallel.Comm() if dist is not None else _Solo()
    # records travel over peer memory (one kernel: NVLink stores + epoch flags) when the GPUs can map each other
    exchange = "device copy (1 rank)"
    if dist is not None:
        exchange = "peer-memory kernel (NVLink P2P, CUDA IPC mailboxes)" if (not args.nccl_gather and comm.enable_peer_exchange(eng)) \
            else "ncclAllGather"

    def cfg5_leg(per_gpu: int, label: str):
        total_n = per_gpu * world
        g0 = rank * per_gpu
        plant_at = (total_n * 5) // 7 + 11
        x, exact = generate_cfg5(torch, g0, per_gpu, device, plant_at, 3.0)
        if dist is not None:
            dist.all_reduce(exact, op=dist.ReduceOp.SUM)         # checker only: exact integer sum of the whole ndarray
        want_sum = int(exact.item())
        px = P.PDL(eng, eng.wrap(x.data_ptr(), per_gpu * 4, x), T.F, [per_gpu])
        res = [None]

        def step5():
            res[0] = parallel.pcollapse(px, comm, ("sum", "max"), offset=g0, total=total_n)

        for _ in range(3):
            step5()
        barrier()
        l0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step5()
        e1.record()
        barrier()
        nl = (eng.launch_count() - l0) // args.steps
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item()) / args.steps
        got_sum, got_max = res[0][0].to_numpy(), res[0][1].to_numpy()
        got_ind = int(parallel.pcollapse(px, comm, ("max_ind",), offset=g0, total=total_n)[0].sclr())
        ok = (abs(want_sum) < 2 ** 24 and got_sum.tobytes() == np.float32(want_sum).tobytes()
              and got_max.tobytes() == np.float32(3.0).tobytes() and got_ind == plant_at)
        same = torch.tensor([int(got_sum.view(np.uint32)), int(got_max.view(np.uint32)), got_ind], dtype=torch.int64, device=device)
        if dist is not None:                                     # identical bits on every rank
            lo, hi = same.clone(), same.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ok = ok and bool(torch.equal(lo, hi))
        del x
        return {"workload": label, "api": "pdl_b200.parallel.pcollapse(flat, comm, ('sum','max'))",
                "ms_per_step": ms, "elements_per_sec": 2 * total_n / (ms / 1e3),
                "gbs_per_gpu": 2 * per_gpu * 4 / ms / 1e6, "frac_per_gpu": 2 * per_gpu * 4 / ms / 1e6 / peak,
                "launches_per_step": int(nl), "exchanges_per_step": 1 if world > 1 else 0, "exchange": exchange,
                "sum": float(got_sum), "max": float(got_max), "max_ind": got_ind, "verified": bool(ok)}

    cfg5_extra = cfg5_leg(N_DIM * ROWS, f"cfg5: sum+max of float[2^30 x {world}] sharded over {world} GPU(s), weak")
    cfg5_strong = cfg5_leg(N_DIM * ROWS // world, f"cfg5 strong: sum+max of float[2^30] split over {world} GPU(s)")

    # ---- the other BASELINE configs and the Perl plugin call, N = 1 only ----
    extra = {"cfg5": cfg5_extra, "cfg5_strong": cfg5_strong}
    if rank == 0 and world == 1 and not args.no_extra:
        sys.path.insert(0, str(ROOT / "tools"))
        import config_legs
        for name, fn in (("cfg1", lambda: config_legs.cfg1(eng, device, peak)),
                         ("cfg3", lambda: config_legs.cfg3(eng, device, peak)),
                         ("cfg4", lambda: config_legs.cfg4(eng, device)),
                         ("next_rows", lambda: config_legs.next_rows(eng, device, peak)),
                         ("perl", run_perl_bench)):
            try:
                extra[name] = fn()
            except Exception as ex:  # an extra leg never takes the headline down with it
                extra[name] = {"error": f"{type(ex).__name__}: {ex}"[:400]}
            torch.cuda.empty_cache()

    # ---- CPU baseline on this box's host cores (rank 0, N == 1 only; bounded sample) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rows = REF_ROWS
        sample = f"{rows} of {ROWS} rows ({rows * N_DIM * 4 >> 20} MiB), 3 ops per step, {args.steps} timed steps"
        try:
            if _have_ref():
                r1 = run_reference_sample(rows, args.steps, max(args.warmup, REF_WARMUP), cores)
                r0 = run_reference_sample(rows, 3, 1, 0)
                cpu_baseline = {"value": r1["elements_per_sec"], "unit": "elements/s",
                                "cores": int(r1.get("autopthread_actual") or 0) or 1, "kind": "reference",
                                "sample": sample + f"; PDL {r1['pdl_version']} with autopthread ({r1.get('autopthread_actual')} threads)",
                                "no_pthread": {"value": r0["elements_per_sec"], "cores": 1},
                                "host_cores": cores}
            else:
                r = run_port_sample(rows, 3, 1)
                cpu_baseline = {"value": r["elements_per_sec"], "unit": "elements/s", "cores": 1, "kind": "port",
                                "sample": sample + "; oracle/pdl_oracle.c single thread", "host_cores": cores}
        except Exception as ex:  # the baseline is reported, never required for the GPU numbers
            cpu_baseline = {"value": None, "unit": "elements/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {
        "metric": "elements/sec", "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements_per_step_per_gpu": 3 * N_DIM * ROWS,
                   "partition": "rows (outermost broadcast dim) across ranks, no collective",
                   "l2": "inputs (4 GiB per GPU) larger than L2", "timer": "cuda events, max over ranks"},
        "roofline": roofline, "per_op": per_op,
        "e2e": {"value": e2e_value, "unit": "elements/s", "h2d_bytes_per_step": nbytes_in, "d2h_bytes_per_step": 3 * ROWS * 4,
                "steps": e2e_steps, "slabs": SLABS, "matches_resident_result": e2e_ok, "staging": staging, "numa_node": numa},
        "gpu_launches": int(launches),
        "clocks": dict(_clocks_summary(samples), window="timed steps + per-op loops + 1.5 s sustained loop of the same step"),
        "sustained": {"ms_per_step": sustained_ms, "steps": sustained_steps,
                      "value": elems_step / world / (sustained_ms / 1e3)},
        "verified_vs_oracle": verified,
        "cpu_baseline": cpu_baseline, "extra": extra,
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg1/cfg3/cfg4/perl legs")
    ap.add_argument("--nccl-gather", action="store_true", help="cfg5: exchange the records with ncclAllGather instead of the peer-memory kernel")
    a = ap.parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
