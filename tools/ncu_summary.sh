#!/usr/bin/env bash
# Run on the GPU box: one `ncu --set full` capture per op, exported as small CSV summaries
# (the .ncu-rep files with imported source are too big to bring back).
#   tools/ncu_summary.sh <outdir> op1 op2 ...
set -u
OUT="$1"; shift
mkdir -p "$OUT"
K="regex:reduce_rows_kernel|ew_tile_kernel|ew_kernel|mm_dmma|mm_exact_kernel|inner_warp_kernel|minmaximum_warp_kernel|scan_chunk_kernel|axisvals_kernel|nind_kernel|mm_dmma_tma_kernel|collapse_records_kernel|scan_onepass_kernel"
for op in "$@"; do
  REP=/tmp/prof_$(echo $op | tr ':' '_')
  ncu --set full --clock-control none --import-source on -k "$K" -s 1 -c 1 -f -o $REP python tools/prof_one.py $op 3 > /dev/null 2>&1
  ncu -i $REP.ncu-rep --page raw --csv 2>/dev/null | python3 -c '
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    hdr, units, vals = rows[0], rows[1], rows[2]
    keep = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput",
            "gpu__dram_throughput", "sm__throughput", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__warps_active", "launch__occupancy", "smsp__inst_executed.sum", "sm__inst_executed", "l1tex__t_bytes", "lts__t_bytes",
            "smsp__cycles_active.avg", "sm__pipe", "smsp__issue_active", "smsp__average_warp", "dram__cycles_active",
            "smsp__warp_issue_stalled", "l1tex__data_pipe", "lts__t_sectors", "sm__inst_executed_pipe")
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(k) for k in keep):
            print(f"{h},{u},{v}")
' > "$OUT/$(echo $op | tr ':' '_').raw.csv"
  # hottest source lines by sampled stalls
  ncu -i $REP.ncu-rep --page source --csv 2>/dev/null | python3 -c '
import csv, sys
rows = list(csv.reader(sys.stdin))
if rows:
    hdr = rows[0]
    try:
        si = hdr.index("# Samples")
    except ValueError:
        si = next((i for i, h in enumerate(hdr) if "Samples" in h), None)
    src = next((i for i, h in enumerate(hdr) if h in ("Source", "Source Line")), 1)
    if si is not None:
        body = [r for r in rows[1:] if len(r) > si and r[si].replace(".", "", 1).isdigit()]
        body.sort(key=lambda r: -float(r[si]))
        print(",".join(hdr[:8]))
        for r in body[:40]:
            print(",".join(x.replace(",", ";")[:160] for x in r[:8]))
' > "$OUT/$(echo $op | tr ':' '_').source_top.csv"
  rm -f $REP.ncu-rep
done
ls -la "$OUT"
