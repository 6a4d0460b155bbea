#!/usr/bin/env python
"""profiles/<op>_r1.ncu_raw.csv (tools/ncu_summary.sh) -> profiles/traffic.json + a markdown table.
traffic = dram__bytes_read.sum + dram__bytes_write.sum of ONE launch; bench.py reads roofline.traffic from it.
Usage: python tools/make_traffic.py [profiles_dir] [suffix]"""
import csv
import json
import sys
from pathlib import Path

d = Path(sys.argv[1] if len(sys.argv) > 1 else "profiles")
suffix = sys.argv[2] if len(sys.argv) > 2 else "_r1.ncu_raw.csv"
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
        "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}
KEYS = {"minimum": "reduce_minimum", "sumover": "reduce_sumover", "average": "reduce_average", "plus": "ew_plus_cfg1",
        "mult_cfg3": "ew_mult_cfg3"}


def num(v):
    return float(v.replace(",", ""))


out, detail, table = {}, {}, []
for f in sorted(d.glob("*" + suffix)):
    op = f.name[: -len(suffix)]
    m = {}
    for row in csv.reader(f.open()):
        if len(row) >= 3:
            m[row[0]] = (row[1], ",".join(row[2:]))
    if "dram__bytes_read.sum" not in m:
        continue

    def val(k, default=0.0):
        if k not in m:
            return default
        u, v = m[k]
        try:
            return num(v) * UNIT.get(u, 1)
        except ValueError:
            return default
    rd, wr, dur = val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("gpu__time_duration.sum")
    detail[op] = {"kernel": m.get("Kernel Name", ("", ""))[1], "dram_bytes_read": rd, "dram_bytes_write": wr,
                  "traffic": rd + wr, "duration_s": dur, "dram_gbs_under_ncu": (rd + wr) / dur / 1e9 if dur else None,
                  "registers": val("launch__registers_per_thread"),
                  "alu_pipe_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                  "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active")}
    out[KEYS.get(op, op)] = rd + wr
    table.append(f"| {op} | `{detail[op]['kernel'][:70]}` | {dur * 1e6:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
                 f"{(rd + wr) / dur / 1e9 if dur else 0:.0f} | {detail[op]['registers']:.0f} | {detail[op]['alu_pipe_pct']:.0f} | "
                 f"{detail[op]['warps_active_pct']:.0f} |")
out["detail"] = detail
(d / "traffic.json").write_text(json.dumps(out, indent=1))
print("| op | kernel | duration us | DRAM read MB | DRAM write MB | DRAM GB/s | regs | ALU pipe % | warps active % |")
print("|---|---|---|---|---|---|---|---|---|")
print("\n".join(table))
