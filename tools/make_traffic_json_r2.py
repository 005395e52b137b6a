"""profiles/r02_traffic.json: DRAM bytes per DP cell of the dominant launch of the parser and envelope kernels, from the
`ncu --set full` captures of the truncated bench command (raw CSV exports) and the launch list of the same command.
cells of the captured launch = family cells per step (bench line) x the launch's share of the family's time per step in
the launch list. The align kernel was not re-captured in round 2 (unchanged): its round-1 entry is carried over.
usage: make_traffic_json_r2.py bench.json launches.csv n_steps_in_launch_list out.json parser=raw.csv wave_env=raw.csv"""
import csv
import json
import os
import sys

bench, launches, nsteps, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raws = dict(a.split("=", 1) for a in sys.argv[5:])
MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def family(n):
    if "mh_parser" in n:
        return "parser"
    if "wave_kernel<8, 1" in n or "wave_kernel<(int)8, (bool)1" in n:
        return "wave_align"
    return "wave_env" if "wave_kernel" in n else None


fam_ms = {"parser": 0.0, "wave_env": 0.0, "wave_align": 0.0}
rows = [r for r in csv.reader(open(launches)) if len(r) > 6]
h = rows[0]
ki, mi, ui, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
for r in rows[1:]:
    f = family(r[ki])
    if f and r[mi] == "gpu__time_duration.sum":
        fam_ms[f] += float(r[vi].replace(",", "")) * MULT.get(r[ui], 1.0)
b = json.load(open(bench))
kk = list(b["roofline"]["kernels"].values())
cells_step = {"parser": kk[0]["cells"] / b["steps"], "wave_env": kk[1]["cells"] / b["steps"], "wave_align": kk[2]["cells"] / b["steps"]}
res = {}
for f, path in raws.items():
    rr = list(csv.reader(open(path)))
    hh, uu = rr[0], rr[1]
    ix = {x: i for i, x in enumerate(hh)}
    best = None
    for v in rr[2:]:
        get = lambda k: float(v[ix[k]].replace(",", "")) * MULT.get(uu[ix[k]], 1.0)
        ms = get("gpu__time_duration.sum")
        if best is None or ms > best[0]:
            best = (ms, get("dram__bytes_read.sum"), get("dram__bytes_write.sum"), v[ix["Kernel Name"]])
    ms, rd, wr, name = best
    share = min(1.0, ms / (fam_ms[f] / nsteps))
    c = cells_step[f] * share
    res[f] = {"kernel": name, "captured_launch_ms": ms, "share_of_family_step_time": share, "dram_bytes_read": rd, "dram_bytes_write": wr,
              "dram_bytes": rd + wr, "cells": c, "dram_bytes_per_cell": (rd + wr) / max(c, 1.0)}
try:
    r1 = json.load(open(os.path.join(os.path.dirname(out), "r01_traffic.json")))["kernels"]["wave_align"]
    r1["note"] = "carried over from profiles/r01_traffic.json (kernel unchanged, not re-captured in round 2)"
    res["wave_align"] = r1
except Exception:
    pass
json.dump({"source": "ncu --set full --clock-control none on: python bench.py --max-queries 640 --max-hmms 48 --slabs 1 --steps 2 --warmup 1 "
                     "--no-cpu-baseline; longest launch of each kernel family (tools/r2_gpu9.sh)",
           "kernels": res,
           "note": "algorithmic bytes/cell: parser 0 (registers + L2-resident special rows), wave_env 8 (Forward match row written once, "
                   "read once), wave_align 33. The parser's DRAM writes are its per-CTA special-state rows evicted from L2."},
          open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
