"""profiles/r01_traffic.json from `ncu --page raw --csv` exports of the dominant launch of each kernel family, captured
on the (truncated) bench command: DRAM bytes per DP cell = (dram__bytes_read + dram__bytes_write of the launch) / cells
of that launch, where the launch's cells = family cells of the bench line x the launch's share of the family's time in
the ncu launch list of the same command.
usage: make_traffic_json.py bench.json launches.csv out.json parser=raw.csv wave_env=raw.csv wave_align=raw.csv"""
import csv
import json
import sys

bench, launches, out = sys.argv[1:4]
raws = dict(a.split("=", 1) for a in sys.argv[4:])
MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def fam_of(name):
    if "mh_parser" in name:
        return "parser"
    if "wave_kernel" in name:
        return "wave_align" if "(bool)1" in name.split(",")[1] or ", 1," in name.split("(")[0] + name else "wave_env"
    return None


# launch list: per family, the duration of every launch
fam_times = {"parser": [], "wave_env": [], "wave_align": []}
rows = [r for r in csv.reader(open(launches)) if len(r) > 6]
hdr = rows[0]
ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    n = r[ki]
    f = "parser" if "mh_parser" in n else ("wave_align" if "wave_kernel<8, 1" in n or "wave_kernel<(int)8, (bool)1" in n else
                                          ("wave_env" if "wave_kernel" in n else None))
    if f:
        fam_times[f].append(float(r[vi].replace(",", "")) * MULT.get(r[ui], 1.0))
b = json.load(open(bench))
kk = list(b["roofline"]["kernels"].values())
cells = {"parser": kk[0]["cells"], "wave_env": kk[1]["cells"], "wave_align": kk[2]["cells"]}
res = {}
for f, path in raws.items():
    rr = list(csv.reader(open(path)))
    h, u, v = rr[0], rr[1], rr[2]
    ix = {x: i for i, x in enumerate(h)}
    get = lambda k: float(v[ix[k]].replace(",", "")) * MULT.get(u[ix[k]], 1.0)
    ms = get("gpu__time_duration.sum")
    if "dram__bytes_read.sum" in ix:
        dram = get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
    else:   # section sets without the explicit byte counters: bytes = average rate x duration
        rate_unit = u[ix["dram__bytes.sum.per_second"]]                       # e.g. "Tbyte/s"
        rate = float(v[ix["dram__bytes.sum.per_second"]].replace(",", "")) * MULT[rate_unit.split("/")[0]]
        dram = rate * ms * 1e-3
    # the bench command runs the step twice (timed leg + end-to-end leg): the launch list holds both, the bench line's
    # cells are for one; within a family cells are taken proportional to launch time
    per_step = sum(fam_times[f]) / 2.0 if fam_times[f] else ms
    share = min(1.0, ms / per_step)
    c = cells[f] * share
    res[f] = {"captured_launch_ms": ms, "family_launches": len(fam_times[f]), "share_of_family_step_time": share,
              "dram_bytes": dram, "cells": c, "dram_bytes_per_cell": dram / max(c, 1.0)}
json.dump({"source": "ncu (sections SpeedOfLight, MemoryWorkloadAnalysis, ...) --clock-control none on: python bench.py --max-queries 640 "
                     "--max-hmms 48 --steps 1 --warmup 0 --no-cpu-baseline; longest launch of each kernel family",
           "kernels": res,
           "note": "algorithmic bytes/cell: parser 0 (registers + L2 scratch), wave_env 8 (Forward match row written once, read "
                   "once), wave_align 33"}, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
