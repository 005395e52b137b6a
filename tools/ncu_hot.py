"""Rank SASS instructions of one kernel launch by warp-stall samples (from `ncu --page source --csv --print-source sass`).
usage: python tools/ncu_hot.py source.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
print("kernels in file:", len(starts), "-> using", which, rows[starts[which] - 1][1][:80] if starts[which] > 0 else "")
hi = starts[which]
end = starts[which + 1] - 1 if which + 1 < len(starts) else len(rows)
rows = rows[:end]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
data = []
tot = 0
for k, r in enumerate(rows[hi + 1:]):
    if len(r) < len(h):
        continue
    s = int(r[col["# Samples"]] or 0)
    tot += s
    data.append((k, r, s))
print("total samples", tot, "instructions", len(data))
agg = {n: 0 for n in stalls}
for k, r, s in data:
    for n in stalls:
        agg[n] += int(r[col[n]] or 0)
print("stall mix:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / max(tot, 1)) for n, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
for k, r, s in sorted(data, key=lambda x: -x[2])[:top]:
    why = sorted(((int(r[col[n]] or 0), n[6:]) for n in stalls), reverse=True)[:2]
    print("%5d %5.2f%% exec %-10s %-60s %s" % (k, 100.0 * s / tot, r[col["Instructions Executed"]], r[col["Source"]].strip()[:60],
                                     " ".join("%s:%d" % (n, v) for v, n in why)))
