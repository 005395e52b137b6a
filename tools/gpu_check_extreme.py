"""Adversarial dynamic-range check: profiles whose consensus is rich in rare residues (W/C: ~5.5 bits per matched
row) so that 32 consecutive rows of a wavefront span > 150 bits. usage: python tools/gpu_check_extreme.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth  # noqa: E402
import witch_b200 as wb  # noqa: E402
from oracle import oracle as O  # noqa: E402

rng = np.random.default_rng(3)
td = tempfile.mkdtemp(prefix="witch_ext_")
A = synth.AMINO
paths, seqs = [], []
for n, (M, rich) in enumerate([(300, 120), (700, 400), (1200, 1200)]):
    cons = rng.integers(0, 20, M)
    s0 = int(rng.integers(0, M - rich + 1))
    cons[s0:s0 + rich] = rng.choice([A.index("W"), A.index("C"), A.index("H"), A.index("M")], rich, p=[0.6, 0.3, 0.05, 0.05])
    counts = np.zeros((M, 20)); counts[np.arange(M), cons] = 1.0
    tc = np.zeros((M + 1, 4)); tc[:, 0] = 1.0
    p = os.path.join(td, "ext_%d.hmm" % n)
    synth.write_hmm(p, "ext_%d" % n, counts, tc, 1, A)
    paths.append(p)
    full = "".join(A[c] for c in cons)
    seqs += [full, full[s0:s0 + rich], full[: M // 2], full[M // 3:], full[s0:s0 + rich][:40] + "A" * 30 + full[s0:s0 + rich][40:]]
E = wb.EHMM(paths); Q = wb.Queries(E, seqs)
sc, rep, pre, fl = wb.score(E, Q)
worst = 0.0; bad = 0
for q in range(Q.n):
    for h in range(E.n):
        p = O.Profile(paths[h]); r = O.score_pair(p, p.abc.digitize(seqs[q]))
        ok = r["reported"] == bool(rep[q, h])
        d = abs(sc[q, h] - r["score"]) if (ok and r["reported"]) else 0.0
        dp = abs(pre[q, h] - r["pre_score"])
        worst = max(worst, d, dp)
        if not ok or d > 0.01 or dp > 0.01 or np.isnan(pre[q, h]):
            bad += 1
            print("  q%d h%d L%d: gpu %.3f/%.3f rep %d  oracle %.3f/%.3f rep %d" % (q, h, len(seqs[q]), sc[q, h], pre[q, h], rep[q, h], r["score"] if r["reported"] else float("nan"), r["pre_score"], r["reported"]))
print("extreme: %d pairs, max|d| %.2e bits, bad %d; best score/residue %.2f bits" % (Q.n * E.n, worst, bad, np.nanmax(sc / Q.lengths[:, None])))
pq = [q for q in range(Q.n)]; ph = [q // 5 for q in range(Q.n)]
cols = wb.align(E, Q, pq, ph)
nres = nbad = 0
for c, q, h in zip(cols, pq, ph):
    p = O.Profile(paths[h]); ref = O.align_pair(p, p.abc.digitize(seqs[q]))
    nres += len(ref); nbad += int((ref != c).sum())
print("extreme align: %d mismatching residues of %d" % (nbad, nres))
import json, shutil
out = os.path.join(ROOT, "gpurun_out", "extreme")
os.makedirs(out, exist_ok=True)
for p_ in paths:
    shutil.copy(p_, out)
json.dump(dict(seqs=seqs, sc=np.nan_to_num(sc, nan=-1e30).tolist(), pre=pre.tolist(), rep=rep.tolist(), fl=fl.tolist(),
               cols=[c.tolist() for c in cols], pq=pq, ph=ph), open(os.path.join(out, "gpu.json"), "w"))
