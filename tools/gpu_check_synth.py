"""GPU-vs-oracle parity on a truncated synthetic workload of a named shape.
usage: python tools/gpu_check_synth.py CONFIG [n_queries] [n_hmms] [n_sample_pairs]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth  # noqa: E402
import witch_b200 as wb  # noqa: E402
from oracle import oracle as O  # noqa: E402

cfg = sys.argv[1]
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 200
nh = int(sys.argv[3]) if len(sys.argv) > 3 else 24
ns = int(sys.argv[4]) if len(sys.argv) > 4 else 150
kw = dict(synth.CONFIGS[cfg]); kw["max_hmms"] = nh
wl = synth.make_workload("/tmp/witch_b200_bench", **kw)
rng = np.random.default_rng(1)
sel = rng.permutation(len(wl["seqs"]))[:nq]
seqs = [wl["seqs"][i] for i in sel]
E = wb.EHMM(wl["hmm_paths"]); Q = wb.Queries(E, seqs)
t0 = time.time(); sc, rep, pre, fl = wb.score(E, Q); t1 = time.time() - t0
print("%s: nq=%d H=%d M=%d..%d L=%d..%d score call %.2fs reported %d flagged %d nan-in-reported %d" % (
    cfg, Q.n, E.n, E.M.min(), E.M.max(), Q.lengths.min(), Q.lengths.max(), t1, rep.sum(), (fl & 1).sum(), int(np.isnan(sc[rep.astype(bool)]).sum())))
idx, w, cnt = wb.weights_topk(E, sc, rep, 10, 1)
profs = {}
pq = rng.integers(0, Q.n, ns); ph = rng.integers(0, E.n, ns)
# bias the sample towards each query's best HMM (where alignment matters)
for z in range(0, ns, 2):
    if cnt[pq[z]] > 0:
        ph[z] = idx[pq[z], 0]
worst = worstp = 0.0; nrep = 0
for q, h in zip(pq, ph):
    if h not in profs:
        profs[h] = O.Profile(wl["hmm_paths"][h])
    r = O.score_pair(profs[h], profs[h].abc.digitize(seqs[q]))
    if r["reported"] != bool(rep[q, h]):
        nrep += 1; print("  REPORT mismatch q%d h%d oracle %s" % (q, h, r)); continue
    worstp = max(worstp, abs(pre[q, h] - r["pre_score"]))
    if r["reported"]:
        d = abs(sc[q, h] - r["score"]); worst = max(worst, d)
        if d > 0.01:
            print("  SCORE q%d h%d L%d gpu %.4f oracle %.4f (pre %.4f / %.4f) flags %d env %s" % (q, h, len(seqs[q]), sc[q, h], r["score"], pre[q, h], r["pre_score"], fl[q, h], r["env"]))
print("  scores vs oracle on %d pairs: max|d| %.2e bits (pre %.2e), reported mismatches %d" % (ns, worst, worstp, nrep))
aq = np.array([q for q in pq[: ns // 2] if cnt[q] > 0], dtype=np.int32)
ah = np.array([idx[q, 0] for q in aq], dtype=np.int32)
cols = wb.align(E, Q, aq, ah)
nres = nbad = 0
for c, q, h in zip(cols, aq, ah):
    if h not in profs:
        profs[h] = O.Profile(wl["hmm_paths"][h])
    ref = O.align_pair(profs[h], profs[h].abc.digitize(seqs[q]))
    nres += len(ref); nbad += int((ref != c).sum())
print("  align vs oracle: %d mismatching residues of %d (%d pairs)" % (nbad, nres, len(aq)))
