"""Parity spot-check of the CUDA path vs the oracle on a synthetic workload. usage: gpu_diag2.py alphabet n_total n_backbone root_len nq [frag_frac]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth
import witch_b200 as wb
from oracle import oracle as O
alph, nt, nb, rl, nq = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
ff = float(sys.argv[6]) if len(sys.argv) > 6 else 0.3
kw = dict(alphabet=alph, n_total=nt, n_backbone=nb, root_len=rl, decomp=10, frag_frac=ff, frag_mean=rl // 3, seed=21)
if alph == "amino": kw.update(mean_blen=0.05, indel_rate=0.004)
wl = synth.make_workload("/tmp/witch_b200_diag", **kw)
seqs = wl["seqs"][:nq]
E = wb.EHMM(wl["hmm_paths"]); Q = wb.Queries(E, seqs)
print("H", E.n, "M", E.M.min(), E.M.max(), "L", Q.lengths.min(), Q.lengths.max())
sc, rep, pre, fl = wb.score(E, Q)
print("reported %d/%d, non-finite among reported %d, score range %.1f..%.1f" % (rep.sum(), rep.size, (rep & ~np.isfinite(sc)).sum(), np.nanmin(sc), np.nanmax(sc)))
rng = np.random.default_rng(0)
profs = {}
def P(h):
    if h not in profs: profs[h] = O.Profile(wl["hmm_paths"][h])
    return profs[h]
md = mdp = 0; nrep = 0; n = 0
pairs = [(int(rng.integers(0, Q.n)), int(rng.integers(0, E.n))) for _ in range(150)]
for q, h in pairs:
    r = O.score_pair(P(h), P(h).abc.digitize(seqs[q])); n += 1
    if r["reported"] != bool(rep[q, h]): nrep += 1; print("  REP mismatch", q, h, r["reported"], rep[q, h], r["max_mocc"]); continue
    mdp = max(mdp, abs(pre[q, h] - r["pre_score"]))
    if r["reported"]:
        d = abs(sc[q, h] - r["score"]); md = max(md, d)
        if d > 0.01: print("  SCORE q%d h%d L%d M%d gpu %.4f oracle %.4f (pre %.4f/%.4f) flags %d/%d" % (q, h, len(seqs[q]), E.M[h], sc[q, h], r["score"], pre[q, h], r["pre_score"], fl[q, h], r["flags"]))
print("spot-check %d pairs: max |dscore| %.2e, max |dpre| %.2e, reported mismatches %d" % (n, md, mdp, nrep))
idx, w, cnt = wb.weights_topk(E, sc, rep, 10, 1)
aq = np.array([q for q in range(min(Q.n, 60)) for j in range(min(2, cnt[q]))], dtype=np.int32)
ah = np.array([idx[q, j] for q in range(min(Q.n, 60)) for j in range(min(2, cnt[q]))], dtype=np.int32)
cols = wb.align(E, Q, aq, ah)
nres = nbad = 0
for c, q, h in zip(cols, aq, ah):
    ref = O.align_pair(P(int(h)), P(int(h)).abc.digitize(seqs[q]))
    nres += len(ref); b = int((ref != c).sum()); nbad += b
    if b: print("  ALIGN q%d h%d L%d M%d mismatches %d" % (q, h, len(ref), E.M[h], b))
print("align: %d mismatching residues of %d (%d pairs)" % (nbad, nres, len(aq)))
