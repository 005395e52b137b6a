"""Diagnose the CUDA path on a (truncated) synthetic workload against the oracle. usage: gpu_diag.py config nq nh"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth
import witch_b200 as wb
from oracle import oracle as O
cfg, nq, nh = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
kw = dict(synth.CONFIGS[cfg]); kw["max_hmms"] = nh
wl = synth.make_workload("/tmp/witch_b200_bench", **kw)
seqs = wl["seqs"][:nq]
E = wb.EHMM(wl["hmm_paths"]); Q = wb.Queries(E, seqs)
print("M", E.M.min(), E.M.max(), "L", Q.lengths.min(), Q.lengths.max())
t0 = time.time(); sc, rep, pre, fl = wb.score(E, Q); print("score %.2fs" % (time.time() - t0))
bad = rep & ~np.isfinite(sc)
print("reported", rep.sum(), "of", rep.size, "non-finite among reported", bad.sum(), "pre non-finite", (~np.isfinite(pre)).sum())
for q, h in list(zip(*np.nonzero(bad)))[:6]:
    p = O.Profile(wl["hmm_paths"][h]); r = O.score_pair(p, p.abc.digitize(seqs[q]))
    print("  bad q%d h%d L%d M%d gpu sc %s pre %s flags %d | oracle %s" % (q, h, len(seqs[q]), E.M[h], sc[q, h], pre[q, h], fl[q, h], r))
rng = np.random.default_rng(0)
md = 0
for _ in range(12):
    q, h = int(rng.integers(0, Q.n)), int(rng.integers(0, E.n))
    p = O.Profile(wl["hmm_paths"][h]); r = O.score_pair(p, p.abc.digitize(seqs[q]))
    d = abs(sc[q, h] - r["score"]) if r["reported"] and rep[q, h] else (0 if r["reported"] == rep[q, h] else 99)
    md = max(md, d)
    if d > 0.01: print("  MISMATCH q%d h%d L%d M%d gpu %s/%s oracle %s" % (q, h, len(seqs[q]), E.M[h], sc[q, h], rep[q, h], r))
print("spot max diff", md)
idx, w, cnt = wb.weights_topk(E, sc, rep, 10, 1)
print("idx range", idx.min(), idx.max(), "cnt", cnt.min(), cnt.max())
bq, bh = np.nonzero(bad)
if len(bq):
    sel = np.arange(min(6, len(bq)))
    pq, ph = bq[sel].astype(np.int32), bh[sel].astype(np.int32)
    for mh in (True, False):
        f, b = wb.debug_fwdbwd(E, Q, pq, ph, mh)
        for z in range(len(pq)):
            p = O.Profile(wl["hmm_paths"][ph[z]])
            of = O.forward_nats(p, p.abc.digitize(seqs[pq[z]]), mh)
            print("  %s q%d h%d fwd %.4f bwd %.4f oracle %.4f" % ("multi" if mh else "uni", pq[z], ph[z], f[z], b[z], of))
