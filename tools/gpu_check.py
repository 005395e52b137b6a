"""Quick GPU-side parity check against the oracle (no torch import): run under gpurun.
usage: python tools/gpu_check.py [set ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import witch_b200 as wb  # noqa: E402
from golden_util import load_set  # noqa: E402
from oracle import oracle as O  # noqa: E402

sets = sys.argv[1:] or ["dna_small", "amino_small", "dna_sub8", "dna_full"]
for s in sets:
    gold, queries, paths = load_set(s)
    profs = [O.Profile(p) for p in paths]
    t0 = time.time()
    E = wb.EHMM(paths)
    Q = wb.Queries(E, [q for _, q in queries])
    print(s, "H", E.n, "M", E.M, "nq", Q.n, "setup %.2fs" % (time.time() - t0))
    nq, H = Q.n, E.n
    qi = np.repeat(np.arange(nq), H).astype(np.int32)
    hi = np.tile(np.arange(H), nq).astype(np.int32)
    for mh in (True, False):
        f, b = wb.debug_fwdbwd(E, Q, qi, hi, mh)
        of = np.array([O.forward_nats(profs[h], profs[h].abc.digitize(queries[q][1]), mh) for q, h in zip(qi, hi)])
        df, db = np.abs(f - of), np.abs(b - of)
        print("  %s fwd maxdiff %.3e  bwd maxdiff %.3e (nats)  nan: %d %d" % (
            "multihit" if mh else "unihit ", np.nanmax(df), np.nanmax(db), np.isnan(f).sum(), np.isnan(b).sum()))
        if np.nanmax(df) > 1e-2 or np.isnan(f).any():
            bad = np.argsort(-np.nan_to_num(df, nan=1e9))[:5]
            for z in bad:
                print("     pair q%d h%d L%d gpu %.4f oracle %.4f bwd %.4f" % (qi[z], hi[z], len(queries[qi[z]][1]), f[z], of[z], b[z]))
    t0 = time.time()
    sc, rep, pre, fl = wb.score(E, Q)
    print("  score call %.3fs" % (time.time() - t0))
    nbad = nrep = 0
    maxd = maxdp = 0.0
    for q in range(nq):
        for h in range(H):
            r = O.score_pair(profs[h], profs[h].abc.digitize(queries[q][1]))
            if r["reported"] != bool(rep[q, h]):
                nrep += 1
                if nrep <= 5:
                    print("     REPORT mismatch q%d h%d oracle %s gpu %s maxmocc %.3f flags %d/%d" % (q, h, r["reported"], rep[q, h], r["max_mocc"], r["flags"], fl[q, h]))
                continue
            maxdp = max(maxdp, abs(pre[q, h] - r["pre_score"]))
            if r["reported"]:
                d = abs(sc[q, h] - r["score"])
                maxd = max(maxd, d)
                if d > 0.01:
                    nbad += 1
                    if nbad <= 5:
                        print("     SCORE q%d h%d gpu %.4f oracle %.4f pre %.4f/%.4f flags %d/%d env %s" % (q, h, sc[q, h], r["score"], pre[q, h], r["pre_score"], fl[q, h], r["flags"], r["env"]))
    print("  scores: max|d| %.2e bits (pre %.2e), >0.01: %d, reported mismatches: %d of %d" % (maxd, maxdp, nbad, nrep, nq * H))
    # weights
    idx, w, cnt = wb.weights_topk(E, sc, rep, 10, 1)
    nw = 0
    for q in range(nq):
        ss = {h: O.printed_score(float(sc[q, h])) for h in range(H) if rep[q, h]}
        if not ss:
            assert cnt[q] == 0
            continue
        ranked = O.rank_bitscores(ss)
        ow = O.calculate_weights([h for h, _ in ranked], [x for _, x in ranked], [int(E.nseq[h]) for h, _ in ranked], 10)
        got = [(int(idx[q, j]), float(w[q, j])) for j in range(cnt[q])]
        ok = len(got) == len(ow) and all(abs(a[1] - b[1]) <= 1e-12 * max(1e-300, b[1]) + 1e-300 for a, b in zip(got, ow)) \
            and sorted(a[0] for a in got) == sorted(b[0] for b in ow)
        if not ok:
            nw += 1
            if nw <= 3:
                print("     WEIGHTS q%d gpu %s oracle %s" % (q, got, ow))
    print("  weights mismatches: %d of %d" % (nw, nq))
    # align
    names = [n for n, _ in queries]
    pq, ph, exp = [], [], []
    for h, hg in enumerate(gold["hmms"]):
        for n, cols in hg["columns"].items():
            pq.append(names.index(n)); ph.append(h); exp.append(np.array(cols, dtype=np.int32))
    t0 = time.time()
    got = wb.align(E, Q, pq, ph)
    print("  align call %.3fs for %d pairs" % (time.time() - t0, len(pq)))
    nres = nmis = npair = 0
    for a, b, q, h in zip(got, exp, pq, ph):
        m = int((a != b).sum())
        nres += len(b); nmis += m
        if m:
            npair += 1
            if npair <= 4:
                w_ = np.nonzero(a != b)[0]
                print("     ALIGN q%d h%d mismatches %d/%d first at %d: gpu %s ref %s" % (q, h, m, len(b), w_[0], a[w_[0]:w_[0] + 6], b[w_[0]:w_[0] + 6]))
    print("  align: %d mismatching residues of %d (%d pairs of %d)" % (nmis, nres, npair, len(pq)))
print("launches", wb.kernel_launches())
