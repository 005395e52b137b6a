"""Runs the library's CUDA sources on the CPU through the SIMT simulator (tools/sim/simt.h) and checks them against the
oracle -- TEST TOOLING for kernel development when no GPU is at hand; never part of the product path (it is the only
place that points the ctypes binding at tools/sim/libwitch_<name>.so, and it says so).

usage: python tools/sim/sim_check.py [lib name = sim] [golden set = dna_small] [n queries = 3] [align pairs = 2]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from witch_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sim"
setname = sys.argv[2] if len(sys.argv) > 2 else "dna_small"
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 3
nalign = int(sys.argv[4]) if len(sys.argv) > 4 else 2
_lib.LIB_PATH = os.path.join(ROOT, "tools", "sim", "libwitch_%s.so" % name)   # the SIMULATION build, explicitly
import witch_b200 as wb  # noqa: E402
from golden_util import load_set  # noqa: E402
from oracle import oracle as O  # noqa: E402

gold, queries, paths = load_set(setname)
queries = sorted(queries, key=lambda x: len(x[1]))[:nq] if os.environ.get("SIM_SHORTEST") else queries[:nq]
if os.environ.get("SIM_BIAS"):   # e.g. SIM_BIAS=0.5:T -- composition-biased copies: null2 corrections of several bits,
    frac, ch = os.environ["SIM_BIAS"].split(":")   # so the envelope kernel's posteriors / domain corrections really matter
    rng = np.random.default_rng(1)
    biased = []
    for n, s in queries:
        a = np.array(list(s.upper()))
        a[rng.random(len(a)) < float(frac)] = ch
        biased.append((n + "_biased", "".join(a)))
    queries = queries + biased
profs = [O.Profile(p) for p in paths]
E = wb.EHMM(paths)
Q = wb.Queries(E, [s for _, s in queries])
print("simulating %s: %d queries (%s nt) x %d HMMs (M = %s)" % (setname, Q.n, [len(s) for _, s in queries], E.n, list(E.M)))
t0 = time.time()
sc, rep, pre, fl = wb.score(E, Q)
print("score stage simulated in %.1f s" % (time.time() - t0))
worst = wpre = wbias = 0.0
nmd = 0
for qi in range(Q.n):
    for h in range(E.n):
        r = O.score_pair(profs[h], profs[h].abc.digitize(queries[qi][1]))
        assert bool(rep[qi, h]) == r["reported"], (qi, h, r, sc[qi, h])
        assert (int(fl[qi, h]) & 1) == (r["flags"] & 1), (qi, h, fl[qi, h], r)
        nmd += r["flags"] & 1
        wpre = max(wpre, abs(float(pre[qi, h]) - r["pre_score"]))
        if r["reported"]:
            worst = max(worst, abs(float(sc[qi, h]) - r["score"]))
            wbias = max(wbias, r["pre_score"] - r["score"])
print("parser + multi-domain + envelope kernels: reported sets identical, %d of %d pairs through the multi-domain branch, max |dpre| %.2e bits, "
      "max |dscore| %.2e bits (largest null2 correction %.2f bits)" % (nmd, Q.n * E.n, wpre, worst, wbias))
assert wpre < 1e-3 and worst < 1e-3
idx, w, cnt = wb.weights_topk(E, sc, rep, 10, 1)
pq = [qi for qi in range(Q.n) if cnt[qi] > 0][:nalign]
ph = [int(idx[qi, 0]) for qi in pq]
if pq:
    t0 = time.time()
    cols = wb.align(E, Q, pq, ph)
    nres = nbad = 0
    for qi, h, c in zip(pq, ph, cols):
        ref = O.align_pair(profs[h], profs[h].abc.digitize(queries[qi][1]))
        nres += len(ref); nbad += int((ref != c).sum())
    print("align kernel: %d residues, %d differ from the oracle (%.1f s)" % (nres, nbad, time.time() - t0))
    assert nbad <= max(1, nres // 2000)
print("SIM CHECK OK (%s)" % name)
