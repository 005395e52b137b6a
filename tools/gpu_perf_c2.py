"""Kernel-level throughput probe on a truncated c2 workload (no torch). Dumps scores so that two builds/variants can be
compared for numerical regressions.  usage: python tools/gpu_perf_c2.py [n_queries] [n_hmms] [tag] [config]"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth  # noqa: E402
import witch_b200 as wb  # noqa: E402
from witch_b200 import _lib  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 640
nh = int(sys.argv[2]) if len(sys.argv) > 2 else 48
tag = sys.argv[3] if len(sys.argv) > 3 else "base"
cfg = sys.argv[4] if len(sys.argv) > 4 else "c2"
kw = dict(synth.CONFIGS[cfg])
kw["max_hmms"] = nh
wl = synth.make_workload("/tmp/witch_b200_bench", **kw)
rng = np.random.default_rng(0)
sel = rng.permutation(len(wl["seqs"]))[:nq]
seqs = [wl["seqs"][i] for i in sel]
if os.environ.get("PERF_LIB"):
    _lib.LIB_PATH = os.path.join(ROOT, os.environ["PERF_LIB"])
lib = _lib.load()
hp = wl["hmm_paths"]
if os.environ.get("PERF_SKIP_ROOT"):
    hp = hp[1:]
E = wb.EHMM(hp)
Q = wb.Queries(E, seqs)
cells = float(Q.lengths.sum()) * float(E.M.sum())
lib.witch_prof_enable(1)
best = [1e30, 1e30]
for it in range(3):
    lib.witch_prof_reset()
    t0 = time.time()
    sc, rep, pre, fl = wb.score(E, Q)
    dt = time.time() - t0
    c = ctypes.c_double(); n = ctypes.c_uint64()
    ms0 = lib.witch_prof_get(0, ctypes.byref(c), ctypes.byref(n)); c0 = c.value
    ms1 = lib.witch_prof_get(1, ctypes.byref(c), ctypes.byref(n)); c1 = c.value
    best = [min(best[0], ms0), min(best[1], ms1)]
print("[%s] %s nq=%d H=%d cells=%.3g wall %.3fs | parser %.1f ms = %.1f Gcell/s | env %.1f ms = %.1f Gcell/s (%.3g cells) | reported %d, flagged %d" % (
    tag, cfg, Q.n, E.n, cells, dt, best[0], c0 / best[0] / 1e6, best[1], c1 / max(best[1], 1e-9) / 1e6, c1, int(rep.sum()), int((fl & 1).sum())))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "scores_%s_%s_%d_%d.npz" % (cfg, tag, nq, nh)), sc=sc, rep=rep, pre=pre, fl=fl)
ref = os.path.join(ROOT, "gpurun_out", "scores_%s_base_%d_%d.npz" % (cfg, nq, nh))
if tag != "base" and os.path.exists(ref):
    r = np.load(ref)
    same = np.array_equal(r["rep"], rep)
    m = rep.astype(bool) & r["rep"].astype(bool)
    print("   vs base: reported identical %s, max|dscore| %.2e, max|dpre| %.2e, flags differ %d" % (
        same, np.abs(sc[m] - r["sc"][m]).max(), np.abs(pre - r["pre"]).max(), int((fl != r["fl"]).sum())))
