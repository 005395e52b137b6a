"""Rough throughput probe (no torch): replicated golden inputs. usage: python tools/gpu_perf.py [nrep_hmm] [nrep_q]"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import witch_b200 as wb  # noqa: E402
from witch_b200 import _lib  # noqa: E402
from golden_util import load_set  # noqa: E402

nh = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nqr = int(sys.argv[2]) if len(sys.argv) > 2 else 12
sets = sys.argv[3:] or ["dna_sub8"]
lib = _lib.load()
for s in sets:
    gold, queries, paths = load_set(s)
    E = wb.EHMM([paths[0]] * nh)
    Q = wb.Queries(E, [q for _, q in queries] * nqr)
    cells = float(Q.lengths.sum()) * float(E.M.sum())
    lib.witch_prof_enable(1)
    for it in range(3):
        lib.witch_prof_reset()
        t0 = time.time()
        sc, rep, pre, fl = wb.score(E, Q)
        dt = time.time() - t0
        c = ctypes.c_double(); n = ctypes.c_uint64()
        ms0 = lib.witch_prof_get(0, ctypes.byref(c), ctypes.byref(n)); c0 = c.value
        ms1 = lib.witch_prof_get(1, ctypes.byref(c), ctypes.byref(n)); c1 = c.value
        print("%s H=%d nq=%d cells=%.3g  score wall %.3fs (%.1f Gcell/s)  parser %.1f ms (%.1f Gcell/s)  env %.1f ms (%.1f Gcell/s of %.3g)  reported %d" % (
            s, E.n, Q.n, cells, dt, cells / dt / 1e9, ms0, c0 / ms0 / 1e6, ms1, c1 / max(ms1, 1e-9) / 1e6, c1, rep.sum()))
    idx, w, cnt = wb.weights_topk(E, sc, rep, 10, 1)
    pq = np.repeat(np.arange(Q.n), 3).astype(np.int32)
    ph = np.tile(np.arange(3) % E.n, Q.n).astype(np.int32)
    lib.witch_prof_reset()
    t0 = time.time()
    cols = wb.align(E, Q, pq, ph)
    dt = time.time() - t0
    c = ctypes.c_double(); n = ctypes.c_uint64()
    ms2 = lib.witch_prof_get(2, ctypes.byref(c), ctypes.byref(n))
    print("   align %d pairs wall %.3fs kernel %.1f ms (%.1f Gcell/s)" % (len(pq), dt, ms2, c.value / ms2 / 1e6))
