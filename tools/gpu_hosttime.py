"""Where the non-kernel time of one DevicePipeline.run goes (WITCH_TIMING=1 prints the library's host phases).
usage: WITCH_TIMING=1 python tools/gpu_hosttime.py [config=c2] [slabs=4]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import synth
import witch_b200 as wb
from witch_b200.gcmm import DevicePipeline
from witch_b200 import _lib
if os.environ.get("PERF_LIB"):
    _lib.LIB_PATH = os.path.join(ROOT, os.environ["PERF_LIB"])
nrep = int(os.environ.get("HOSTTIME_REPS", "3"))
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = synth.make_workload("/tmp/witch_b200_bench", **synth.CONFIGS[cfg])
L = np.array([len(s) for s in wl["seqs"]])
ids = np.sort(np.argsort(-L, kind="stable")[0::ns])
seqs = [wl["seqs"][i] for i in ids]
E = wb.EHMM(wl["hmm_paths"])
pipe = DevicePipeline(E, k=10)
for it in range(nrep):
    t0 = time.perf_counter()
    Q = wb.Queries(E, seqs)
    t1 = time.perf_counter()
    sys.stderr.write("---- run %d (queries upload %.1f ms)\n" % (it, 1e3 * (t1 - t0)))
    res = pipe.run(Q)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    sys.stderr.write("---- pipe.run total %.1f ms\n" % (1e3 * (t2 - t1)))
