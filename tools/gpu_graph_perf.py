"""Time the device alignment-graph DP on a synthetic workload and spot-check it against the oracle.
usage: gpu_graph_perf.py config nq nh"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth
import witch_b200 as wb
from witch_b200.gcmm import BatchedSearch
from oracle import oracle as O
cfg, nq, nh = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
kw = dict(synth.CONFIGS[cfg]); kw["max_hmms"] = nh
wl = synth.make_workload("/tmp/witch_b200_bench", **kw)
seqs = wl["seqs"][:nq]; names = wl["names"][:nq]
bs = BatchedSearch(wl["hmm_paths"], num_hmms=10)
t0 = time.time(); bs.search(names, seqs); t1 = time.time()
t2w = bs.writeWeights(); t2 = time.time()
bb = bs.getBackbones(t2w); t3 = time.time()
ret = {h: wl["retained_columns"][h] for h in range(len(wl["hmm_paths"]))}
ng = {h: wl["nongaps_per_column"][h] for h in range(len(wl["hmm_paths"]))}
rows = bs.alignSubQueriesNew(wl["backbone_length"], ret, ng, t2w); t4 = time.time()
npairs = sum(len(v[2]) for v in bb.values() if v[0] != "N/A")
cells = sum(len(s) for s in seqs) * 1.0
print("search %.2fs weights %.2fs getBackbones %.2fs (%d pairs) alignSubQueriesNew(total incl. align) %.2fs for %d queries, backbone %d" % (
    t1 - t0, t2 - t1, t3 - t2, npairs, t4 - t3, len(seqs), wl["backbone_length"]))
rng = np.random.default_rng(1)
bad = 0
for q in rng.choice(len(seqs), size=min(12, len(seqs)), replace=False):
    t = names[q]
    _, wmap, s2c = bb[t]
    want = O.compress_insertions(O.graph_align(seqs[q].upper(), wl["backbone_length"], wmap, s2c, ret, ng))
    if rows[t] != want: bad += 1
print("oracle spot-check mismatches:", bad)
