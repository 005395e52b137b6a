"""Compare two score dumps written by tools/gpu_perf_c2.py. usage: python tools/cmp_scores.py a.npz b.npz"""
import sys
import numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
m = a["rep"].astype(bool) & b["rep"].astype(bool)
print("reported identical:", np.array_equal(a["rep"], b["rep"]), " max|dscore| %.3e  max|dpre| %.3e  flags differ %d  nan mismatch %d" % (
    np.abs(a["sc"][m] - b["sc"][m]).max(), np.abs(a["pre"] - b["pre"]).max(), int((a["fl"] != b["fl"]).sum()),
    int((np.isnan(a["sc"]) != np.isnan(b["sc"])).sum())))
