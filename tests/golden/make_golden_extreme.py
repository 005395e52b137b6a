"""Golden set `amino_extreme`: dynamic-range stress for the CUDA path, pinned by the reference's own HMMER binaries.

Profiles are written by hand (tools/synth.write_hmm: consensus rich in W/C, the rarest residues, so a matched row is
worth ~5.5 bits and whole-query scores reach thousands of bits -- far beyond what one FP32 or FP64 exponent holds);
the expected scores / envelopes / column lists come from hmmsearch / hmmalign 3.1b2 exactly as in make_golden.py.

Usage (build container only): python tests/golden/make_golden_extreme.py
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, HERE)
import synth  # noqa: E402
from make_golden import hmmsearch, hmmalign_columns, write_fasta  # noqa: E402
from oracle.make_ref import build as build_ref  # noqa: E402


def main():
    assert build_ref(), "reference binaries not staged"
    rng = np.random.default_rng(3)
    A = synth.AMINO
    d = os.path.join(HERE, "amino_extreme")
    os.makedirs(d, exist_ok=True)
    hmms, queries = [], []
    for n, (M, rich) in enumerate([(200, 90), (400, 260), (700, 700)]):
        cons = rng.integers(0, 20, M)
        s0 = int(rng.integers(0, M - rich + 1))
        cons[s0:s0 + rich] = rng.choice([A.index("W"), A.index("C"), A.index("H"), A.index("M")], rich, p=[0.6, 0.3, 0.05, 0.05])
        counts = np.zeros((M, 20)); counts[np.arange(M), cons] = 1.0
        tc = np.zeros((M + 1, 4)); tc[:, 0] = 1.0
        p = os.path.join(d, "hmm_%d.hmm" % n)
        synth.write_hmm(p, "ext_%d" % n, counts, tc, 1, A)
        hmms.append(p)
        full = "".join(A[c] for c in cons)
        rs = full[s0:s0 + rich]
        for k, s in enumerate([full, rs, full[: M // 2], full[M // 3:], rs[:40] + "A" * 30 + rs[40:]]):
            queries.append(("X%d_%d" % (n, k), s))
    write_fasta(os.path.join(d, "queries.fasta"), queries)
    gold = dict(molecule="amino", hmms=[])
    qd = dict(queries)
    for hi, hmm in enumerate(hmms):
        hits = hmmsearch(hmm, os.path.join(d, "queries.fasta"))
        own = [n for n, _ in queries if n.startswith("X%d_" % hi)]
        columns = {n: hmmalign_columns(hmm, n, qd[n]) for n in own}
        M = [int(ln.split()[1]) for ln in open(hmm) if ln.startswith("LENG")][0]
        with open(hmm, "rb") as f, gzip.GzipFile(hmm + ".gz", "wb", mtime=0) as g:
            g.write(f.read())
        os.remove(hmm)
        gold["hmms"].append(dict(file="hmm_%d.hmm.gz" % hi, M=M, nseq=1, retained_columns=list(range(M)),
                                 nongaps_per_column=[1] * M, hits=hits, columns=columns, taxa=["consensus"]))
        print("amino_extreme hmm", hi, "M", M, "reported", len(hits), "/", len(queries), "aligned", len(columns),
              "max score", max(h["score"] for h in hits.values()))
    with open(os.path.join(d, "golden.json"), "w") as f:
        json.dump(gold, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
