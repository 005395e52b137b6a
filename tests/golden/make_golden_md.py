"""Golden vectors for the multi-domain branch of hmmsearch's domain definition (oracle/hmm_md.c).

Run in the build container only (needs oracle/_ref/hmmer/hmmsearch staged by oracle/make_ref.py). For every pair of
the committed golden sets whose region fails HMMER's single-domain test (oracle flag bit 0) the reference binary is
run on that single query with WITCH's command line (gcmm/algorithm.py:526-532) plus `--domE 99999999 --tblout
--domtblout` (machine-readable copies of the same numbers, every domain listed) under tools/hmmer_probe's qsort
interposer, which captures the cluster list {i, j, k, m, count of 200 traces} that p7_spensemble_Cluster builds for
each such region. Written: tests/golden/md_golden.json
    {set: {"<hmm index>": {query: {"score", "bias", "domains": [[score, bias, ienv, jenv]], "clusters": [[i,j,k,m,count]]}}}}

Usage: python tests/golden/make_golden_md.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools", "hmmer_probe"))
from golden_util import SETS, load_set  # noqa: E402
from oracle import oracle as O  # noqa: E402
import probe  # noqa: E402


def main():
    out = {}
    for setname in SETS:
        gold, queries, paths = load_set(setname)
        out[setname] = {}
        for h, path in enumerate(paths):
            prof = O.Profile(path)
            flagged = [(n, s) for n, s in queries if O.score_pair(prof, prof.abc.digitize(s), multidomain=False)["flags"] & 1]
            ref = probe.hmmsearch_probe(path, flagged)
            out[setname][str(h)] = {n: dict(score=r["score"], bias=r["bias"], domains=[list(d) for d in r["domains"]],
                                            clusters=[list(c) for reg in r["clusters"] for c in reg])
                                    for n, r in ref.items()}
            print(setname, h, "flagged pairs", len(flagged), "with clusters", sum(1 for r in ref.values() if r["clusters"]))
    with open(os.path.join(HERE, "md_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"), sort_keys=True)


if __name__ == "__main__":
    main()
