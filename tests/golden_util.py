"""Helpers shared by the parity tests: load tests/golden/<set> (written by tests/golden/make_golden.py)."""
import gzip
import json
import os
import tempfile

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# amino_extreme (tests/golden/make_golden_extreme.py): W/C-rich hand-written profiles, scores of thousands of bits
SETS = ("dna_small", "dna_sub8", "dna_full", "amino_small", "amino_extreme")


def read_fasta(path):
    out = []
    with open(path) as f:
        for ln in f:
            ln = ln.strip()
            if ln.startswith(">"):
                out.append([ln[1:].split()[0], ""])
            elif ln:
                out[-1][1] += ln
    return [(a, b) for a, b in out]


def load_set(name, workdir=None):
    """-> (gold dict, [(qname, seq)], [unpacked hmm paths])"""
    d = os.path.join(GOLDEN, name)
    gold = json.load(open(os.path.join(d, "golden.json")))
    queries = read_fasta(os.path.join(d, "queries.fasta"))
    workdir = workdir or tempfile.mkdtemp(prefix="witch_golden_")
    os.makedirs(workdir, exist_ok=True)
    paths = []
    for h in gold["hmms"]:
        p = os.path.join(workdir, name + "_" + h["file"][:-3])
        if not os.path.exists(p):
            with gzip.open(os.path.join(d, h["file"]), "rb") as g, open(p, "wb") as f:
                f.write(g.read())
        paths.append(p)
    return gold, queries, paths
