"""Generate the golden vectors that PIN the oracle (and, through it, the CUDA path) to the reference.

Run in the build container only (needs /root/reference for the example data and oracle/_ref/hmmer/* staged by
oracle/make_ref.py).  Everything written here is small and committed under tests/golden/:

    <set>/hmm_<i>.hmm.gz      profile built by the reference's hmmbuild with WITCH's flags
                              (witch_msa/gcmm/algorithm.py:463-470: --cpu 1 --<mol> --ere 0.59 --symfrac 0.0 --informat afa)
    <set>/queries.fasta       the query sequences
    <set>/golden.json         per HMM: NSEQ, M; per (HMM, query): hmmsearch's printed score / bias / domain
                              envelopes (command of algorithm.py:526-532 plus --tblout/--domtblout, which only add
                              machine-readable copies of the same numbers), and hmmalign's per-residue column list
                              (command of aligner.py:98-100, Stockholm -> columns as aligner.py:126-142)

Usage: python tests/golden/make_golden.py
"""
import gzip
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.make_ref import ref_tool, build as build_ref  # noqa: E402

DATA = "/root/reference/examples/data"


def read_fasta(path):
    op = gzip.open if path.endswith(".gz") else open
    name, seqs, order = None, {}, []
    with op(path, "rt") as f:
        for ln in f:
            ln = ln.strip()
            if not ln:
                continue
            if ln[0] == ">":
                name = ln[1:].split()[0]
                seqs[name] = []
                order.append(name)
            else:
                seqs[name].append(ln)
    return [(n, "".join(seqs[n])) for n in order]


def write_fasta(path, items):
    with open(path, "w") as f:
        for n, s in items:
            f.write(">%s\n%s\n" % (n, s))


def hmmbuild(aln_items, out_hmm, mol):
    """Same column handling as subset_alignment_and_hmmbuild: drop all-gap columns, then hmmbuild."""
    cols = [j for j in range(len(aln_items[0][1])) if any(s[j] != "-" for _, s in aln_items)]
    sub = [(n, "".join(s[j] for j in cols)) for n, s in aln_items]
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "in.fasta")
        write_fasta(fa, sub)
        cmd = [ref_tool("hmmbuild"), "--cpu", "1", "--" + mol, "--ere", "0.59", "--symfrac", "0.0",
               "--informat", "afa", "-o", "/dev/null", out_hmm, fa]
        subprocess.check_call(cmd)
    nongaps = [sum(1 for _, s in sub if s[j] != "-") for j in range(len(cols))]
    return cols, nongaps


def hmmsearch(hmm, fasta):
    """WITCH's command + tblout/domtblout. Returns {name: dict(score,bias,ndom,domains=[(score,bias,envfrom,envto)])}."""
    with tempfile.TemporaryDirectory() as td:
        out, tbl, dom = (os.path.join(td, x) for x in ("out", "tbl", "dom"))
        cmd = [ref_tool("hmmsearch"), "--cpu", "1", "--noali", "-E", "99999999", "-o", out, "--max",
               "--tblout", tbl, "--domtblout", dom, hmm, fasta]
        subprocess.check_call(cmd)
        res = {}
        for ln in open(tbl):
            if ln.startswith("#"):
                continue
            t = ln.split()
            res[t[0]] = dict(score=float(t[5]), bias=float(t[6]), domains=[])
        for ln in open(dom):
            if ln.startswith("#"):
                continue
            t = ln.split()
            res[t[0]]["domains"].append((float(t[13]), float(t[14]), int(t[19]), int(t[20])))
    return res


def hmmalign_columns(hmm, name, seq):
    """One query per process, exactly like getBackbones; Stockholm row -> column list (aligner.py:126-142)."""
    with tempfile.TemporaryDirectory() as td:
        fa, out = os.path.join(td, "c1.fasta"), os.path.join(td, "out.sto")
        write_fasta(fa, [(name, seq)])
        subprocess.check_call([ref_tool("hmmalign"), "-o", out, hmm, fa])
        row = []
        for ln in open(out):
            if ln.startswith("#") or ln.startswith("//") or not ln.strip():
                continue
            t = ln.split()
            if t[0] == name:
                row.append(t[1])
        row = "".join(row).replace(".", "-")
    cols, regular = [], 0
    for ch in row:
        if ch == "-":
            regular += 1
        elif ch.islower():
            cols.append(-1)
        else:
            cols.append(regular)
            regular += 1
    assert len(cols) == len(seq), (len(cols), len(seq))
    return cols


def make_set(setname, mol, hmm_alns, queries, align_pairs="reported", max_align=None, seed=0):
    d = os.path.join(HERE, setname)
    os.makedirs(d, exist_ok=True)
    write_fasta(os.path.join(d, "queries.fasta"), queries)
    gold = dict(molecule=mol, hmms=[])
    rng = random.Random(seed)
    for hi, aln in enumerate(hmm_alns):
        hmm = os.path.join(d, "hmm_%d.hmm" % hi)
        cols, nongaps = hmmbuild(aln, hmm, mol)
        hits = hmmsearch(hmm, os.path.join(d, "queries.fasta"))
        names = [n for n, _ in queries]
        cand = [n for n in names if n in hits] if align_pairs == "reported" else list(names)
        if max_align is not None and len(cand) > max_align:
            cand = sorted(rng.sample(cand, max_align), key=names.index)
        qd = dict(queries)
        columns = {n: hmmalign_columns(hmm, n, qd[n]) for n in cand}
        M = NSEQ = None
        for ln in open(hmm):
            t = ln.split()
            if t and t[0] == "LENG":
                M = int(t[1])
            if t and t[0] == "NSEQ":
                NSEQ = int(t[1])
            if t and t[0] == "HMM":
                break
        with open(hmm, "rb") as f, gzip.GzipFile(hmm + ".gz", "wb", mtime=0) as g:
            g.write(f.read())
        os.remove(hmm)
        gold["hmms"].append(dict(file="hmm_%d.hmm.gz" % hi, M=M, nseq=NSEQ, retained_columns=cols,
                                 nongaps_per_column=nongaps, hits=hits, columns=columns,
                                 taxa=[n for n, _ in aln]))
        print(setname, "hmm", hi, "M", M, "nseq", NSEQ, "reported", len(hits), "/", len(queries),
              "aligned", len(columns))
    with open(os.path.join(d, "golden.json"), "w") as f:
        json.dump(gold, f, separators=(",", ":"))


def evolve_protein_family(seed, nleaf, length):
    """Tiny seeded protein family (true alignment known): root -> star/binary mix with substitutions + indels."""
    rng = random.Random(seed)
    aa = "ACDEFGHIKLMNPQRSTVWY"
    root = [rng.choice(aa) for _ in range(length)]
    rows = []
    for i in range(nleaf):
        base = root if i < 2 or rng.random() < 0.5 else list(rows[rng.randrange(len(rows))])
        row = list(base)
        for j in range(length):
            if row[j] != "-" and rng.random() < 0.25:
                row[j] = rng.choice(aa)
        if rng.random() < 0.6:  # a deletion
            s = rng.randrange(length - 12)
            for j in range(s, s + rng.randrange(2, 10)):
                row[j] = "-"
        rows.append(row)
    return [("P%02d" % i, "".join(r)) for i, r in enumerate(rows)]


def main():
    assert build_ref(), "reference binaries not staged"
    bb = read_fasta(os.path.join(DATA, "backbone.aln.fasta.gz"))
    frags = read_fasta(os.path.join(DATA, "unaligned_frag.fasta"))
    full = read_fasta(os.path.join(DATA, "unaligned_all.fasta"))
    rng = random.Random(7)

    # set 1: small windowed DNA profiles (M ~ 150-330) x 60 fragments: local hits with N/C flanks, unreported pairs
    win = lambda items, a, b: [(n, s[a:b]) for n, s in items if any(c != "-" for c in s[a:b])]
    q1 = [frags[i] for i in sorted(rng.sample(range(len(frags)), 60))]
    make_set("dna_small", "dna", [win(bb[:10], 300, 800), win(bb[40:70], 900, 1500), win(bb[100:104], 0, 600)], q1)

    # set 2: a real sub-HMM of WITCH's decomposition size (8 leaves, M ~ 1052) and the whole 500-seq backbone is too
    # big to commit; use 8-seq full-width profile x 40 fragments (+ a query with degenerate residues)
    q2 = [frags[i] for i in sorted(rng.sample(range(len(frags)), 40))]
    s = list(q2[0][1])
    for pos, ch in ((5, "N"), (17, "N"), (40, "R"), (41, "Y"), (77, "N"), (90, "N"), (91, "N"), (120, "N"), (150, "N")):
        if pos < len(s):
            s[pos] = ch
    q2.append(("DEGEN1", "".join(s)))
    make_set("dna_sub8", "dna", [bb[:8]], q2, max_align=24, seed=1)

    # set 3: full-length queries (~1 kb) against a 60-sequence profile window: multi-domain-flagged cases
    q3 = [full[i] for i in sorted(rng.sample(range(len(full)), 24))]
    make_set("dna_full", "dna", [win(bb[200:260], 0, 1200)], q3, max_align=10, seed=2)

    # set 4: protein family (null2 biases of several bits)
    fam = evolve_protein_family(11, 14, 120)
    rngp = random.Random(5)
    aa = "ACDEFGHIKLMNPQRSTVWY"
    q4 = []
    for i in range(16):
        src = fam[rngp.randrange(len(fam))][1].replace("-", "")
        a = rngp.randrange(0, 40)
        b = rngp.randrange(70, len(src))
        sq = list(src[a:b])
        for j in range(len(sq)):
            if rngp.random() < 0.15:
                sq[j] = rngp.choice(aa)
        if i % 5 == 0:  # low-complexity tail -> null2 bias
            sq += list(rngp.choice(aa) * rngp.randrange(10, 25))
        if i == 3:
            sq[4] = "X"; sq[9] = "B"
        q4.append(("Q%02d" % i, "".join(sq)))
    make_set("amino_small", "amino", [fam[:12], fam[3:9]], q4)


if __name__ == "__main__":
    main()
