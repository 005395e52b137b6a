"""Seeded synthetic workloads of the shapes BASELINE.json names (SURVEY.md 8(d) "Synthetic inputs").

Bench/test tooling, not product code. Generates:
  * a family of sequences evolved on a random binary tree with substitutions and indels, homology tracked so
    that the true alignment of any subset is known (no MAFFT/MAGUS needed),
  * a backbone (full-length leaves) + its true alignment, queries (full length and/or fragments),
  * the eHMM: hierarchical centroid bisection of the backbone tree down to <= `decomp` leaves
    (reference behaviour: witch_msa/gcmm/tree.py:384-438 keeps every subtree, 1+2+4+... subsets),
  * one HMMER3/f text profile per subset, built by the reference's own `hmmbuild` with WITCH's command line
    (`--cpu 1 --dna|--amino --ere 0.59 --symfrac 0.0 --informat afa`, gcmm/algorithm.py:463-470) when the staged
    binary oracle/_ref/hmmer/hmmbuild is present (meta["profiles"] == "hmmbuild"). Fallback, stated in
    meta["profiles"] == "pseudocount": a simple pseudocount estimator (every non-all-gap column is a match state,
    as `--symfrac 0.0` does). Both the CUDA path and the reference CPU binaries read these same files.
"""
import hashlib
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

DNA = "ACGT"
AMINO = "ACDEFGHIKLMNPQRSTVWY"


class Node:
    __slots__ = ("left", "right", "leaves", "blen")

    def __init__(self):
        self.left = self.right = None
        self.leaves = None
        self.blen = 0.0


def _evolve(rng, n_leaves, root_len, K, sub_rate, indel_rate, mean_blen):
    """Returns (leaf residue arrays, leaf coordinate arrays, tree root). Coordinates are floats; equal coordinate ==
    homologous site."""
    res = rng.integers(0, K, root_len).astype(np.int8)
    coord = np.arange(root_len, dtype=np.float64)
    # random (Yule-like) topology built top-down by random splits of leaf counts
    leaves_res, leaves_coord = [], []

    def grow(n, res, coord):
        node = Node()
        if n == 1:
            node.leaves = [len(leaves_res)]
            leaves_res.append(res)
            leaves_coord.append(coord)
            return node
        nl = int(rng.integers(1, n)) if n > 2 else 1
        if n > 8:  # keep the tree reasonably balanced so that depth stays O(log n)
            nl = int(np.clip(rng.normal(n / 2, n / 6), 1, n - 1))
        kids = []
        for m in (nl, n - nl):
            b = rng.exponential(mean_blen)
            r, c = _mutate(rng, res, coord, K, sub_rate * b, indel_rate * b)
            kids.append(grow(m, r, c))
        node.left, node.right = kids
        node.leaves = kids[0].leaves + kids[1].leaves
        return node

    import sys
    sys.setrecursionlimit(100000)
    root = grow(n_leaves, res, coord)
    return leaves_res, leaves_coord, root


def _mutate(rng, res, coord, K, psub, pindel):
    res = res.copy()
    n = len(res)
    m = rng.random(n) < min(psub, 0.75)
    res[m] = rng.integers(0, K, int(m.sum())).astype(np.int8)
    nev = rng.poisson(pindel * n)
    if nev == 0:
        return res, coord
    keep = np.ones(n, dtype=bool)
    ins_at, ins_res, ins_coord = [], [], []
    for _ in range(nev):
        p = int(rng.integers(0, n))
        ln = int(rng.geometric(1.0 / 3.0))
        if rng.random() < 0.5:
            keep[p:p + ln] = False
        else:
            lo = coord[p]
            hi = coord[p + 1] if p + 1 < n else coord[p] + 1.0
            cs = np.sort(lo + (hi - lo) * rng.random(ln))
            ins_at.append(p + 1)
            ins_res.append(rng.integers(0, K, ln).astype(np.int8))
            ins_coord.append(cs)
    if ins_at:
        order = np.argsort(ins_at, kind="stable")
        pieces_r, pieces_c, last = [], [], 0
        for o in order:
            a = ins_at[o]
            pieces_r += [res[last:a][keep[last:a]], ins_res[o]]
            pieces_c += [coord[last:a][keep[last:a]], ins_coord[o]]
            last = a
        pieces_r.append(res[last:][keep[last:]])
        pieces_c.append(coord[last:][keep[last:]])
        res, coord = np.concatenate(pieces_r), np.concatenate(pieces_c)
        o = np.argsort(coord, kind="stable")
        return res[o], coord[o]
    return res[keep], coord[keep]


def _decompose(node, max_size, out):
    """Hierarchical centroid-style bisection: keep every subtree, stop splitting at <= max_size leaves."""
    out.append(node.leaves)
    if len(node.leaves) <= max_size or node.left is None:
        return
    _decompose(node.left, max_size, out)
    _decompose(node.right, max_size, out)


def _subtree(root, leafset):
    """Smallest subtree whose leaves are all in leafset order (restrict a tree to a set of leaves)."""
    def rec(n):
        if n.left is None:
            if n.leaves[0] in leafset:
                m = Node(); m.leaves = list(n.leaves); return m
            return None
        a, b = rec(n.left), rec(n.right)
        if a is None:
            return b
        if b is None:
            return a
        m = Node(); m.left, m.right = a, b; m.leaves = a.leaves + b.leaves
        return m
    return rec(root)


def write_hmm(path, name, counts, trans_counts, nseq, alphabet):
    """HMMER3/f ASCII. counts[M][K] residue counts per match column; trans_counts[M+1][4] = (MM, MD, DM, DD) counts
    leaving node k (k = 0 is the begin node)."""
    K = len(alphabet)
    M = counts.shape[0]
    bg = np.full(K, 1.0 / K)
    # entropy-flattening stand-in for --ere: total column weight capped
    tot = counts.sum(1, keepdims=True)
    w = np.minimum(1.0, 3.0 / np.maximum(tot, 1.0))
    mat = (counts * w + 1.0 * bg) / (tot * w + 1.0)
    tc = trans_counts.astype(np.float64)
    wt = np.minimum(1.0, 4.0 / np.maximum(tc.sum(1, keepdims=True), 1.0))
    tc = tc * wt
    a = 2.0
    mm, md = tc[:, 0] + a * 0.95, tc[:, 1] + a * 0.03
    mi = np.full(M + 1, a * 0.02)
    sm = mm + md + mi
    dm, dd = tc[:, 2] + a * 0.6, tc[:, 3] + a * 0.4
    sd = dm + dd
    t = np.stack([mm / sm, mi / sm, md / sm, np.full(M + 1, 0.75), np.full(M + 1, 0.25), dm / sd, dd / sd], 1)

    def fmt(v):
        return "  ".join("%7.5f" % x for x in v)

    lines = ["HMMER3/f [3.1b2 | February 2015]", "NAME  %s" % name, "LENG  %d" % M,
             "ALPH  %s" % ("amino" if K == 20 else "DNA"), "RF    no", "MM    no", "CONS  yes", "CS    no", "MAP   no",
             "NSEQ  %d" % nseq, "EFFN  %f" % min(float(nseq), 3.0),
             "STATS LOCAL MSV      -11.0000  0.70000", "STATS LOCAL VITERBI  -12.0000  0.70000",
             "STATS LOCAL FORWARD   -5.0000  0.70000",
             "HMM     " + "".join("     %s   " % c for c in alphabet),
             "            m->m     m->i     m->d     i->m     i->i     d->m     d->d"]
    ins = "          " + fmt(-np.log(bg))
    nl = -np.log(t)
    lm = -np.log(mat)
    lines.append(ins)
    lines.append("          " + fmt(nl[0][:5]) + "  0.00000        *")
    for k in range(1, M + 1):
        lines.append("%7d   %s %6s %s - - -" % (k, fmt(lm[k - 1]), "-", alphabet[int(np.argmax(mat[k - 1]))].lower()))
        lines.append(ins)
        if k < M:
            lines.append("          " + fmt(nl[k]))
        else:
            mmM = t[k][0] / (t[k][0] + t[k][1])
            lines.append("          %7.5f  %7.5f        *  %7.5f  %7.5f  0.00000        *" % (
                -np.log(mmM), -np.log(1 - mmM), nl[k][3], nl[k][4]))
    lines.append("//")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


HMMBUILD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "hmmer", "hmmbuild")


def _hmmbuild_leng(path):
    try:
        with open(path) as f:
            for ln in f:
                if ln.startswith("LENG"):
                    return int(ln.split()[1])
                if ln.startswith("HMM "):
                    break
    except OSError:
        pass
    return -1


def _run_hmmbuild(job):
    """One subset: write hmmbuild.input.<label>.fasta (the subset alignment, all-gap columns removed) and run the
    reference's hmmbuild with WITCH's flags (gcmm/algorithm.py:463-470). -> True when the profile has one match state
    per retained column."""
    hmm_path, aln_path, rows, lut, molecule, ncols = job
    if _hmmbuild_leng(hmm_path) == ncols:
        return True
    with open(aln_path, "w") as f:
        for r in range(rows.shape[0]):
            f.write(">s%d\n%s\n" % (r, "".join(lut[rows[r]])))
    subprocess.run([HMMBUILD, "--cpu", "1", "--" + molecule, "--ere", "0.59", "--symfrac", "0.0", "--informat", "afa",
                    "-o", "/dev/null", hmm_path, aln_path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    os.remove(aln_path)
    return _hmmbuild_leng(hmm_path) == ncols


def make_workload(outdir, alphabet="dna", c1=False, n_total=10000, n_backbone=1000, root_len=1550, decomp=10, frag_frac=0.25,
                  frag_mean=400, seed=1, n_queries=None, max_hmms=None, mean_blen=0.02, indel_rate=0.0065,
                  profiles="auto"):
    """Build (or reuse) a workload directory. Returns dict(hmm_paths, nseq, names, seqs, retained_columns,
    nongaps_per_column, backbone_length, meta). profiles: "hmmbuild" (the reference binary), "pseudocount" (the
    built-in estimator) or "auto" (hmmbuild when oracle/_ref is staged)."""
    if profiles == "auto":
        profiles = "hmmbuild" if os.access(HMMBUILD, os.X_OK) else "pseudocount"
    if c1:
        return make_workload_c1(outdir, max_hmms=max_hmms)
    key = hashlib.sha1(repr((alphabet, n_total, n_backbone, root_len, decomp, frag_frac, frag_mean, seed, n_queries,
                             max_hmms, mean_blen, indel_rate, "v6", profiles)).encode()).hexdigest()[:12]
    wd = os.path.join(outdir, "synth_" + key)
    abc = DNA if alphabet == "dna" else AMINO
    K = len(abc)
    rng = np.random.default_rng(seed)
    leaves_res, leaves_coord, root = _evolve(rng, n_total, root_len, K, sub_rate=1.0, indel_rate=indel_rate,
                                             mean_blen=mean_blen)
    # backbone = a random subset of leaves (full length); queries = the rest
    perm = rng.permutation(n_total)
    bb = np.sort(perm[:n_backbone])
    qs = np.sort(perm[n_backbone:])
    if n_queries is not None:
        qs = qs[rng.integers(0, len(qs), n_queries)] if n_queries > len(qs) else qs[:n_queries]
    allc = np.unique(np.concatenate([leaves_coord[i] for i in bb]))
    backbone_length = len(allc)
    bb_rows = np.full((len(bb), backbone_length), -1, dtype=np.int8)
    for r, i in enumerate(bb):
        bb_rows[r, np.searchsorted(allc, leaves_coord[i])] = leaves_res[i]
    bbtree = _subtree(root, set(int(x) for x in bb))
    subsets = []
    _decompose(bbtree, decomp, subsets)
    if max_hmms:
        subsets = subsets[:max_hmms]
    row_of = {int(i): r for r, i in enumerate(bb)}
    os.makedirs(wd, exist_ok=True)
    hmm_paths, nseqs, retained, nongaps, jobs = [], [], [], [], []
    for si, leaves in enumerate(subsets):
        rows = bb_rows[[row_of[i] for i in leaves]]
        present = rows >= 0
        cols = np.nonzero(present.any(0))[0]
        sub = rows[:, cols]
        pres = sub >= 0
        counts = np.zeros((len(cols), K))
        for x in range(K):
            counts[:, x] = (sub == x).sum(0)
        # transitions between consecutive retained columns (M = residue, D = gap); node 0 = begin
        a, b = pres[:, :-1], pres[:, 1:]
        tcn = np.zeros((len(cols) + 1, 4))
        tcn[1:-1, 0] = (a & b).sum(0); tcn[1:-1, 1] = (a & ~b).sum(0)
        tcn[1:-1, 2] = (~a & b).sum(0); tcn[1:-1, 3] = (~a & ~b).sum(0)
        tcn[0, 0] = pres[:, 0].sum(); tcn[0, 1] = (~pres[:, 0]).sum()
        tcn[-1, 0] = pres[:, -1].sum(); tcn[-1, 2] = (~pres[:, -1]).sum()
        p = os.path.join(wd, "hmmbuild.model.A_0_%d" % si)
        if profiles == "hmmbuild":
            jobs.append((p, os.path.join(wd, "hmmbuild.input.A_0_%d.fasta" % si), sub, np.array(list(abc + "-")),
                         "dna" if alphabet == "dna" else "amino", len(cols)))
        elif not os.path.exists(p):
            write_hmm(p, "A_0_%d" % si, counts, tcn, len(leaves), abc)
        hmm_paths.append(p); nseqs.append(len(leaves))
        retained.append(cols.astype(np.int32)); nongaps.append(pres.sum(0).astype(np.int32))
    if jobs:
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
            ok = list(ex.map(_run_hmmbuild, jobs))
        if not all(ok):
            raise RuntimeError("hmmbuild did not produce one match state per retained column for %d subsets" % ok.count(False))
    names, seqs = [], []
    lut = np.array(list(abc))
    for n, i in enumerate(qs):
        r = leaves_res[i]
        # fresh 2% point mutations so that resampled leaves are not identical
        r = r.copy()
        m = rng.random(len(r)) < 0.02
        r[m] = rng.integers(0, K, int(m.sum()))
        if rng.random() < frag_frac:
            ln = int(np.clip(rng.normal(frag_mean, 0.15 * frag_mean), 0.2 * frag_mean, min(2 * frag_mean, len(r))))
            s = int(rng.integers(0, max(1, len(r) - ln + 1)))
            r = r[s:s + ln]
        names.append("Q%06d" % n)
        seqs.append("".join(lut[r]))
    meta = dict(alphabet=alphabet, n_queries=len(seqs), H=len(hmm_paths), sumL=int(sum(len(s) for s in seqs)),
                sumM=int(sum(len(c) for c in retained)), backbone_length=int(backbone_length), seed=seed, dir=wd,
                profiles=profiles)
    return dict(hmm_paths=hmm_paths, nseq=nseqs, names=names, seqs=seqs, retained_columns=retained,
                nongaps_per_column=nongaps, backbone_length=backbone_length, meta=meta)


def make_workload_c1(outdir, max_hmms=None, decomp=10, **_):
    """BASELINE config c1: the reference's bundled nucleotide example (examples/data: 500-sequence backbone alignment,
    500 fragments), from the copies committed under tests/golden/c1 by tests/golden/make_golden_c1.py. The eHMM is the
    hierarchical decomposition WITCH builds with `-A 10` (every subtree above 10 leaves plus the leaves' subtrees:
    1 + 2 + ... + 64 = 127 subsets); the tree itself is not on the hot path, so the backbone rows are halved in file order
    (subset SIZES match the reference's). Profiles: the reference's hmmbuild with WITCH's flags (oracle/_ref)."""
    import gzip
    gdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1")

    def rd(path):
        out = []
        with gzip.open(path, "rt") as f:
            for ln in f:
                ln = ln.strip()
                if ln.startswith(">"):
                    out.append([ln[1:].split()[0], ""])
                elif ln:
                    out[-1][1] += ln
        return out
    bb = rd(os.path.join(gdir, "backbone.fasta.gz"))
    qs = rd(os.path.join(gdir, "queries_all.fasta.gz"))
    if not os.access(HMMBUILD, os.X_OK):
        raise RuntimeError("config c1 needs the staged reference hmmbuild (oracle/_ref/hmmer/hmmbuild)")
    lut = {c: i for i, c in enumerate(DNA)}
    rows = np.full((len(bb), len(bb[0][1])), -1, dtype=np.int8)
    for r, (_, s) in enumerate(bb):
        a = np.frombuffer(s.upper().encode(), dtype=np.uint8)
        for c, i in lut.items():
            rows[r, a == ord(c)] = i
        rows[r, (a != ord("-")) & (rows[r] < 0)] = 0   # (degenerate backbone letters, if any, count as residues)
    subsets = []

    def split(lo, hi):
        subsets.append((lo, hi))
        if hi - lo > decomp:
            mid = (lo + hi) // 2
            split(lo, mid); split(mid, hi)
    split(0, len(bb))
    subsets.sort(key=lambda x: (-(x[1] - x[0]), x[0]))   # breadth-first order like the reference's labels
    if max_hmms:
        subsets = subsets[:max_hmms]
    wd = os.path.join(outdir, "c1_%d_%s" % (decomp, max_hmms))
    os.makedirs(wd, exist_ok=True)
    hmm_paths, nseqs, retained, nongaps, jobs = [], [], [], [], []
    for si, (lo, hi) in enumerate(subsets):
        sub = rows[lo:hi]
        cols = np.nonzero((sub >= 0).any(0))[0]
        sub = sub[:, cols]
        p = os.path.join(wd, "hmmbuild.model.A_0_%d" % si)
        jobs.append((p, os.path.join(wd, "hmmbuild.input.A_0_%d.fasta" % si), sub, np.array(list(DNA + "-")), "dna", len(cols)))
        hmm_paths.append(p); nseqs.append(hi - lo)
        retained.append(cols.astype(np.int32)); nongaps.append((sub >= 0).sum(0).astype(np.int32))
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        ok = list(ex.map(_run_hmmbuild, jobs))
    if not all(ok):
        raise RuntimeError("hmmbuild failed for %d c1 subsets" % ok.count(False))
    names, seqs = [n for n, _ in qs], [s.upper() for _, s in qs]
    meta = dict(alphabet="dna", n_queries=len(seqs), H=len(hmm_paths), sumL=int(sum(len(s) for s in seqs)),
                sumM=int(sum(len(c) for c in retained)), backbone_length=int(rows.shape[1]), seed=None, dir=wd, profiles="hmmbuild")
    return dict(hmm_paths=hmm_paths, nseq=nseqs, names=names, seqs=seqs, retained_columns=retained, nongaps_per_column=nongaps,
                backbone_length=int(rows.shape[1]), meta=meta)


CONFIGS = {
    # name: kwargs; "c1" is not synthetic: make_workload dispatches it to make_workload_c1
    "c1": dict(c1=True),
    "c2": dict(alphabet="dna", n_total=10000, n_backbone=1000, root_len=1550, decomp=10, frag_frac=0.25, frag_mean=400, seed=1),
    "c3": dict(alphabet="dna", n_total=27643, n_backbone=1000, root_len=1500, decomp=10, frag_frac=1.0, frag_mean=560, seed=2),
    "c4": dict(alphabet="amino", n_total=20000, n_backbone=800, root_len=300, decomp=10, frag_frac=0.0, frag_mean=150, seed=3, mean_blen=0.05, indel_rate=0.004),
    "c5": dict(alphabet="dna", n_total=10000, n_backbone=1000, root_len=1550, decomp=10, frag_frac=1.0, frag_mean=400, seed=4, n_queries=100000),
    "tiny": dict(alphabet="dna", n_total=300, n_backbone=60, root_len=300, decomp=10, frag_frac=0.5, frag_mean=120, seed=5),
    "tiny_aa": dict(alphabet="amino", n_total=200, n_backbone=40, root_len=120, decomp=10, frag_frac=0.3, frag_mean=60, seed=6),
}
