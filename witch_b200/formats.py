"""On-disk formats at the edges of the hot path: the eHMM directory WITCH builds
(`tree_decomp/root/A_0_<i>/{hmmbuild.input.<label>.fasta, hmmbuild.model.<label>, hmmsearch.results.*}`), the
per-HMM bit-score files a later `-p` run re-reads, `weights.txt`, and the gzip checkpoint of finished query rows.
Readers accept what the reference writes; writers produce files the reference's own readers accept
(ast.literal_eval here where the reference uses eval). Host-side Python only: nothing here touches the GPU."""
import ast
import gzip
import os
import re

import numpy as np


def read_fasta(path):
    op = gzip.open if path.endswith(".gz") else open
    name, chunks = None, []
    with op(path, "rt") as f:
        for ln in f:
            ln = ln.rstrip("\n")
            if ln.startswith(">"):
                if name is not None:
                    yield name, "".join(chunks)
                name, chunks = ln[1:].strip(), []
            elif ln:
                chunks.append(ln.strip())
    if name is not None:
        yield name, "".join(chunks)


class HMMSubset:
    """gcmm/loader.py:17-58: one `A_0_<index>` directory (alignment path, model path, NSEQ of the model header)."""

    def __init__(self, path, index):
        self.alignment_dir, self.index, self.num_taxa = path, index, 0
        files = sorted(os.listdir(path))
        aln = [f for f in files if f.startswith("hmmbuild.input")]
        mod = [f for f in files if f.startswith("hmmbuild.model.")]
        if not mod:
            raise FileNotFoundError("no hmmbuild.model.* in %s" % path)
        self.alignment_path = os.path.realpath(os.path.join(path, aln[0])) if aln else None
        self.hmm_model_path = os.path.realpath(os.path.join(path, mod[0]))
        with open(self.hmm_model_path) as f:
            for _ in range(20):                      # the reference looks at the first 20 header lines
                t = f.readline().split()
                if t and t[0] == "NSEQ":
                    self.num_taxa = int("".join(t[1:]))
                    break
        if self.num_taxa == 0:
            raise ValueError("Cannot find field: NSEQ, from {}".format(path))


def getAlignmentSubsets(path):
    """gcmm/loader.py:248-272: every `A_0_*` directory below `path` -> {index: HMMSubset}."""
    index_to_hmm = {}
    for root, dirs, _ in os.walk(path):
        for d in dirs:
            m = re.fullmatch(r"A_0_(\d+)", d)
            if m:
                index_to_hmm[int(m.group(1))] = HMMSubset(os.path.join(root, d), int(m.group(1)))
    return index_to_hmm


def obtainRetainedColumns(backbone_path, index_to_hmm):
    """gcmm/algorithm.py:551-574 (`-p` path) == :405-429: for every subset, the backbone columns that are not all-gap
    within the subset's rows and the number of non-gap characters in each of them.
    -> (subset_to_retained_columns, subset_to_nongaps_per_column, backbone_length)."""
    names, rows = [], []
    for n, s in read_fasta(backbone_path):
        names.append(n); rows.append(np.frombuffer(s.encode(), dtype=np.uint8))
    mat = np.stack(rows) != ord("-")
    row_of = {n: i for i, n in enumerate(names)}
    retained, nongaps = {}, {}
    for idx, sub in index_to_hmm.items():
        taxa = [n for n, _ in read_fasta(sub.alignment_path)]
        cnt = mat[[row_of[t] for t in taxa]].sum(0)
        cols = np.nonzero(cnt)[0]
        retained[idx] = tuple(int(c) for c in cols)
        nongaps[idx] = tuple(int(c) for c in cnt[cols])
    return retained, nongaps, mat.shape[1]


def writeHMMSearchResults(index_to_hmm, names, scores, reported, chunk=0):
    """The file subset_frag_chunk_hmmsearch leaves behind (gcmm/algorithm.py:482-544): `str({name: (evalue, score)})`
    at <dir>/hmmsearch.results.<label>.fragment_chunk_<i>. Scores are the printed 1-decimal values; the E-value slot
    is never read downstream (loader.py:293 takes scores[1]) and is written as 0.0. Column h of `scores` belongs to
    the h-th smallest index of index_to_hmm."""
    paths = []
    for h, idx in enumerate(sorted(index_to_hmm)):
        sub = index_to_hmm[idx]
        label = os.path.basename(sub.hmm_model_path)[len("hmmbuild.model."):]
        d = {names[q]: (0.0, float("%.1f" % scores[q, h])) for q in range(len(names)) if reported[q, h]}
        p = os.path.join(sub.alignment_dir, "hmmsearch.results.{}.fragment_chunk_{}".format(label, chunk))
        with open(p, "w") as f:
            f.write(str(d))
        paths.append(p)
    return paths


def readHMMSearch(subset):
    """gcmm/loader.py:277-294: {taxon: [(subset.index, score), ...]} from every hmmsearch.results.* of one subset."""
    ranks = {}
    for f in sorted(os.listdir(subset.alignment_dir)):
        if f.startswith("hmmsearch.results."):
            with open(os.path.join(subset.alignment_dir, f)) as fh:
                for taxon, sc in ast.literal_eval(fh.read()).items():
                    ranks.setdefault(taxon, []).append((subset.index, sc[1]))
    return ranks


def readAndRankBitscore(index_to_hmm, renamed_taxa=None):
    """gcmm/loader.py:299-332: merge the per-subset files and sort each query's list by score, descending (ties keep
    ascending subset index here; the reference's order among ties depends on which future finishes first)."""
    ranks = {}
    for idx in sorted(index_to_hmm):
        for taxon, lst in readHMMSearch(index_to_hmm[idx]).items():
            ranks.setdefault(taxon, []).extend(lst)
    renamed_taxa = renamed_taxa or {}
    return {renamed_taxa.get(t, t): sorted(v, key=lambda x: x[1], reverse=True) for t, v in ranks.items()}


def writeWeightsToLocal(taxon_to_weights, path):
    """gcmm/weighting.py:174-178; plain ints/floats so that readWeightsFromLocal's eval works under any numpy."""
    with open(path, "w") as f:
        for taxon, weights in taxon_to_weights.items():
            f.write("{}:{}\n".format(taxon, tuple((int(i), float(w)) for i, w in weights)))


def readWeightsFromLocal(path):
    """gcmm/weighting.py:184-194 (the taxon is everything before the LAST ':' here, so names may contain colons)."""
    out = {}
    with open(path) as f:
        for line in f:
            if line.strip():
                taxon, w = line.rstrip("\n").rsplit(":", 1)
                out[taxon] = ast.literal_eval(w)
    return out


def writeCheckpointAlignments(path, taxon_to_row, append=True):
    """gcmm/callback.py:19-29 / loader.py:79-92: one gzip member per call, one `taxon<TAB>row` line per finished query
    (a batch of queries per call instead of one query per call; the reference's reader concatenates members)."""
    with gzip.open(path, "ab" if append else "wb") as f:
        for taxon, row in taxon_to_row.items():
            f.write("{}\t{}\n".format(taxon, row).encode("utf-8"))


def readCheckpointAlignments(path):
    """gcmm/loader.py:95-150: {taxon: row}; the taxon is everything before the last TAB (names may contain tabs)."""
    out = {}
    with gzip.open(path, "rb") as f:
        for line in f.read().decode("utf-8").split("\n")[:-1]:
            parts = line.split("\t")
            out["\t".join(parts[:-1])] = parts[-1]
    return out
