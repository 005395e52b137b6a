"""BASELINE config c1 end to end against the REFERENCE PIPELINE's own outputs (tests/golden/make_golden_c1.py ran
`SearchAlgorithm.search` and `witch.py -p ... --save-weight 1` of the reference on 100 fragments of examples/data against
a 7-subset eHMM built by the reference's `subset_alignment_and_hmmbuild`): bit-score tables, weights.txt, the "top N
weights" log lines, checkpoint rows, aligned.fasta and aligned.masked.fasta.

  * CPU: the oracle restatement reproduces all of it (a sample of the queries, to keep the CPU suite short).
  * GPU: the CUDA path through the mirror interface (BatchedSearch -> writeWeights -> getBackbones ->
    alignSubQueriesNew -> mergeAlignmentsCollapsed) reproduces all of it, character for character; a score that prints
    differently is only tolerated within 2e-3 bits of a rounding boundary and is counted.
"""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from golden_util import GOLDEN, read_fasta
from oracle import oracle as O

C1 = os.path.join(GOLDEN, "c1")


def _load(tmpdir):
    G = json.loads(gzip.open(os.path.join(C1, "c1_golden.json.gz")).read())
    paths = []
    for i in range(G["n_subsets"]):
        p = os.path.join(str(tmpdir), "c1_hmm_%d.hmm" % i)
        with gzip.open(os.path.join(C1, "hmm_%d.hmm.gz" % i), "rb") as g, open(p, "wb") as f:
            f.write(g.read())
        paths.append(p)
    queries = read_fasta(os.path.join(C1, "queries.fasta"))
    bbp = os.path.join(str(tmpdir), "c1_backbone.fasta")
    with gzip.open(os.path.join(C1, "backbone.fasta.gz"), "rb") as g, open(bbp, "wb") as f:
        f.write(g.read())
    backbone = read_fasta(bbp)
    return G, paths, queries, backbone


def _fasta_sha(names, rows):
    return hashlib.sha256("".join(">%s\n%s\n" % (n, r) for n, r in zip(names, rows)).encode()).hexdigest()


def test_c1_inputs_match_reference_directory_reader(tmp_path):
    """formats.obtainRetainedColumns / HMMSubset against the reference's readHMMDirectory / getAlignmentSubsets output."""
    from witch_b200 import formats
    G, paths, queries, backbone = _load(tmp_path)
    root = tmp_path / "root"
    for i, p in enumerate(paths):   # rebuild the eHMM directory layout from the committed pieces
        d = root / ("A_0_%d" % i)
        d.mkdir(parents=True)
        os.rename(p, d / ("hmmbuild.model.A_0_%d" % i))
        taxa = set(G["subset_taxa"][i])
        with open(d / ("hmmbuild.input.A_0_%d.fasta" % i), "w") as f:
            for n, s in backbone:
                if n in taxa:
                    f.write(">%s\n%s\n" % (n, s))
    idx = formats.getAlignmentSubsets(str(root))
    assert sorted(idx) == list(range(G["n_subsets"]))
    assert {str(i): idx[i].num_taxa for i in idx} == G["nseq"]
    bbp = tmp_path / "bb.fasta"
    with open(bbp, "w") as f:
        for n, s in backbone:
            f.write(">%s\n%s\n" % (n, s))
    ret, ng, B = formats.obtainRetainedColumns(str(bbp), idx)
    assert B == G["backbone_length"]
    for i in idx:
        assert list(ret[i]) == G["retained"][str(i)] and list(ng[i]) == G["nongaps"][str(i)]


def test_c1_oracle_reproduces_reference_pipeline(tmp_path):
    G, paths, queries, backbone = _load(tmp_path)
    profs = [O.Profile(p) for p in paths]
    H, B = len(paths), G["backbone_length"]
    ret = {h: G["retained"][str(h)] for h in range(H)}
    ng = {h: G["nongaps"][str(h)] for h in range(H)}
    sample = queries[::2]   # 50 of the 100 queries
    rows = {}
    for taxon, seq in sample:
        scores = {}
        for h, p in enumerate(profs):
            r = O.score_pair(p, p.abc.digitize(seq))
            assert r["reported"] == (taxon in G["hmmsearch"][str(h)]), (taxon, h)
            if r["reported"]:
                assert O.printed_score(r["score"]) == G["hmmsearch"][str(h)][taxon], (taxon, h, r["score"])
                scores[h] = O.printed_score(r["score"])
        ranked = O.rank_bitscores(scores)
        sw = O.calculate_weights([h for h, _ in ranked], [s for _, s in ranked], [profs[h].nseq for h, _ in ranked], 10)
        gw = G["weights"][taxon]
        assert len(sw) == len(gw)
        for (i, w), (gi, gwt) in zip(sw, gw):
            assert abs(w - gwt) <= 1e-12 * gwt
        assert sorted(i for i, _ in sw) == sorted(i for i, _ in gw)
        inc = O.adaptive_inclusion([tuple(x) for x in gw])
        assert len(inc) == G["log_top"][taxon][0] and [i for i, _ in inc] == [i for i, _ in G["log_top"][taxon][1]]
        s2c = {h: O.align_pair(profs[h], profs[h].abc.digitize(seq)) for h, _ in inc}
        row = O.compress_insertions(O.graph_align(seq.upper(), B, dict((i, w) for i, w in gw), s2c, ret, ng))
        assert row == G["checkpoint_rows"][taxon], taxon
        rows[taxon] = row
    # merged alignment of the sampled queries == the reference's rows restricted to them (insertion columns are per query
    # set, so compare through the masked rows and the ungapped content)
    merged, masked, _ = O.merge_rows([r for _, r in backbone] + list(rows.values()), B)
    for (taxon, _), mk in zip(sample, masked[len(backbone):]):
        assert mk == G["masked"][taxon]


@pytest.mark.gpu
def test_c1_cuda_path_reproduces_reference_pipeline(tmp_path):
    from witch_b200.gcmm import BatchedSearch, mergeAlignmentsCollapsed
    G, paths, queries, backbone = _load(tmp_path)
    H, B = len(paths), G["backbone_length"]
    names = [n for n, _ in queries]
    bs = BatchedSearch(paths, num_hmms=10)
    bs.search(names, [s for _, s in queries])
    # (1) bit-score tables == the files SearchAlgorithm.search left behind, as the reference's readHMMSearch parses them
    n_scores = n_boundary = 0
    bad_queries = set()
    for h in range(H):
        res = bs.hmmsearch_results(h)
        assert set(res) == set(G["hmmsearch"][str(h)]), h
        for t, (_, sc) in res.items():
            n_scores += 1
            if sc != G["hmmsearch"][str(h)][t]:
                raw = float(bs.scores[names.index(t), h]) * 10.0
                assert abs(raw - np.floor(raw) - 0.5) < 0.02, (t, h, raw / 10.0, G["hmmsearch"][str(h)][t])   # within 2e-3 bits of a print boundary
                n_boundary += 1
                bad_queries.add(t)
    assert n_scores == sum(len(v) for v in G["hmmsearch"].values()) and n_boundary <= 2
    # (2) weights.txt, (3) the log line's inclusion count
    t2w = bs.writeWeights()
    assert set(t2w) == set(G["weights"])
    for t, sw in t2w.items():
        if t in bad_queries:
            continue
        gw = G["weights"][t]
        assert len(sw) == len(gw)
        for (i, w), (gi, gwt) in zip(sw, gw):
            assert abs(w - gwt) <= 1e-12 * gwt, (t, sw, gw)
        assert sorted(i for i, _ in sw) == sorted(i for i, _ in gw)
    bb = bs.getBackbones(t2w)
    for t in t2w:
        if t not in bad_queries:
            log = bb[t][0]
            assert log.startswith("%s\tpassed to main pipeline with top %d weights: " % (t, G["log_top"][t][0])), log
    # (4) checkpoint rows, (5) aligned.fasta / aligned.masked.fasta
    ret = {h: G["retained"][str(h)] for h in range(H)}
    ng = {h: G["nongaps"][str(h)] for h in range(H)}
    rows = bs.alignSubQueriesNew(B, ret, ng, t2w)
    n_rows_same = sum(1 for t in rows if rows[t] == G["checkpoint_rows"][t])
    assert set(rows) == set(G["checkpoint_rows"])
    assert n_rows_same >= len(rows) - len(bad_queries) - 1, (n_rows_same, len(rows))   # (one documented FP32 near-tie class)
    order = [n for n in G["aligned_order"] if n in rows]
    full, mask = mergeAlignmentsCollapsed(bs.ehmm, backbone, {t: rows[t] for t in order}, B)
    assert list(full.keys()) == G["aligned_order"]
    if n_rows_same == len(rows):
        assert len(next(iter(full.values()))) == G["aligned_width"]
        assert _fasta_sha(full.keys(), full.values()) == G["aligned_sha256"]
        assert _fasta_sha(mask.keys(), mask.values()) == G["masked_sha256"]
    for t in rows:
        if rows[t] == G["checkpoint_rows"][t]:
            assert mask[t] == G["masked"][t]
            if n_rows_same == len(rows):
                assert full[t] == G["aligned"][t]
