"""CPU: the oracle (oracle/hmm_oracle.c + oracle/oracle.py) against the reference binaries' golden vectors.

Pins: (1) reported-or-not, (2) the printed 1-decimal score, (3) hmmalign column lists, for every pair in
tests/golden/*.  The only tolerated differences are the documented multi-domain deviation (oracle flag bit 0:
HMMER resolves such regions by stochastic traceback clustering, the restatement keeps one envelope).
"""
import numpy as np
import pytest

from golden_util import SETS, load_set
from oracle import oracle as O


@pytest.mark.parametrize("setname", SETS)
def test_scores_and_columns_match_reference_binaries(setname, tmp_path):
    gold, queries, paths = load_set(setname, str(tmp_path))
    n_pairs = n_flagged_dev = 0
    for h, path in zip(gold["hmms"], paths):
        prof = O.Profile(path)
        assert prof.M == h["M"] and prof.nseq == h["nseq"]
        for name, seq in queries:
            dsq = prof.abc.digitize(seq)
            r = O.score_pair(prof, dsq)
            hit = h["hits"].get(name)
            n_pairs += 1
            if r["flags"] & 1:
                # documented deviation: reported-or-not / null2 may differ, pre-score must still be close
                n_flagged_dev += 1
                # (amino_extreme: low-complexity W/C repeats, where HMMER's clustering splits the region into many
                #  domains and the scores legitimately differ by much more -- DESIGN.md section 2)
                if hit is not None and setname != "amino_extreme":
                    assert abs(r["score"] - hit["score"]) < 0.15
                continue
            assert r["reported"] == (hit is not None), (setname, name)
            if hit is not None:
                assert O.printed_score(r["score"]) == hit["score"], (setname, name, r, hit)
                assert abs((r["pre_score"] - r["score"]) - hit["bias"]) < 0.11 or (r["flags"] & 2)
                if len(hit["domains"]) == 1 and r["nregions"] == 1:
                    assert tuple(hit["domains"][0][2:4]) == r["env"]
            if name in h["columns"]:
                cols = O.align_pair(prof, dsq)
                assert np.array_equal(cols, np.array(h["columns"][name], dtype=np.int32)), (setname, name)
    assert n_pairs > 0 and n_flagged_dev < n_pairs


def test_columns_match_even_for_flagged_pairs(tmp_path):
    gold, queries, paths = load_set("dna_sub8", str(tmp_path))
    prof = O.Profile(paths[0])
    qd = dict(queries)
    n = 0
    for name, cols in gold["hmms"][0]["columns"].items():
        got = O.align_pair(prof, prof.abc.digitize(qd[name]))
        assert np.array_equal(got, np.array(cols, dtype=np.int32))
        n += len(cols)
    assert n > 5000


def test_forward_equals_backward():
    gold, queries, paths = load_set("amino_small")
    prof = O.Profile(paths[0])
    for name, seq in queries[:6]:
        dsq = prof.abc.digitize(seq)
        for mh in (True, False):
            f = O.forward_nats(prof, dsq, mh)
            b = O.backward_nats(prof, dsq, mh)
            assert abs(f - b) < 1e-9


def test_weights_formula_matches_softmax_form():
    rng = np.random.default_rng(0)
    scores = [float("%.1f" % x) for x in rng.uniform(-5, 300, 40)]
    sizes = [int(x) for x in rng.integers(2, 500, 40)]
    w = O.calculate_weights(list(range(40)), scores, sizes, 10)
    a = np.array(scores) + np.log2(np.array(sizes, dtype=np.float64))
    sm = np.exp2(a - a.max())
    sm /= sm.sum()
    order = np.argsort(-sm, kind="stable")[:10]
    assert [i for i, _ in w] == list(order)
    assert np.allclose([x for _, x in w], sm[order], rtol=1e-12)
    inc = O.adaptive_inclusion(w)
    assert 1 <= len(inc) <= 10 and sum(x for _, x in inc[:-1]) < 0.999


def test_compress_insertions():
    assert O.compress_insertions("--ac-AC-gt-T--g-a-") == "ac---AC-gt-T----ga"
    assert O.compress_insertions("--ac--") == "--ac--"


def test_merge_restatement_matches_reference_python():
    """oracle.merge_rows (closed form) == the reference's ExtendedAlignment.merge_in applied query by query."""
    import gzip, json, os
    from golden_util import GOLDEN
    Mg = json.loads(gzip.open(os.path.join(GOLDEN, "dna_small", "merge_golden.json.gz")).read())
    G = json.loads(gzip.open(os.path.join(GOLDEN, "dna_small", "graph_golden.json.gz")).read())
    names = [n for n, _ in Mg["backbone"]] + Mg["order"]
    rows = [r for _, r in Mg["backbone"]] + [G["queries"][t]["row"] for t in Mg["order"]]
    merged, masked, width = O.merge_rows(rows, G["backbone_length"])
    assert len(merged[0]) == G["backbone_length"] + int(width.sum()) > G["backbone_length"]
    for n, a, b in zip(names, merged, masked):
        assert a == Mg["merged"][n] and b == Mg["masked"][n], n
