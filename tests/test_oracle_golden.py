"""CPU: the oracle (oracle/hmm_oracle.c + oracle/oracle.py) against the reference binaries' golden vectors.

Pins: (1) reported-or-not, (2) the printed 1-decimal score, (3) hmmalign column lists, for every pair in
tests/golden/* -- including the pairs whose region fails HMMER's single-domain test and goes through the
stochastic-traceback clustering branch (oracle/hmm_md.c): for those, tests/golden/md_golden.json also pins the cluster
list {i, j, k, m, count of 200 traces} captured from inside the binary and every domain envelope it printed.
"""
import json
import os
import numpy as np
import pytest

from golden_util import SETS, load_set
from oracle import oracle as O


@pytest.mark.parametrize("setname", SETS)
def test_scores_and_columns_match_reference_binaries(setname, tmp_path):
    gold, queries, paths = load_set(setname, str(tmp_path))
    n_pairs = n_flagged = 0
    for h, path in zip(gold["hmms"], paths):
        prof = O.Profile(path)
        assert prof.M == h["M"] and prof.nseq == h["nseq"]
        for name, seq in queries:
            dsq = prof.abc.digitize(seq)
            r = O.score_pair(prof, dsq)
            hit = h["hits"].get(name)
            n_pairs += 1
            n_flagged += r["flags"] & 1          # multi-domain pairs are held to the same assertions as all others
            assert r["reported"] == (hit is not None), (setname, name)
            if hit is not None:
                assert O.printed_score(r["score"]) == hit["score"], (setname, name, r, hit)
                # (bias is printed from pre_score - score of the same float arithmetic; amino_extreme prints biases of
                #  thousands of bits, i.e. with an absolute rounding of the float32 score pair)
                assert abs((r["pre_score"] - r["score"]) - hit["bias"]) < 0.11 + 1e-4 * abs(hit["bias"]) or (r["flags"] & 2)
                # golden.json lists the domains hmmsearch REPORTED (domain E-value <= 10); each must be one of the envelopes
                envs = {(e[0], e[1]) for e in r["envelopes"]}
                for d in hit["domains"]:
                    assert (d[2], d[3]) in envs, (setname, name, d, envs)
            if name in h["columns"]:
                cols = O.align_pair(prof, dsq)
                assert np.array_equal(cols, np.array(h["columns"][name], dtype=np.int32)), (setname, name)
    assert n_pairs > 0
    if setname in ("dna_small", "dna_sub8", "dna_full", "amino_extreme"):
        assert n_flagged >= 10   # these sets do exercise the multi-domain branch


@pytest.mark.parametrize("setname", SETS)
def test_multidomain_branch_matches_reference_binary(setname, tmp_path):
    """oracle/hmm_md.c against what was captured from inside hmmsearch (tests/golden/make_golden_md.py): the clusters of
    the 200 sampled traces, every domain envelope, the printed per-sequence score and bias, the printed per-domain
    score and bias."""
    from golden_util import GOLDEN
    G = json.load(open(os.path.join(GOLDEN, "md_golden.json")))[setname]
    gold, queries, paths = load_set(setname, str(tmp_path))
    qd = dict(queries)
    n = 0
    for h, path in enumerate(paths):
        prof = O.Profile(path)
        for name, ref in G[str(h)].items():
            L = len(qd[name])
            r = O.score_pair(prof, prof.abc.digitize(qd[name]))
            assert r["flags"] & 1
            assert sorted(r["clusters"]) == sorted(tuple(c) for c in ref["clusters"]), (setname, h, name)
            assert sorted((e[0], e[1]) for e in r["envelopes"]) == sorted((d[2], d[3]) for d in ref["domains"]), (setname, h, name)
            assert r["reported"] == (ref["score"] is not None)
            if r["reported"]:
                assert O.printed_score(r["score"]) == ref["score"], (setname, h, name, r["score"], ref)
            p1 = L / (L + 1.0)
            nullsc = L * np.log(p1) + np.log(1.0 - p1)
            for e, d in zip(sorted(r["envelopes"]), sorted(ref["domains"], key=lambda d: (d[2], d[3]))):
                i, j, envsc, corr, _ = e
                y = corr + np.log(1.0 / 256.0)
                dombias = y + np.log1p(np.exp(-y)) if y > 40 else np.log1p(np.exp(y))
                bits = (envsc + (L - (j - i + 1)) * np.log(L / (L + 3.0)) - (nullsc + dombias)) / np.log(2.0)
                assert abs(bits - d[0]) < 0.051 + 2e-5 * abs(dombias), (setname, h, name, e, d)
                assert abs(dombias / np.log(2.0) - d[1]) < 0.051 + 2e-5 * abs(dombias), (setname, h, name, e, d)
            n += 1
    if setname != "amino_small":
        assert n >= 10


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
