"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the reference binaries' golden
vectors. Tolerances (BASELINE.json north_star): bit scores within 0.01 bits of the float64 oracle; identical
reported sets; identical top-k HMM ranking and weights (1e-12 relative, the reference's own float64 summation-order
noise); alignment columns bit-exact except on documented float near-ties (<= 1 residue in 2000)."""
import os
import sys

import numpy as np
import pytest

from golden_util import SETS, load_set
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
pytestmark = pytest.mark.gpu

SCORE_TOL_BITS = 0.01


@pytest.fixture(scope="module")
def wb():
    import witch_b200
    from witch_b200 import _lib
    assert _lib.load().witch_device_count() > 0, "CUDA library loaded but no device: no CPU fallback exists"
    return witch_b200


def _oracle_scores(profs, queries):
    out = {}
    for qi, (_, s) in enumerate(queries):
        for h, p in enumerate(profs):
            out[(qi, h)] = O.score_pair(p, p.abc.digitize(s))
    return out


@pytest.mark.parametrize("setname", SETS)
def test_scores_weights_columns_vs_oracle_and_golden(wb, setname, tmp_path):
    gold, queries, paths = load_set(setname, str(tmp_path))
    profs = [O.Profile(p) for p in paths]
    E = wb.EHMM(paths)
    Q = wb.Queries(E, [s for _, s in queries])
    assert list(E.M) == [h["M"] for h in gold["hmms"]] and list(E.nseq) == [h["nseq"] for h in gold["hmms"]]
    sc, rep, pre, fl = wb.score(E, Q)
    ora = _oracle_scores(profs, queries)
    for (qi, h), r in ora.items():
        assert bool(rep[qi, h]) == r["reported"], (setname, qi, h, r)
        assert abs(pre[qi, h] - r["pre_score"]) < SCORE_TOL_BITS
        assert (fl[qi, h] & 1) == (r["flags"] & 1)
        if r["reported"]:
            # amino_extreme: null2 corrections of up to 2,000 bits (W/C repeats); FP32 sums of 700 x M posteriors carry
            # ~1e-5 relative error, so the tolerance there is 0.01 bits + 2e-5 of the correction itself
            tol = SCORE_TOL_BITS + (2e-5 * abs(r["pre_score"] - r["score"]) if setname == "amino_extreme" else 0.0)
            assert abs(sc[qi, h] - r["score"]) < tol, (setname, qi, h, sc[qi, h], r)
        else:
            assert np.isnan(sc[qi, h])
    # printed 1-decimal scores against the reference binary, multi-domain pairs included (values within 1e-3 of a
    # rounding boundary may legitimately print differently); the reported sets must be hmmsearch's
    names = [n for n, _ in queries]
    nprint = nmd = 0
    for h, hg in enumerate(gold["hmms"]):
        assert {names[qi] for qi in range(Q.n) if rep[qi, h]} == set(hg["hits"].keys()), (setname, h)
        for n, hit in hg["hits"].items():
            qi = names.index(n)
            nmd += int(fl[qi, h] & 1)
            x = float(sc[qi, h]) * 10.0
            if abs(x - np.floor(x) - 0.5) < 0.01:
                continue
            assert O.printed_score(float(sc[qi, h])) == hit["score"], (setname, n, h, sc[qi, h], hit)
            nprint += 1
    assert nprint > 0
    if setname in ("dna_small", "dna_sub8", "dna_full", "amino_extreme"):
        assert nmd >= 10   # these sets exercise the multi-domain branch (stochastic-trace clustering) on the device
    # weights / top-k
    idx, w, cnt = wb.weights_topk(E, sc, rep, 10, 1)
    for qi in range(Q.n):
        ss = {h: O.printed_score(float(sc[qi, h])) for h in range(E.n) if rep[qi, h]}
        if not ss:
            assert cnt[qi] == 0 and (idx[qi] == -1).all()
            continue
        ranked = O.rank_bitscores(ss)
        ow = O.calculate_weights([h for h, _ in ranked], [x for _, x in ranked], [int(E.nseq[h]) for h, _ in ranked], 10)
        assert cnt[qi] == len(ow)
        got = [(int(idx[qi, j]), float(w[qi, j])) for j in range(cnt[qi])]
        for (gi, gw), (oi, owt) in zip(got, ow):
            assert abs(gw - owt) <= 1e-12 * owt + 1e-300
        assert sorted(i for i, _ in got) == sorted(i for i, _ in ow)
        assert abs(sum(x for _, x in got) - 1.0) < 1e-9 or len(ss) > 10
    # alignment columns against hmmalign's (golden) and the oracle's
    pq, ph, exp = [], [], []
    for h, hg in enumerate(gold["hmms"]):
        for n, cols in hg["columns"].items():
            pq.append(names.index(n)); ph.append(h); exp.append(np.array(cols, dtype=np.int32))
    got = wb.align(E, Q, pq, ph)
    nres = nbad = 0
    for a, b in zip(got, exp):
        assert len(a) == len(b)
        nres += len(b); nbad += int((a != b).sum())
    assert nbad <= nres // 2000, (setname, nbad, nres)


def test_profile_cache_gives_identical_results(wb, tmp_path):
    """An eHMM created from the serialised profile cache scores and aligns exactly like one parsed from the HMM text."""
    gold, queries, paths = load_set("dna_sub8", str(tmp_path))
    cache = str(tmp_path / "witch_b200.profiles")
    E0 = wb.EHMM(paths)
    E1 = wb.EHMM(paths, cache=cache)
    E2 = wb.EHMM(paths, cache=cache)
    assert not E1.cache_hit and E2.cache_hit
    assert list(E2.M) == list(E0.M) and list(E2.nseq) == list(E0.nseq)
    seqs = [s for _, s in queries[:40]]
    a = wb.score(E0, wb.Queries(E0, seqs))
    b = wb.score(E2, wb.Queries(E2, seqs))
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    pq = [qi for qi in range(len(seqs)) if a[1][qi, 0]][:10]
    ca = wb.align(E0, wb.Queries(E0, seqs), pq, [0] * len(pq))
    cb = wb.align(E2, wb.Queries(E2, seqs), pq, [0] * len(pq))
    assert all(np.array_equal(x, y) for x, y in zip(ca, cb))


def test_edge_cases(wb, tmp_path):
    gold, queries, paths = load_set("dna_small", str(tmp_path))
    E = wb.EHMM(paths)
    prof = O.Profile(paths[0])
    base = queries[0][1]
    seqs = ["", "A", "ACGT", base, base.lower(), base[:50] + "NNNRYK" + base[50:], base, "T" * 300]
    Q = wb.Queries(E, seqs)
    sc, rep, pre, fl = wb.score(E, Q)
    assert not rep[0].any() and np.isnan(sc[0]).all()          # empty query: never reported
    assert np.array_equal(rep[3], rep[4]) and np.allclose(sc[3], sc[4], equal_nan=True)  # case-insensitive
    assert np.allclose(sc[3], sc[6], equal_nan=True)            # duplicates score identically
    for qi in (1, 2, 5, 7):
        r = O.score_pair(prof, prof.abc.digitize(seqs[qi]))
        assert bool(rep[qi, 0]) == r["reported"]
        if r["reported"]:
            assert abs(sc[qi, 0] - r["score"]) < SCORE_TOL_BITS
    cols = wb.align(E, Q, [0, 1, 5, 3], [0, 0, 0, 1])
    assert len(cols[0]) == 0 and len(cols[1]) == 1
    ref = O.align_pair(prof, prof.abc.digitize(seqs[5]))
    assert np.array_equal(cols[2], ref)
    with pytest.raises(wb.WitchError):
        wb.Queries(E, ["AC-GT"])       # gaps are not valid in unaligned queries
    with pytest.raises(wb.WitchError):
        wb.align(E, Q, [99], [0])      # pair index out of range


@pytest.mark.parametrize("root_len,degenerate", [(600, False), (1500, False), (1500, True), (1750, False), (1750, True), (2600, False)])
def test_parser_classes_vs_oracle(wb, tmp_path, root_len, degenerate):
    """Every launch class of the packed two-query parser (witch_abi.cu:s_classes) against the float64 oracle: C = 4 pairs
    (<= 1,024 nodes), 13 x 128 (<= 1,664), 16 x 128 (<= 2,048, plain ACGT only), 13 x 256 (<= 3,328), and the fall-back
    classes when degenerate symbols leave no room for two CTAs' tables. An odd number of queries of very unequal
    lengths: the last one is paired with itself, partners differ in length by up to 10x, one query is empty."""
    import synth
    wl = synth.make_workload(str(tmp_path), alphabet="dna", n_total=150, n_backbone=40, root_len=root_len, decomp=10,
                             frag_frac=0.6, frag_mean=max(150, root_len // 6), seed=31 + root_len)
    paths = wl["hmm_paths"][:6]
    rng = np.random.default_rng(root_len)
    seqs = [wl["seqs"][i] for i in rng.choice(len(wl["seqs"]), size=20, replace=False)]
    seqs += [seqs[0][:37], seqs[1][:5], "", seqs[2][: len(seqs[2]) // 2]]
    seqs.append(seqs[3])   # 25 queries: odd
    if degenerate:   # 11 IUPAC codes on top of ACGT -> 15 distinct symbols: the 13 x 128 class no longer fits twice per SM
        codes = "NRYKMSWBDHV"
        for z in range(0, 12):
            a = list(seqs[z])
            for pos in rng.integers(0, len(a), 6):
                a[pos] = codes[(z + pos) % len(codes)]
            seqs[z] = "".join(a)
    E = wb.EHMM(paths)
    Q = wb.Queries(E, seqs)
    sc, rep, pre, fl = wb.score(E, Q)
    profs = [O.Profile(p) for p in paths]
    worst = wpre = 0.0
    for qi, s in enumerate(seqs):
        for h, p in enumerate(profs):
            if len(s) == 0:
                assert not rep[qi, h]
                continue
            r = O.score_pair(p, p.abc.digitize(s))
            assert bool(rep[qi, h]) == r["reported"], (root_len, qi, h)
            assert (int(fl[qi, h]) & 1) == (r["flags"] & 1), (root_len, qi, h)
            wpre = max(wpre, abs(float(pre[qi, h]) - r["pre_score"]))
            if r["reported"]:
                worst = max(worst, abs(float(sc[qi, h]) - r["score"]))
    assert wpre < SCORE_TOL_BITS and worst < SCORE_TOL_BITS, (root_len, wpre, worst)
    # the same queries in another order (other partners, other pair slots) give the same numbers
    perm = rng.permutation(len(seqs))
    sc2, rep2, pre2, _ = wb.score(E, wb.Queries(E, [seqs[i] for i in perm]))
    ne = np.array([len(seqs[i]) > 0 for i in perm])   # (an empty query has no Forward score: its `pre` entry is not defined)
    assert np.array_equal(rep2, rep[perm]) and np.allclose(sc2, sc[perm], equal_nan=True, atol=1e-5)
    assert np.allclose(pre2[ne], pre[perm][ne], atol=1e-5)


def test_properties_at_scale(wb, tmp_path):
    """Size-independent properties on a workload the oracle could not finish quickly."""
    import synth
    wl = synth.make_workload(str(tmp_path), alphabet="dna", n_total=1500, n_backbone=200, root_len=900, decomp=10,
                             frag_frac=0.5, frag_mean=300, seed=11)
    E = wb.EHMM(wl["hmm_paths"])
    seqs = wl["seqs"][:600]
    Q = wb.Queries(E, seqs)
    sc, rep, pre, fl = wb.score(E, Q)
    # (1) Forward == Backward totals for both the multihit and unihit engines
    rng = np.random.default_rng(0)
    pq = rng.integers(0, Q.n, 300).astype(np.int32)
    ph = rng.integers(0, E.n, 300).astype(np.int32)
    for mh in (True, False):
        f, b = wb.debug_fwdbwd(E, Q, pq, ph, mh)
        assert np.all(np.abs(f - b) < 2e-3 + 2e-6 * np.abs(f)), np.abs(f - b).max()
    # (2) permutation invariance: scoring a shuffled query set gives the same numbers
    perm = rng.permutation(Q.n)
    Q2 = wb.Queries(E, [seqs[i] for i in perm])
    sc2, rep2, _, _ = wb.score(E, Q2)
    assert np.array_equal(rep2, rep[perm]) and np.allclose(sc2, sc[perm], equal_nan=True, atol=1e-5)
    # (3) every query is reported by the root HMM (it contains the whole backbone) and weights sum to one
    assert rep[:, 0].mean() > 0.99
    idx, w, cnt = wb.weights_topk(E, sc, rep, 10, 1)
    full = cnt == 10
    assert np.all(w.sum(1)[~full & (cnt > 0)] > 1 - 1e-9) and np.all(w.sum(1) <= 1 + 1e-9)
    assert np.all(np.diff(w, axis=1) <= 1e-18)
    # (4) alignment columns are strictly increasing over matched residues and inside the model
    keep = np.minimum(cnt, 2)
    aq = np.repeat(np.arange(Q.n), keep).astype(np.int32)[:800]
    ah = np.concatenate([idx[q, :keep[q]] for q in range(Q.n)]).astype(np.int32)[:800]
    cols = wb.align(E, Q, aq, ah)
    for c, q, h in zip(cols, aq, ah):
        m = c[c >= 0]
        assert len(c) == len(seqs[q]) and np.all(np.diff(m) > 0) and (len(m) == 0 or m[-1] < E.M[h])
    # (5) spot-check against the oracle
    for z in rng.integers(0, len(aq), 12):
        p = O.Profile(wl["hmm_paths"][ah[z]])
        d = p.abc.digitize(seqs[aq[z]])
        r = O.score_pair(p, d)
        assert abs(sc[aq[z], ah[z]] - r["score"]) < SCORE_TOL_BITS
        assert (cols[z] != O.align_pair(p, d)).sum() <= 1


def test_mirror_interface_end_to_end(wb, tmp_path):
    """BatchedSearch mirrors the reference's function contracts (types and content)."""
    from witch_b200.gcmm import BatchedSearch
    gold, queries, paths = load_set("dna_small", str(tmp_path))
    rt = str(tmp_path / "runtime_breakdown.txt")
    bs = BatchedSearch(paths, num_hmms=10, runtime_path=rt, profile_cache=str(tmp_path / "witch_b200.profiles"))
    bs.search([n for n, _ in queries], [s for _, s in queries])
    ranked = bs.rankBitscores()
    t2w = bs.writeWeights()
    bb = bs.getBackbones(t2w)
    lines = open(rt).read().splitlines()   # the reference's runtime_breakdown.txt format: "(tag) Time to ... (s): x"
    assert [ln.split(")")[0] for ln in lines] == ["(gpu_load", "(gpu_score", "(gpu_weights", "(gpu_align"]
    assert all(" (s): " in ln and float(ln.rsplit(": ", 1)[1]) >= 0 for ln in lines)
    names = [n for n, _ in queries]
    for h, hg in enumerate(gold["hmms"]):
        res = bs.hmmsearch_results(h)
        for n, hit in hg["hits"].items():
            assert n in res
            assert abs(res[n][1] - hit["score"]) <= 0.1 + 1e-9
    for t, sw in t2w.items():
        assert isinstance(sw, tuple) and all(isinstance(i, int) and isinstance(x, float) for i, x in sw)
        assert [x[1] for x in ranked[t]] == sorted((x[1] for x in ranked[t]), reverse=True)
        log, wmap, s2c = bb[t]
        assert log.startswith(t + "\tpassed to main pipeline with top ")
        for h, cols in s2c.items():
            assert len(cols) == len(dict(queries)[t]) and h in wmap
            if t in gold["hmms"][h]["columns"]:
                assert sum(a != b for a, b in zip(cols, gold["hmms"][h]["columns"][t])) <= 2


def test_device_pipeline_matches_staged_calls(wb, tmp_path):
    import torch
    from witch_b200.gcmm import DevicePipeline
    gold, queries, paths = load_set("amino_small", str(tmp_path))
    E = wb.EHMM(paths)
    Q = wb.Queries(E, [s for _, s in queries])
    res = DevicePipeline(E, k=10).run(Q)
    torch.cuda.synchronize()
    sc, rep, _, _ = wb.score(E, Q)
    assert np.allclose(res["scores"].cpu().numpy(), sc, equal_nan=True)
    idx, w, cnt = wb.weights_topk(E, sc, rep, 10, 1)
    assert np.array_equal(res["idx"].cpu().numpy(), idx) and np.allclose(res["w"].cpu().numpy(), w)
    cols = wb.align(E, Q, res["pair_q"], res["pair_h"])
    flat = res["cols"].cpu().numpy()
    for p, c in enumerate(cols):
        assert np.array_equal(flat[res["col_off"][p]:res["col_off"][p + 1]], c)


def _graph_golden():
    import gzip, json
    from golden_util import GOLDEN
    return json.loads(gzip.open(os.path.join(GOLDEN, "dna_small", "graph_golden.json.gz")).read())


def test_graph_dp_matches_reference_python_rows(wb, tmp_path):
    """Device alignment-graph DP fed with the reference's own weights and hmmalign columns must reproduce, character
    for character, the rows the reference's alignSubQueriesNew returned (tests/golden/make_golden_graph.py)."""
    gold, queries, paths = load_set("dna_small", str(tmp_path))
    G = _graph_golden()
    E = wb.EHMM(paths)
    ret = [[c + G["window_offsets"][i] for c in gold["hmms"][i]["retained_columns"]] for i in range(3)]
    ng = [gold["hmms"][i]["nongaps_per_column"] for i in range(3)]
    seqs, pb, ph, pw, cols, exp = [], [0], [], [], [], []
    for taxon, seq in queries:
        if taxon not in G["queries"]:
            continue
        sw = [tuple(x) for x in G["queries"][taxon]["weights"]]
        for h, w in O.adaptive_inclusion(sw):
            ph.append(h); pw.append(dict(sw)[h]); cols.append(gold["hmms"][h]["columns"][taxon])
        pb.append(len(ph)); seqs.append(seq.upper()); exp.append(G["queries"][taxon]["row"])
    rows = wb.graph_align(E, seqs, pb, ph, pw, cols, ret, ng, G["backbone_length"])
    assert len(rows) == len(exp) >= 50
    for r, e in zip(rows, exp):
        assert r == e


def test_graph_dp_full_device_path_vs_oracle(wb, tmp_path):
    """score -> weights -> align -> graph DP entirely through the CUDA path, against the oracle's restatement."""
    from witch_b200.gcmm import BatchedSearch
    gold, queries, paths = load_set("dna_small", str(tmp_path))
    G = _graph_golden()
    ret = {i: [c + G["window_offsets"][i] for c in gold["hmms"][i]["retained_columns"]] for i in range(3)}
    ng = {i: gold["hmms"][i]["nongaps_per_column"] for i in range(3)}
    bs = BatchedSearch(paths, num_hmms=10)
    bs.search([n for n, _ in queries], [s for _, s in queries])
    t2w = bs.writeWeights()
    rows = bs.alignSubQueriesNew(G["backbone_length"], ret, ng, t2w)
    bb = bs.getBackbones(t2w)
    nsame_ref = 0
    for taxon, seq in queries:
        if taxon not in t2w:
            assert taxon not in rows
            continue
        _, wmap, s2c = bb[taxon]
        want = O.compress_insertions(O.graph_align(seq.upper(), G["backbone_length"], wmap, s2c, ret, ng))
        assert rows[taxon] == want, taxon
        assert len(rows[taxon]) >= G["backbone_length"]
        assert rows[taxon].replace("-", "").upper() == seq.upper()
        if taxon in G["queries"] and rows[taxon] == G["queries"][taxon]["row"]:
            nsame_ref += 1
    assert nsame_ref >= len(G["queries"]) - 1   # documented: one FP32 near-tie alignment (2 of 35,293 residues)


def test_merge_matches_reference_python(wb, tmp_path):
    """Device transitivity merge fed with the reference's own rows must reproduce, character for character, what the
    reference's ExtendedAlignment.merge_in / remove_insertion_columns produced (tests/golden/make_golden_merge.py)."""
    import gzip, json
    from golden_util import GOLDEN
    from witch_b200.gcmm import mergeAlignmentsCollapsed
    gold, queries, paths = load_set("dna_small", str(tmp_path))
    G = _graph_golden()
    Mg = json.loads(gzip.open(os.path.join(GOLDEN, "dna_small", "merge_golden.json.gz")).read())
    E = wb.EHMM(paths)
    qrows = {t: G["queries"][t]["row"] for t in Mg["order"]}
    out = str(tmp_path / "aligned.fasta")
    full, mask = mergeAlignmentsCollapsed(E, [tuple(x) for x in Mg["backbone"]], qrows, G["backbone_length"], outpath=out)
    assert list(full.keys()) == [n for n, _ in Mg["backbone"]] + Mg["order"]
    for n in full:
        assert full[n] == Mg["merged"][n], n
        assert mask[n] == Mg["masked"][n], n
    assert os.path.exists(out) and os.path.exists(str(tmp_path / "aligned.masked.fasta"))
    # oracle restatement agrees too, and an empty query set leaves the backbone untouched
    m2, k2, w2 = O.merge_rows([r for _, r in Mg["backbone"]] + list(qrows.values()), G["backbone_length"])
    assert m2 == list(full.values()) and k2 == list(mask.values())
    f0, m0 = mergeAlignmentsCollapsed(E, [tuple(x) for x in Mg["backbone"]], {}, G["backbone_length"])
    assert list(f0.values()) == [r for _, r in Mg["backbone"]] == list(m0.values())


def test_next_rows_at_scale_vs_oracle(wb, tmp_path):
    """score -> weights -> align -> graph DP -> transitivity merge on a workload the golden sets do not cover (longer
    models, 40+ HMMs, hundreds of queries): properties that hold for every row, the oracle on a sample, and the
    merged alignment against the oracle's closed form."""
    import synth
    from witch_b200.gcmm import BatchedSearch, mergeAlignmentsCollapsed
    wl = synth.make_workload(str(tmp_path), alphabet="dna", n_total=1500, n_backbone=200, root_len=900, decomp=10,
                             frag_frac=0.5, frag_mean=300, seed=11)
    B = wl["backbone_length"]
    names, seqs = wl["names"][:400], wl["seqs"][:400]
    bs = BatchedSearch(wl["hmm_paths"], num_hmms=10)
    bs.search(names, seqs)
    t2w = bs.writeWeights()
    ret = {h: wl["retained_columns"][h] for h in range(len(wl["hmm_paths"]))}
    ng = {h: wl["nongaps_per_column"][h] for h in range(len(wl["hmm_paths"]))}
    rows = bs.alignSubQueriesNew(B, ret, ng, t2w)
    assert len(rows) == len(t2w) > 390
    qd = dict(zip(names, seqs))
    for t, r in rows.items():
        assert r.replace("-", "").upper() == qd[t]                       # every residue exactly once, in order
        assert sum(1 for ch in r if not ("a" <= ch <= "z")) == B          # exactly one cell per backbone column
    bb = bs.getBackbones(t2w)
    rng = np.random.default_rng(5)
    for t in rng.choice(sorted(rows), 25, replace=False):
        _, wmap, s2c = bb[t]
        want = O.compress_insertions(O.graph_align(qd[t], B, wmap, s2c, ret, ng))
        assert rows[t] == want, t
    backbone = [("bb0", "A" * B), ("bb1", "-" * (B // 2) + "C" * (B - B // 2))]
    full, mask = mergeAlignmentsCollapsed(bs.ehmm, backbone, rows, B)
    m2, k2, w2 = O.merge_rows([r for _, r in backbone] + list(rows.values()), B)
    assert list(full.values()) == m2 and list(mask.values()) == k2
    width = B + int(w2.sum())
    assert all(len(r) == width for r in full.values()) and all(len(r) == B for r in mask.values())
    for t, r in rows.items():
        assert full[t].replace("-", "").upper() == qd[t] and mask[t] == "".join(ch for ch in r if not ("a" <= ch <= "z"))


def test_limits_long_queries_and_model_size(wb, tmp_path):
    """Sizes at the edges: queries far longer than any length class boundary (residue staging and scratch are sized per
    launch), a 6,000-node model runs through the slow parser class, and a model beyond the supported 8,192 nodes must be refused
    loudly, not mis-scored."""
    import synth
    gold, queries, paths = load_set("dna_small", str(tmp_path))
    rng = np.random.default_rng(9)
    base = queries[0][1].upper()
    long1 = "".join(rng.choice(list("ACGT"), 2500)) + base + "".join(rng.choice(list("ACGT"), 2600))   # 5,000+ nt, one hit
    long2 = (base + "".join(rng.choice(list("ACGT"), 700))) * 3                                          # three copies
    E = wb.EHMM(paths[:2])
    Q = wb.Queries(E, [long1, long2, base])
    sc, rep, pre, fl = wb.score(E, Q)
    for h in range(2):
        prof = O.Profile(paths[h])
        for qi, s in enumerate([long1, long2, base]):
            r = O.score_pair(prof, prof.abc.digitize(s))
            assert bool(rep[qi, h]) == r["reported"], (qi, h)
            assert abs(pre[qi, h] - r["pre_score"]) < SCORE_TOL_BITS
            if r["reported"] and r["nregions"] <= 6:
                assert abs(sc[qi, h] - r["score"]) < SCORE_TOL_BITS, (qi, h, sc[qi, h], r)
    # more than 6 envelopes for one pair: the first 6 are scored, WITCH_FLAG_ENVCAP (4) says so
    hits = gold["hmms"][0]["hits"]
    best = max(hits, key=lambda n: hits[n]["score"])
    d0 = hits[best]["domains"][0]
    seg = dict(queries)[best].upper()[d0[2] - 1:d0[3]][:150]
    many = "".join(seg + "".join(rng.choice(list("ACGT"), 200)) for _ in range(9))
    Q9 = wb.Queries(E, [many])
    sc9, rep9, pre9, fl9 = wb.score(E, Q9)
    p0 = O.Profile(paths[0])
    r9 = O.score_pair(p0, p0.abc.digitize(many))
    assert r9["nregions"] > 6 and abs(pre9[0, 0] - r9["pre_score"]) < SCORE_TOL_BITS
    assert (fl9[0, 0] & 4) and rep9[0, 0] and np.isfinite(sc9[0, 0])
    cols = wb.align(E, Q, [0, 1], [0, 0])
    prof = O.Profile(paths[0])
    for c, s in zip(cols, [long1, long2]):
        ref = O.align_pair(prof, prof.abc.digitize(s))
        assert len(c) == len(s) and int((c != ref).sum()) <= 1
    # a 6,000-node model (slow parser class, own wave launches) still scores and aligns like the oracle
    M = 6000
    cons = rng.integers(0, 4, M)
    counts = np.zeros((M, 4)); counts[np.arange(M), cons] = 1.0
    tc = np.zeros((M + 1, 4)); tc[:, 0] = 1.0
    mid = str(tmp_path / "mid.hmm")
    synth.write_hmm(mid, "mid", counts, tc, 1, synth.DNA)
    Em = wb.EHMM([mid, paths[0]])
    frag = "".join(synth.DNA[c] for c in cons[2000:2600])
    qs = [frag, frag[:200] + "ACGTTGCA" * 5 + frag[200:], base]
    Qm = wb.Queries(Em, qs)
    scm, repm, prem, flm = wb.score(Em, Qm)
    pm = O.Profile(mid)
    for qi, s_ in enumerate(qs):
        r = O.score_pair(pm, pm.abc.digitize(s_))
        assert bool(repm[qi, 0]) == r["reported"] and abs(prem[qi, 0] - r["pre_score"]) < SCORE_TOL_BITS
        if r["reported"]:
            assert abs(scm[qi, 0] - r["score"]) < SCORE_TOL_BITS, (qi, scm[qi, 0], r)
    cm = wb.align(Em, Qm, [0, 1], [0, 0])
    for c, s_ in zip(cm, qs[:2]):
        assert int((c != O.align_pair(pm, pm.abc.digitize(s_))).sum()) <= 1
    # a 9,000-node model is outside the supported range of the parser kernels
    M = 9000
    cons = rng.integers(0, 4, M)
    counts = np.zeros((M, 4)); counts[np.arange(M), cons] = 1.0
    tc = np.zeros((M + 1, 4)); tc[:, 0] = 1.0
    big = str(tmp_path / "big.hmm")
    synth.write_hmm(big, "big", counts, tc, 1, synth.DNA)
    Eb = wb.EHMM([big])
    Qb = wb.Queries(Eb, ["ACGT" * 50])
    with pytest.raises(wb.WitchError, match="8192"):
        wb.score(Eb, Qb)
