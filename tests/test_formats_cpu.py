"""CPU: on-disk formats at the edges of the hot path (witch_b200/formats.py)."""
import gzip
import os

import numpy as np

from golden_util import load_set
from witch_b200 import formats as F


def _make_dir(tmp_path, paths, taxa_sets, backbone):
    root = tmp_path / "tree_decomp" / "root"
    for i, (p, taxa) in enumerate(zip(paths, taxa_sets)):
        d = root / ("A_0_%d" % i)
        d.mkdir(parents=True)
        (d / ("hmmbuild.model.A_0_%d" % i)).write_text(open(p).read())
        with open(d / ("hmmbuild.input.A_0_%d.fasta" % i), "w") as f:
            for t in taxa:
                f.write(">%s\n%s\n" % (t, backbone[t]))
    bb = tmp_path / "backbone.fasta"
    with open(bb, "w") as f:
        for t, s in backbone.items():
            f.write(">%s\n%s\n" % (t, s))
    return str(root), str(bb)


def test_directory_layout_scores_weights_checkpoint_roundtrip(tmp_path):
    gold, queries, paths = load_set("dna_small", str(tmp_path / "hmm"))
    backbone = {"t0": "AC-GT-A", "t1": "A--GTCA", "t2": "-C-G--A", "t3": "ACTG--A"}
    root, bb = _make_dir(tmp_path, paths, [["t0", "t1"], ["t2"], ["t1", "t2", "t3"]], backbone)
    i2h = F.getAlignmentSubsets(root)
    assert sorted(i2h) == [0, 1, 2]
    assert [i2h[i].num_taxa for i in range(3)] == [h["nseq"] for h in gold["hmms"]]
    assert all(os.path.basename(i2h[i].hmm_model_path) == "hmmbuild.model.A_0_%d" % i for i in range(3))
    ret, ng, B = F.obtainRetainedColumns(bb, i2h)
    assert B == 7
    assert ret[0] == (0, 1, 3, 4, 5, 6) and ng[0] == (2, 1, 2, 2, 1, 2)       # columns that are not all-gap in {t0,t1}
    assert ret[1] == (1, 3, 6) and ng[1] == (1, 1, 1)
    assert ret[2] == (0, 1, 2, 3, 4, 5, 6) and ng[2] == (2, 2, 1, 3, 1, 1, 3)
    # bit-score files: what the reference's readHMMSearch (eval of a dict of (evalue, score)) expects
    names = [n for n, _ in queries[:5]]
    scores = np.array([[10.04, np.nan, -3.26], [5.55, 7.0, np.nan], [np.nan] * 3, [1.0, 1.0, 1.0], [99.95, 0.0, 2.5]], dtype=np.float32)
    rep = ~np.isnan(scores)
    files = F.writeHMMSearchResults(i2h, names, scores, rep)
    assert os.path.basename(files[0]) == "hmmsearch.results.A_0_0.fragment_chunk_0"
    assert eval(open(files[0]).read())[names[0]] == (0.0, 10.0)                  # the reference reads with eval
    ranked = F.readAndRankBitscore(i2h, renamed_taxa={names[3]: "renamed"})
    assert ranked[names[0]] == [(0, 10.0), (2, -3.3)] and names[2] not in ranked
    assert ranked["renamed"] == [(0, 1.0), (1, 1.0), (2, 1.0)]
    assert ranked[names[4]][0] == (0, 100.0) or ranked[names[4]][0] == (0, 99.9)   # %.1f of the float32 value
    # weights.txt
    t2w = {"a_b": ((3, 0.75), (1, 0.25)), "q": ((0, 1.0),)}   # (the reference's reader splits at ':', so no colons in names)
    wp = str(tmp_path / "weights.txt")
    F.writeWeightsToLocal(t2w, wp)
    assert open(wp).read().splitlines()[1] == "q:((0, 1.0),)"
    assert F.readWeightsFromLocal(wp) == t2w
    # checkpoint_alignments.txt.gz: appended batches, one "taxon<TAB>row" line each
    cp = str(tmp_path / "checkpoint_alignments.txt.gz")
    F.writeCheckpointAlignments(cp, {"q1": "ac-GT", "q 2": "--ACg"})
    F.writeCheckpointAlignments(cp, {"q3": "A"})
    assert gzip.open(cp, "rb").read().decode() == "q1\tac-GT\nq 2\t--ACg\nq3\tA\n"
    assert F.readCheckpointAlignments(cp) == {"q1": "ac-GT", "q 2": "--ACg", "q3": "A"}


def test_writers_are_byte_identical_to_files_the_reference_readers_accepted(tmp_path):
    """tests/golden/make_golden_c1.py wrote these files with formats.py and parsed them back with THE REFERENCE'S OWN
    readHMMSearch (gcmm/loader.py:277-294), readWeightsFromLocal (gcmm/weighting.py:184-194) and
    readOneCheckpointAlignment (gcmm/loader.py:95-111), asserting the content survives; the exact texts are pinned here."""
    import json
    from golden_util import GOLDEN
    fmt = json.loads(gzip.open(os.path.join(GOLDEN, "c1", "c1_golden.json.gz")).read())["formats"]
    gold, queries, paths = load_set("dna_small", str(tmp_path / "hmm"))
    backbone = {"t0": "AC-GT-A", "t1": "A--GTCA"}
    root, bb = _make_dir(tmp_path, paths[:2], [["t0", "t1"], ["t1"]], backbone)
    i2h = F.getAlignmentSubsets(root)
    files = F.writeHMMSearchResults(i2h, fmt["names"], np.array(fmt["scores"], dtype=np.float32), np.array(fmt["reported"], dtype=bool))
    for i in (0, 1):
        assert open(files[i]).read() == fmt["hmmsearch_files"][str(i)]
        assert F.readHMMSearch(i2h[i]) == {t: [tuple(x) for x in v] for t, v in fmt["hmmsearch_parsed"][str(i)].items()}
    wp = str(tmp_path / "w.txt")
    order = [ln.split(":")[0] for ln in fmt["weights_file"].splitlines()]   # (the golden json is key-sorted; the file is not)
    F.writeWeightsToLocal({t: tuple(tuple(x) for x in fmt["weights"][t]) for t in order}, wp)
    assert open(wp).read() == fmt["weights_file"]
    cp = str(tmp_path / "cp.txt.gz")
    rows = fmt["checkpoint_rows"]
    F.writeCheckpointAlignments(cp, {"SHFB": rows["SHFB"]}, append=False)
    F.writeCheckpointAlignments(cp, {k: rows[k] for k in ("Q_b", "t\tab")}, append=True)
    assert gzip.open(cp, "rb").read().decode() == fmt["checkpoint_text"]
    assert F.readCheckpointAlignments(cp) == rows
