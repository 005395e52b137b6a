"""Golden vectors for the "next" row: WITCH's weights + weighted alignment-graph DP, produced by
IMPORTING THE REFERENCE'S OWN PYTHON (witch_msa.gcmm.weighting.calculateWeights and
witch_msa.gcmm.aligner.alignSubQueriesNew, which launches the bundled hmmalign exactly as WITCH does).

Run in the build container only. The reference tree is read-only and writes a file into its package directory on
import, so it is copied to a scratch directory first; `dendropy` (not installed, imported but unused on this path)
is stubbed by name. Output: tests/golden/dna_small/graph_golden.json with, per query,
    weights : the reference's taxon_to_weights entry  [(hmm_idx, w), ...]
    row     : the final aligned row alignSubQueriesNew returns (upper = aligned, lower = insertion, '-' = gap)
computed over the three window profiles of the dna_small set placed on one common backbone coordinate system.
"""
import json
import os
import shutil
import sys
import tempfile
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WINDOW_OFFSETS = [300, 900, 0]   # column windows used for dna_small's three profiles (make_golden.py)
BACKBONE_LENGTH = 2574


def main():
    from golden_util import load_set
    from oracle.make_ref import ref_tool, build as build_ref
    assert build_ref()
    work = tempfile.mkdtemp(prefix="witch_ref_")
    shutil.copytree("/root/reference/witch_msa", os.path.join(work, "witch_msa"))
    stub = os.path.join(work, "dendropy")
    os.makedirs(os.path.join(stub, "datamodel"))
    open(os.path.join(stub, "__init__.py"), "w").write(
        "class Tree: pass\nclass Taxon: pass\nclass DataSet: pass\nclass treecalc: pass\n")
    open(os.path.join(stub, "datamodel", "__init__.py"), "w").write("")
    open(os.path.join(stub, "datamodel", "treemodel.py"), "w").write("class Tree: pass\n")
    open(os.path.join(stub, "datamodel", "taxonmodel.py"), "w").write("class Taxon: pass\n")
    os.environ["HOME"] = os.path.join(work, "home")
    os.makedirs(os.environ["HOME"])
    sys.path.insert(0, work)
    from witch_msa.configs import Configs
    from witch_msa.gcmm import weighting, aligner

    gold, queries, paths = load_set("dna_small", os.path.join(work, "hmms"))
    outdir = os.path.join(work, "out")
    os.makedirs(outdir)
    Configs.outdir = outdir
    Configs.hmmalignpath = ref_tool("hmmalign")
    Configs.use_weight = True
    Configs.keeptemp = False
    Configs.num_hmms = 10
    for name in ("log", "warning", "runtime", "debug", "error"):
        setattr(Configs, name, staticmethod(lambda *a, **k: None))

    class HMM:
        def __init__(self, p, n):
            self.hmm_model_path, self.num_taxa = p, n
    index_to_hmm = {i: HMM(p, gold["hmms"][i]["nseq"]) for i, p in enumerate(paths)}
    retained = {i: tuple(c + WINDOW_OFFSETS[i] for c in gold["hmms"][i]["retained_columns"]) for i in range(len(paths))}
    nongaps = {i: tuple(gold["hmms"][i]["nongaps_per_column"]) for i in range(len(paths))}
    aligner.alignSubQueriesNew.subset_to_retained_columns = retained
    aligner.alignSubQueriesNew.subset_to_nongaps_per_column = nongaps
    lock = threading.Lock()
    out = {"backbone_length": BACKBONE_LENGTH, "window_offsets": WINDOW_OFFSETS, "queries": {}}
    for qi, (taxon, seq) in enumerate(queries):
        scores = [(h, gold["hmms"][h]["hits"][taxon]["score"]) for h in range(len(paths)) if taxon in gold["hmms"][h]["hits"]]
        if not scores:
            continue
        ranked = sorted(scores, key=lambda x: x[1], reverse=True)                       # gcmm/loader.py:318-330
        res = weighting.calculateWeights((taxon, [x[0] for x in ranked], [x[1] for x in ranked],
                                          [index_to_hmm[x[0]].num_taxa for x in ranked]))
        sw = res[taxon]
        query, _, _ = aligner.alignSubQueriesNew(None, BACKBONE_LENGTH, index_to_hmm, lock, 120, taxon, seq.upper(), sw, qi)
        row = query[taxon] if len(query) else None
        out["queries"][taxon] = {"weights": [(int(i), float(w)) for i, w in sw], "row": row}
        print(taxon, len(seq), [(i, round(float(w), 4)) for i, w in sw], None if row is None else len(row))
    with open(os.path.join(HERE, "dna_small", "graph_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
