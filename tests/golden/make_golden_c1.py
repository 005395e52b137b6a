"""End-to-end golden of BASELINE config c1 (the bundled nucleotide example) produced by RUNNING THE REFERENCE PIPELINE.

Run in the build container only (needs /root/reference and the staged oracle/_ref binaries). Recipe of SURVEY.md 8(c):
the read-only reference is copied to a scratch directory (it writes a file into its package on import), `dendropy`
(imported, unused on this path) is stubbed by name, HOME is redirected. Then, with the reference's own code:
  1. `subset_alignment_and_hmmbuild` (gcmm/algorithm.py:394-477) builds a 7-subset eHMM directory from
     examples/data/backbone.aln.fasta.gz (the whole backbone, its halves and quarters: tree decomposition itself needs
     a real dendropy and is not on the hot path),
  2. `SearchAlgorithm.search` (gcmm/algorithm.py:273-336) runs the all-against-all hmmsearch jobs for 100 seeded
     fragments of examples/data/unaligned_frag.fasta,
  3. `witch.py -p <eHMM dir> -b ... -q ... --save-weight 1` (examples/run.sh:33-37 with -p) does the rest:
     rankBitscores, writeWeights, getBackbones/hmmalign, the graph DP, the transitivity merge.
Committed under tests/golden/c1/: the 7 profiles, the queries, the upper-cased backbone WITCH worked on, and
c1_golden.json.gz with what the reference produced -- per-subset retained columns / non-gap counts / NSEQ
(readHMMDirectory), the bit-score tables as the reference's own readHMMSearch parses them, weights.txt as its
readWeightsFromLocal parses it, the "passed to main pipeline with top N weights" log lines, the checkpoint rows, the
query rows of aligned.fasta / aligned.masked.fasta and the sha256 of both complete files.

It also cross-checks the on-disk formats of witch_b200/formats.py against the reference's own readers (SURVEY.md
8f-3/8f-4): files written by formats.writeHMMSearchResults / writeWeightsToLocal / writeCheckpointAlignments are parsed
with the reference's readHMMSearch / readWeightsFromLocal / readOneCheckpointAlignment and must give back the content;
the exact file texts are stored so that tests/test_formats_cpu.py can pin the writers byte for byte.

Usage: python tests/golden/make_golden_c1.py
"""
import glob
import gzip
import hashlib
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DATA = "/root/reference/examples/data"
OUT = os.path.join(HERE, "c1")
N_QUERIES = 100


def read_fasta(path):
    op = gzip.open if path.endswith(".gz") else open
    out = []
    with op(path, "rt") as f:
        for ln in f:
            ln = ln.strip()
            if ln.startswith(">"):
                out.append([ln[1:].split()[0], ""])
            elif ln:
                out[-1][1] += ln
    return [(a, b) for a, b in out]


def main():
    sys.path.insert(0, ROOT)
    from oracle.make_ref import ref_tool, build as build_ref
    assert build_ref(), "reference binaries not staged"
    work = tempfile.mkdtemp(prefix="witch_c1_")
    shutil.copytree("/root/reference", os.path.join(work, "ref"))
    stub = os.path.join(work, "stub", "dendropy")
    os.makedirs(os.path.join(stub, "datamodel"))
    open(os.path.join(stub, "__init__.py"), "w").write("class Tree: pass\nclass Taxon: pass\nclass DataSet: pass\nclass treecalc: pass\n")
    open(os.path.join(stub, "datamodel", "__init__.py"), "w").write("")
    open(os.path.join(stub, "datamodel", "treemodel.py"), "w").write("class Tree: pass\n")
    open(os.path.join(stub, "datamodel", "taxonmodel.py"), "w").write("class Taxon: pass\n")
    home = os.path.join(work, "home")
    os.makedirs(home)
    os.environ["HOME"] = home
    sys.path.insert(0, os.path.join(work, "stub"))
    sys.path.insert(0, os.path.join(work, "ref"))
    from multiprocessing import Manager
    from concurrent.futures import ProcessPoolExecutor
    from witch_msa.configs import Configs
    from witch_msa.gcmm.algorithm import SearchAlgorithm, subset_alignment_and_hmmbuild
    from witch_msa.gcmm import loader as ref_loader, weighting as ref_weighting
    for name in ("log", "warning", "runtime", "debug", "error"):
        setattr(Configs, name, staticmethod(lambda *a, **k: None))

    # ---- inputs
    bb = read_fasta(os.path.join(DATA, "backbone.aln.fasta.gz"))
    names = [n for n, _ in bb]
    bb_path = os.path.join(work, "backbone.aln.fasta")
    with open(bb_path, "w") as f:
        for n, s in bb:
            f.write(">%s\n%s\n" % (n, s))
    frags = read_fasta(os.path.join(DATA, "unaligned_frag.fasta"))
    rng = random.Random(3)
    queries = [frags[i] for i in sorted(rng.sample(range(len(frags)), N_QUERIES))]
    q_path = os.path.join(work, "queries.fasta")
    with open(q_path, "w") as f:
        for n, s in queries:
            f.write(">%s\n%s\n" % (n, s))
    subsets = [names, names[:250], names[250:], names[:125], names[125:250], names[250:375], names[375:]]

    # ---- 1. the eHMM directory, with the reference's own function
    out1 = os.path.join(work, "out")
    hmmdir = os.path.join(out1, "tree_decomp", "root")
    os.makedirs(hmmdir)
    m = Manager()
    lock = m.Lock()
    for i, taxa in enumerate(subsets):
        subset_alignment_and_hmmbuild(lock, ref_tool("hmmbuild"), hmmdir, "dna", 0.59, 0.0, "afa", bb_path, ("A_0_%d" % i, set(taxa)))

    # ---- 2. all-against-all searches, with the reference's own class
    Configs.hmmsearchpath = ref_tool("hmmsearch")
    Configs.outdir = out1
    Configs.num_cpus = 8
    Configs.query_path = q_path
    Configs.molecule = "dna"
    Configs.max_concurrent_jobs = 16
    paths = sorted(glob.glob(os.path.join(hmmdir, "A_0_*", "hmmbuild.model.*")), key=lambda p: int(p.rsplit("_", 1)[1]))
    pool = ProcessPoolExecutor(8)
    s = SearchAlgorithm(paths)
    s.molecule = "dna"
    s.search(lock, pool)

    # ---- 3. the rest of the pipeline through the reference's CLI
    out2 = os.path.join(work, "out2")
    env = dict(os.environ, PYTHONPATH=os.path.join(work, "stub"), HOME=home)
    subprocess.check_call([sys.executable, os.path.join(work, "ref", "witch.py"), "-p", hmmdir, "-b", bb_path, "-q", q_path, "-d", out2,
                           "-t", "8", "--save-weight", "1", "--molecule", "dna"], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)

    # ---- what the reference produced, read back with the reference's own readers
    Configs.hmmdir = hmmdir
    _, retained, nongaps = SearchAlgorithm(None).readHMMDirectory(lock, pool)
    index_to_hmm = ref_loader.getAlignmentSubsets(hmmdir, lock, pool)
    gold = dict(n_subsets=len(subsets), backbone_length=len(bb[0][1]),
                subset_taxa=[list(t) for t in subsets],
                nseq={str(i): int(index_to_hmm[i].num_taxa) for i in index_to_hmm},
                retained={str(i): [int(c) for c in retained[i]] for i in retained},
                nongaps={str(i): [int(c) for c in nongaps[i]] for i in nongaps})
    gold["hmmsearch"] = {}
    for i in sorted(index_to_hmm):
        ranks = ref_loader.readHMMSearch(lock, len(index_to_hmm), index_to_hmm[i])
        gold["hmmsearch"][str(i)] = {t: float(v[0][1]) for t, v in ranks.items()}
    w = ref_weighting.readWeightsFromLocal(os.path.join(out2, "weights.txt"))
    gold["weights"] = {t: [(int(i), float(x)) for i, x in sw] for t, sw in w.items()}
    gold["log_top"] = {}
    for ln in open(os.path.join(out2, "log.txt")):   # "<date>\t[LOG] <taxon>\tpassed to main pipeline with top N weights: [(idx, w), ...]"
        if "passed to main pipeline with top" in ln:
            parts = ln.rstrip("\n").split("\t")
            taxon = parts[-2].split("] ")[-1]
            body = parts[-1]
            n = int(body.split("with top ")[1].split()[0])
            import ast
            lst = ast.literal_eval(body.split("weights: ", 1)[1])
            gold["log_top"][taxon] = [n, [(int(i), float(x)) for i, x in lst]]
    with gzip.open(os.path.join(out2, "checkpoint_alignments.txt.gz"), "rb") as f:
        lines = f.read().decode("utf-8").split("\n")[:-1]
    gold["checkpoint_rows"] = {}
    for qa in ref_loader.readOneCheckpointAlignment(lines):
        t = list(qa.keys())[0]
        gold["checkpoint_rows"][t] = qa[t]
    qnames = {n for n, _ in queries}
    for key, fn in (("aligned", "aligned.fasta"), ("masked", "aligned.masked.fasta")):
        rows = read_fasta(os.path.join(out2, fn))
        gold[key] = {n: r for n, r in rows if n in qnames}
        gold[key + "_order"] = [n for n, _ in rows]
        gold[key + "_sha256"] = hashlib.sha256("".join(">%s\n%s\n" % (n, r) for n, r in rows).encode()).hexdigest()
        gold[key + "_width"] = len(rows[0][1])

    # ---- formats cross-check: OUR writers -> the REFERENCE's readers
    sys.path.insert(0, ROOT)
    import numpy as np
    from witch_b200 import formats
    fx = os.path.join(work, "fmt")
    os.makedirs(os.path.join(fx, "A_0_0")); os.makedirs(os.path.join(fx, "A_0_1"))
    for i in (0, 1):
        shutil.copy(paths[i], os.path.join(fx, "A_0_%d" % i, "hmmbuild.model.A_0_%d" % i))
        shutil.copy(os.path.join(hmmdir, "A_0_%d" % i, "hmmbuild.input.A_0_%d.fasta" % i), os.path.join(fx, "A_0_%d" % i))
    ours = formats.getAlignmentSubsets(fx)
    fnames = ["SHFB", "Q two", "x_3"]
    fscores = np.array([[12.34, -3.21], [100.05, 7.0], [0.04, 55.55]], dtype=np.float32)
    frep = np.array([[1, 1], [1, 0], [0, 1]], dtype=bool)
    wpaths = formats.writeHMMSearchResults(ours, fnames, fscores, frep, chunk=0)
    theirs = ref_loader.getAlignmentSubsets(fx, lock, pool)
    fmt = dict(names=fnames, scores=fscores.tolist(), reported=frep.tolist(), hmmsearch_files={}, hmmsearch_parsed={})
    for i in (0, 1):
        assert theirs[i].num_taxa == ours[i].num_taxa and theirs[i].hmm_model_path == ours[i].hmm_model_path
        parsed = ref_loader.readHMMSearch(lock, 2, theirs[i])
        fmt["hmmsearch_parsed"][str(i)] = {t: [list(x) for x in v] for t, v in parsed.items()}
        assert {t: v for t, v in parsed.items()} == formats.readHMMSearch(ours[i]), (parsed, formats.readHMMSearch(ours[i]))
        fmt["hmmsearch_files"][str(i)] = open(wpaths[i]).read()
    t2w = {"SHFB": ((3, 0.9572797479864741), (1, 0.024298524040165876)), "Q_b": ((0, 1.0),), "z": ((5, 0.5), (2, 0.25), (6, 0.25))}
    wp = os.path.join(fx, "weights.txt")
    formats.writeWeightsToLocal(t2w, wp)
    back = ref_weighting.readWeightsFromLocal(wp)
    assert {t: tuple((int(i), float(x)) for i, x in v) for t, v in back.items()} == t2w, back
    assert formats.readWeightsFromLocal(wp) == t2w
    fmt["weights"] = {t: [list(x) for x in v] for t, v in t2w.items()}
    fmt["weights_file"] = open(wp).read()
    rows = {"SHFB": "--ACgtAC-", "Q_b": "acA----gt", "t\tab": "ACGT-----"}
    cp = os.path.join(fx, "checkpoint_alignments.txt.gz")
    formats.writeCheckpointAlignments(cp, {"SHFB": rows["SHFB"]}, append=False)
    formats.writeCheckpointAlignments(cp, {k: rows[k] for k in ("Q_b", "t\tab")}, append=True)
    with gzip.open(cp, "rb") as f:
        lines = f.read().decode("utf-8").split("\n")[:-1]
    back = {}
    for qa in ref_loader.readOneCheckpointAlignment(lines):
        t = list(qa.keys())[0]
        back[t] = qa[t]
    assert back == rows and formats.readCheckpointAlignments(cp) == rows, back
    fmt["checkpoint_rows"] = rows
    fmt["checkpoint_text"] = "\n".join(lines) + "\n"
    gold["formats"] = fmt
    pool.shutdown()

    # ---- commit
    os.makedirs(OUT, exist_ok=True)
    for i, p in enumerate(paths):
        with open(p, "rb") as f, gzip.GzipFile(os.path.join(OUT, "hmm_%d.hmm.gz" % i), "wb", mtime=0) as g:
            g.write(f.read())
    shutil.copy(q_path, os.path.join(OUT, "queries.fasta"))
    tmp_bb = read_fasta(os.path.join(out2, "tree_decomp", "backbone", "backbone.aln.fasta")) if os.path.exists(
        os.path.join(out2, "tree_decomp", "backbone", "backbone.aln.fasta")) else [(n, s.upper()) for n, s in bb]
    with gzip.GzipFile(os.path.join(OUT, "backbone.fasta.gz"), "wb", mtime=0) as g:
        g.write("".join(">%s\n%s\n" % (n, s.upper()) for n, s in tmp_bb).encode())
    with gzip.GzipFile(os.path.join(OUT, "c1_golden.json.gz"), "wb", mtime=0) as g:
        g.write(json.dumps(gold, separators=(",", ":"), sort_keys=True).encode())
    print("c1 golden written:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in gold.items() if k not in ("formats",)})
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
