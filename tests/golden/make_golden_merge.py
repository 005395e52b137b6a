"""Golden vectors for the next row: the final transitivity merge, produced by IMPORTING THE REFERENCE'S
OWN PYTHON (witch_msa.helpers.alignment_tools.ExtendedAlignment.merge_in / remove_insertion_columns, driven exactly
like gcmm/merger.py:42-102 mergeAlignmentsCollapsed).

Inputs: the query rows of tests/golden/dna_small/graph_golden.json.gz (the reference's alignSubQueriesNew output,
labelled as gcmm/aligner.py:489-495 does) and the first 6 rows of the example backbone alignment.
Output: tests/golden/dna_small/merge_golden.json.gz = {backbone: [(name,row)], order: [...], merged: {name: row},
masked: {name: row}}. Run in the build container only (copies the read-only reference to a scratch directory and
stubs `dendropy`, which this path imports but never uses).
"""
import gzip
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    from make_golden import read_fasta
    work = tempfile.mkdtemp(prefix="witch_ref_")
    shutil.copytree("/root/reference/witch_msa", os.path.join(work, "witch_msa"))
    stub = os.path.join(work, "dendropy")
    os.makedirs(os.path.join(stub, "datamodel"))
    open(os.path.join(stub, "__init__.py"), "w").write("class Tree: pass\nclass Taxon: pass\nclass DataSet: pass\nclass treecalc: pass\n")
    open(os.path.join(stub, "datamodel", "__init__.py"), "w").write("")
    open(os.path.join(stub, "datamodel", "treemodel.py"), "w").write("class Tree: pass\n")
    open(os.path.join(stub, "datamodel", "taxonmodel.py"), "w").write("class Taxon: pass\n")
    os.environ["HOME"] = os.path.join(work, "home")
    os.makedirs(os.environ["HOME"])
    sys.path.insert(0, work)
    from witch_msa.helpers.alignment_tools import ExtendedAlignment

    G = json.loads(gzip.open(os.path.join(HERE, "dna_small", "graph_golden.json.gz")).read())
    bb = read_fasta("/root/reference/examples/data/backbone.aln.fasta.gz")[:6]
    assert all(len(s) == G["backbone_length"] for _, s in bb)
    bbpath = os.path.join(work, "bb.fasta")
    with open(bbpath, "w") as f:
        for n, s in bb:
            f.write(">%s\n%s\n" % (n, s))
    # queries: ExtendedAlignment objects labelled like gcmm/aligner.py:486-495
    queries, order = [], []
    for taxon, q in G["queries"].items():
        row = q["row"]
        if row is None:
            continue
        ea = ExtendedAlignment([])
        ea[taxon] = row
        ea._reset_col_names()
        insertion, regular = -1, 0
        for i in range(len(row)):
            if row[i].islower():
                ea._col_labels[i] = insertion; insertion -= 1
            else:
                ea._col_labels[i] = regular; regular += 1
        queries.append(ea); order.append(taxon)
    # gcmm/merger.py:69-78
    full_aln = ExtendedAlignment([])
    full_aln.read_file_object(bbpath)
    full_aln.from_string_to_bytearray()
    for query in queries:
        full_aln.merge_in(query, False)
    full_aln.from_bytearray_to_string()
    merged = {k: str(v) for k, v in full_aln.items()}
    full_aln.remove_insertion_columns()                      # gcmm/merger.py:97
    masked = {k: str(v) for k, v in full_aln.items()}
    out = dict(backbone=bb, order=order, merged=merged, masked=masked)
    with gzip.GzipFile(os.path.join(HERE, "dna_small", "merge_golden.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    w = len(next(iter(merged.values())))
    print("merged %d rows, width %d (backbone %d + %d insertion columns), masked width %d" % (
        len(merged), w, G["backbone_length"], w - G["backbone_length"], len(next(iter(masked.values())))))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
