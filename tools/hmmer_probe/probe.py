"""TEST TOOLING (build container only): runs the reference's hmmsearch binary on one profile and a few queries with WITCH's
command line and captures, besides the printed per-sequence / per-domain numbers (--domE raised so that every domain is
listed), the cluster list of p7_spensemble_Cluster for every region that went through the stochastic-trace branch: an
LD_PRELOAD interposer on qsort() dumps the 24-byte {idx,i,j,k,m,prob} records the binary sorts by start coordinate.
Used by tests/golden/make_golden_md.py to pin oracle/hmm_md.c."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build_interposer():
    so = os.path.join(tempfile.gettempdir(), "witch_qsort_trace.so")
    src = os.path.join(HERE, "qsort_trace.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", so, src, "-ldl"])
    return so


def hmmsearch_probe(hmm_path, queries, hmmsearch=None):
    """queries: [(name, seq)] -> {name: dict(score, bias, domains=[(score, bias, envfrom, envto)], clusters=[[(i,j,k,m,count)]])}.
    `clusters` has one list per multi-domain region of that query, in sequence order (before the dominated-domain filter)."""
    hmmsearch = hmmsearch or os.path.join(ROOT, "oracle", "_ref", "hmmer", "hmmsearch")
    so = build_interposer()
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name, seq in queries:   # one query per process: the qsort trace is then unambiguous
            fa, tbl, dom, tr = (os.path.join(td, x) for x in ("q.fa", "tbl", "dom", "trace"))
            with open(fa, "w") as f:
                f.write(">%s\n%s\n" % (name, seq))
            env = dict(os.environ, LD_PRELOAD=so, QSORT_TRACE=tr)
            subprocess.check_call([hmmsearch, "--cpu", "1", "--noali", "-E", "99999999", "--domE", "99999999", "-o", os.devnull,
                                   "--max", "--tblout", tbl, "--domtblout", dom, hmm_path, fa], env=env)
            rec = dict(score=None, bias=None, domains=[], clusters=[])
            for ln in open(tbl):
                if not ln.startswith("#"):
                    t = ln.split(); rec["score"] = float(t[5]); rec["bias"] = float(t[6])
            for ln in open(dom):
                if not ln.startswith("#"):
                    t = ln.split(); rec["domains"].append((float(t[13]), float(t[14]), int(t[19]), int(t[20])))
            if os.path.exists(tr):
                for ln in open(tr):
                    t = ln.split()
                    if t and t[0] == "SIGC":
                        cl = []
                        for e in t[2:]:
                            v = e.split(",")
                            cl.append((int(v[1]), int(v[2]), int(v[3]), int(v[4]), int(round(float(v[5]) * 200))))
                        rec["clusters"].append(sorted(cl))
                os.remove(tr)
            out[name] = rec
    return out
