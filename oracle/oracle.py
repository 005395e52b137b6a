"""CPU oracle for WITCH's eHMM score + align hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product (witch_b200/) never does and has no CPU fallback.

What is restated (reference = /root/reference, WITCH v1.0.10 with bundled HMMER 3.1b2 binaries):
  * HMMER3/f profile parsing + local-mode profile configuration  (SURVEY.md 8a "Score semantics" 1-2;
    HMMER 3.1b2 is a third-party dependency present only as binaries, so its published algorithm is restated
    and pinned by golden vectors produced with those binaries: tests/golden/make_golden.py)
  * hmmsearch --max per-sequence score  -> hmm_oracle.c:orc_score_pair   (call site gcmm/algorithm.py:526-532)
  * the 1-decimal print/parse           -> printed_score()                (gcmm/algorithm.py:579-605)
  * rankBitscores / calculateWeights    -> rank_bitscores(), calculate_weights()  (gcmm/loader.py:299-332,
    gcmm/weighting.py:58-74)
  * adaptive inclusion                  -> adaptive_inclusion()           (gcmm/aligner.py:52-63)
  * hmmalign + Stockholm -> column list -> hmm_oracle.c:orc_align_pair    (gcmm/aligner.py:96-100,126-142)
  * weighted alignment-graph DP         -> graph_align()                  (gcmm/aligner.py:387-495,
    helpers/alignment_tools.py:1356-1384)
Parity status: PINNED against the reference binaries' outputs (tests/golden/*.json, tests/test_oracle_golden.py).
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DNA_SYMS = "ACGT-RYMKSWHBVDN*~"
RNA_SYMS = "ACGU-RYMKSWHBVDN*~"
AMINO_SYMS = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"
_DNA_DEGEN = {"R": "AG", "Y": "CT", "M": "AC", "K": "GT", "S": "CG", "W": "AT", "H": "ACT", "B": "CGT",
              "V": "ACG", "D": "AGT", "N": "ACGT"}
_AMINO_DEGEN = {"B": "ND", "J": "IL", "Z": "QE", "O": "K", "U": "C", "X": "ACDEFGHIKLMNPQRSTVWY"}
AMINO_BG = [0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198, 0.0590092, 0.0594422,
            0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639, 0.0540978, 0.0683364, 0.0540687, 0.0673417,
            0.0114135, 0.0304133]


def build_lib(force=False):
    """Compile hmm_oracle.c -> oracle/liboracle.so (gcc)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, "hmm_oracle.c"), os.path.join(_HERE, "hmm_md.c")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(x) for x in srcs):
        # -ffp-contract=off: hmm_md.c restates HMMER's SSE float arithmetic operation by operation (no FMA contraction)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so] + srcs + ["-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_lib())
        _LIB.orc_forward.restype = ctypes.c_double
        _LIB.orc_backward.restype = ctypes.c_double
        _LIB.orc_graph_dp.restype = ctypes.c_int
        _LIB.md_oprofile_create.restype = ctypes.c_void_p
        _LIB.md_oprofile_free.argtypes = [ctypes.c_void_p]
    return _LIB


class Alphabet:
    def __init__(self, kind):
        kind = kind.lower()
        if kind in ("dna", "rna"):
            self.syms = DNA_SYMS if kind == "dna" else RNA_SYMS
            self.K = 4
            degen = dict(_DNA_DEGEN)
            self.bg = np.full(4, 0.25)
            canon = "ACGT"
        elif kind == "amino":
            self.syms = AMINO_SYMS
            self.K = 20
            degen = dict(_AMINO_DEGEN)
            self.bg = np.array(AMINO_BG)
            canon = AMINO_SYMS[:20]
        else:
            raise ValueError(kind)
        self.kind = kind
        self.Kp = len(self.syms)
        self.code = {c: i for i, c in enumerate(self.syms)}
        if kind == "dna":
            self.code.update({"U": self.code["T"], "X": self.code["N"], "I": self.code["A"]})
        if kind == "rna":
            self.code.update({"T": self.code["U"], "X": self.code["N"], "I": self.code["A"]})
        self.code.update({"_": self.code["-"], ".": self.code["-"]})
        # degeneracy sets as index lists
        self.degen_n = np.zeros(self.Kp, dtype=np.int32)
        self.degen_set = np.zeros((self.Kp, self.K), dtype=np.int32)
        cidx = {c: i for i, c in enumerate(canon)}
        if kind == "rna":
            cidx = {c: i for i, c in enumerate("ACGU")}
            degen = {k: v.replace("T", "U") for k, v in degen.items()}
        for s, members in degen.items():
            x = self.code[s]
            self.degen_n[x] = len(members)
            for a, c in enumerate(members):
                self.degen_set[x, a] = cidx[c]

    def digitize(self, seq):
        return np.array([self.code[c] for c in seq.upper()], dtype=np.uint8)


class Profile:
    """Local-mode Plan-7 profile in probability space (mode/length are applied at DP time)."""

    def __init__(self, path):
        self.path = path
        self._parse(path)
        self._config()

    def _parse(self, path):
        with open(path) as f:
            lines = f.read().split("\n")
        it = iter(lines)
        self.nseq = None
        for ln in it:
            tok = ln.split()
            if not tok:
                continue
            if tok[0] == "LENG":
                self.M = int(tok[1])
            elif tok[0] == "ALPH":
                self.abc = Alphabet(tok[1])
            elif tok[0] == "NSEQ":
                self.nseq = int(tok[1])
            elif tok[0] == "NAME":
                self.name = tok[1]
            elif tok[0] == "HMM":
                break
        next(it)  # transition header line
        K, M = self.abc.K, self.M

        def p(v):
            return 0.0 if v == "*" else math.exp(-float(v))

        def raw(v):
            return math.inf if v == "*" else float(v)

        ln = next(it).split()
        if ln[0] == "COMPO":
            ln = next(it).split()
        # ln = node-0 insert emissions (ignored); next: node-0 transitions
        tok0 = next(it).split()
        t0 = [p(v) for v in tok0]
        self.mat = np.zeros((M + 1, K))
        self.t = np.zeros((M + 1, 7))
        self.t[0] = t0
        # the numbers of the text file as written (-ln p, inf for '*'): input of the float restatement in hmm_md.c
        self.raw_t = np.full((M + 1, 7), np.inf)
        self.raw_mat = np.full((M + 1, K), np.inf)
        self.raw_t[0] = [raw(v) for v in tok0]
        for k in range(1, M + 1):
            tok = next(it).split()
            assert int(tok[0]) == k, (tok, k)
            self.mat[k] = [p(v) for v in tok[1:1 + K]]
            self.raw_mat[k] = [raw(v) for v in tok[1:1 + K]]
            next(it)  # insert emissions: ignored (insert score hard-wired to 0)
            tokt = next(it).split()
            self.t[k] = [p(v) for v in tokt]
            self.raw_t[k] = [raw(v) for v in tokt]

    def _config(self):
        M, K, Kp = self.M, self.abc.K, self.abc.Kp
        t = self.t
        occ = np.zeros(M + 1)
        occ[1] = t[0][1] + t[0][0]
        for k in range(2, M + 1):
            occ[k] = occ[k - 1] * (t[k - 1][0] + t[k - 1][1]) + (1.0 - occ[k - 1]) * t[k - 1][5]
        Z = float(np.sum(occ[1:] * (M - np.arange(1, M + 1) + 1)))
        self.entry = np.zeros(M + 1)
        self.entry[1:] = occ[1:] / Z
        self.tr = t.copy()
        self.tr[0] = 0.0
        self.tr[M] = 0.0
        bg = self.abc.bg
        sc = np.full((Kp, M + 1), -np.inf)
        with np.errstate(divide="ignore"):
            sc[:K, 1:] = np.log(self.mat[1:, :].T / bg[:, None])
        for x in range(K + 1, Kp - 2):
            n = self.abc.degen_n[x]
            if n == 0:
                continue
            idx = self.abc.degen_set[x, :n]
            w = bg[idx]
            sc[x, 1:] = (sc[idx, 1:] * w[:, None]).sum(0) / w.sum()
        self.emis = np.exp(sc)
        self.emis[:, 0] = 0.0
        self.emis = np.ascontiguousarray(self.emis)
        self.tr = np.ascontiguousarray(self.tr)

    def _args(self):
        dp = ctypes.POINTER(ctypes.c_double)
        return (self.tr.ctypes.data_as(dp), self.entry.ctypes.data_as(dp), self.emis.ctypes.data_as(dp))


def _u8(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def forward_nats(prof, dsq, multihit=True, Lmodel=None):
    L = len(dsq)
    return lib().orc_forward(prof.M, prof.abc.Kp, *prof._args(), _u8(dsq), L, int(multihit), Lmodel or L)


def backward_nats(prof, dsq, multihit=True, Lmodel=None):
    L = len(dsq)
    return lib().orc_backward(prof.M, prof.abc.Kp, *prof._args(), _u8(dsq), L, int(multihit), Lmodel or L)


def oprofile(prof):
    """Handle of the float restatement of HMMER's optimized profile (hmm_md.c), cached on the Profile."""
    if getattr(prof, "_om", None) is None:
        ip = ctypes.POINTER(ctypes.c_int)
        dp = ctypes.POINTER(ctypes.c_double)
        prof._bg32 = np.ascontiguousarray(prof.abc.bg, dtype=np.float32)
        prof._raw_t = np.ascontiguousarray(prof.raw_t, dtype=np.float64)
        prof._raw_mat = np.ascontiguousarray(prof.raw_mat, dtype=np.float64)
        prof._om = ctypes.c_void_p(lib().md_oprofile_create(
            prof.M, prof.abc.K, prof.abc.Kp, prof._raw_t.ctypes.data_as(dp), prof._raw_mat.ctypes.data_as(dp),
            prof._bg32.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
            prof.abc.degen_n.ctypes.data_as(ip), prof.abc.degen_set.ctypes.data_as(ip)))
    return prof._om


def score_pair(prof, dsq, multidomain=True):
    """dict(reported, pre_score, score, nregions, flags, env, envelopes, ...) for one (profile, digitized sequence).
    multidomain=True: regions that fail the single-domain test go through HMMER's stochastic-trace clustering
    (hmm_md.c); False: such a region is kept as one envelope (the round-1 simplification, for comparison)."""
    res = np.zeros(13)
    env = np.zeros((64, 5))
    sig = np.zeros((128, 5), dtype=np.int32)
    nsig = ctypes.c_int(0)
    ip = ctypes.POINTER(ctypes.c_int)
    dp = ctypes.POINTER(ctypes.c_double)
    lib().orc_score_pair2(prof.M, prof.abc.Kp, prof.abc.K, *prof._args(),
                          prof.abc.degen_n.ctypes.data_as(ip), prof.abc.degen_set.ctypes.data_as(ip),
                          _u8(dsq), len(dsq), oprofile(prof) if multidomain else None,
                          res.ctypes.data_as(dp), env.ctypes.data_as(dp), 64, sig.ctypes.data_as(ip), 128, ctypes.byref(nsig))
    nenv = int(res[12])
    return dict(reported=bool(res[0]), pre_score=res[1], score=res[2], nregions=int(res[3]), flags=int(res[4]),
                env=(int(res[5]), int(res[6])), fwd=res[7], max_mocc=res[8], mdstat=res[9], seq_score=res[10],
                sum_score=res[11], nenv=nenv, clusters=[tuple(int(x) for x in c) for c in sig[:min(nsig.value, 128)]],
                envelopes=[(int(e[0]), int(e[1]), float(e[2]), float(e[3]), bool(e[4])) for e in env[:min(nenv, 64)]])


def align_pair(prof, dsq):
    """Column list as consumed at gcmm/aligner.py:399-418: len L, match-state index or -1."""
    cols = np.full(len(dsq), -1, dtype=np.int32)
    oasc = ctypes.c_double(0)
    lib().orc_align_pair(prof.M, prof.abc.Kp, *prof._args(), _u8(dsq), len(dsq),
                         cols.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.byref(oasc))
    return cols


def printed_score(score):
    """hmmsearch prints '%6.1f' and WITCH parses that text back (gcmm/algorithm.py:596-599)."""
    return float("%6.1f" % score)


def rank_bitscores(scores):
    """scores: {hmm_idx: printed score} for one query -> list sorted by score desc (stable; gcmm/loader.py:318-330).
    The reference's tie order depends on future completion order; here ties keep ascending hmm_idx."""
    return sorted(sorted(scores.items()), key=lambda x: x[1], reverse=True)


def calculate_weights(indexes, bitscores, sizes, num_hmms):
    """gcmm/weighting.py:58-74 restated, same summation order (numpy f64)."""
    weights = {}
    bs = np.array(bitscores, dtype=np.float64)
    sz = np.array(sizes, dtype=np.float64)
    for i in range(len(bitscores)):
        exponents = bs - bs[i] + np.log2(sz / sz[i])
        weights[indexes[i]] = 1.0 / np.sum(np.power(2, exponents))
    keep = min(num_hmms, len(weights))
    return tuple(sorted(weights.items(), key=lambda x: x[1], reverse=True)[:keep])


def adaptive_inclusion(sorted_weights, target=0.999):
    """gcmm/aligner.py:58-63."""
    cur, idx = 0.0, 0
    while idx < len(sorted_weights) and cur < target:
        cur += sorted_weights[idx][1]
        idx += 1
    return [(w[0], float(w[1])) for w in sorted_weights[:idx]]


def compress_insertions(seq):
    """helpers/alignment_tools.py:1356-1384 restated: in the stretch before the first and after the last
    upper-case (aligned) character, lower-case insertions are packed against the row's ends
    (front: letters then gaps; back: gaps then letters). Rows without any aligned character are unchanged."""
    up = [i for i, c in enumerate(seq) if "A" <= c <= "Z"]
    if not up:
        return seq
    f_end, b_start = up[0], up[-1] + 1
    front = seq[:f_end].replace("-", "")
    back = seq[b_start:].replace("-", "")
    return front + "-" * (f_end - len(front)) + seq[f_end:b_start] + "-" * (len(seq) - b_start - len(back)) + back


def graph_align(seq, backbone_length, subset_to_weight, subset_to_aligned_columns, retained_columns,
                nongaps_per_column):
    """gcmm/aligner.py:387-482 (graph + DP + backtrace + padding), before compressInsertions."""
    ci, cj, cw = [], [], []
    min_col, max_col = backbone_length + 1, -1
    for subset, cols in subset_to_aligned_columns.items():
        for i, c in enumerate(cols):
            if c == -1:
                continue
            j = int(retained_columns[subset][c])
            ci.append(i)
            cj.append(j)
            cw.append(int(nongaps_per_column[subset][c]) * subset_to_weight[subset])
            min_col, max_col = min(min_col, j), max(max_col, j)
    L = len(seq)
    ci = np.array(ci, dtype=np.int32)
    cj = np.array(cj, dtype=np.int32)
    cw = np.array(cw, dtype=np.float64)
    buf = ctypes.create_string_buffer(L + backbone_length + 8)
    i32 = ctypes.POINTER(ctypes.c_int32)
    n = lib().orc_graph_dp(int(L), int(min_col), int(max_col), len(ci), ci.ctypes.data_as(i32), cj.ctypes.data_as(i32),
                           cw.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), buf)
    ops = buf.raw[:n].decode()
    out, i = [], 0
    for o in ops:
        if o == "M":
            out.append(seq[i]); i += 1
        elif o == "I":
            out.append(seq[i].lower()); i += 1
        else:
            out.append("-")
    row = "-" * min_col + "".join(out) + "-" * (backbone_length - max_col - 1)
    return row


def merge_rows(rows, backbone_length):
    """helpers/alignment_tools.py:1183-1316 (ExtendedAlignment.merge_in) + gcmm/merger.py:69-78 restated in closed
    form for the rows WITCH merges: every row has exactly `backbone_length` regular columns (upper case or '-') and
    any number of insertion columns (lower case) between them. Insertion runs that sit in the same backbone gap are
    overlaid from the left, so the merged width of gap g is the longest run any row has there; a row's own run is
    left-aligned in the gap's block and padded with '-'. -> (merged rows, masked rows, gap widths [backbone_length+1])."""
    B = backbone_length
    runs = []
    width = np.zeros(B + 1, dtype=np.int64)
    for r in rows:
        g, cur, rr = 0, [], []
        for ch in r:
            if "a" <= ch <= "z":
                cur.append(ch)
            else:
                rr.append("".join(cur)); cur = []; g += 1
        rr.append("".join(cur))
        assert g == B, (g, B)
        runs.append(rr)
        width = np.maximum(width, [len(x) for x in rr])
    merged, masked = [], []
    for r, rr in zip(rows, runs):
        reg = [ch for ch in r if not ("a" <= ch <= "z")]
        out = []
        for g in range(B):
            out.append(rr[g] + "-" * (int(width[g]) - len(rr[g])))
            out.append(reg[g])
        out.append(rr[B] + "-" * (int(width[B]) - len(rr[B])))
        merged.append("".join(out))
        masked.append("".join(reg))
    return merged, masked, width
