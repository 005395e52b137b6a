"""Thin Python objects over the C ABI: an eHMM resident on the GPU, a packed query set, and the three stage
calls (score, weights/top-k, align). numpy in / numpy out; the *_dev variants take raw device pointers
(e.g. torch tensors' data_ptr()) so scores and weights never leave HBM between stages."""
import ctypes

import numpy as np

from . import _lib
from ._lib import WitchError, check

ALPHABETS = {0: "dna", 1: "rna", 2: "amino"}


def _ptr(a, ctype):
    return a.ctypes.data_as(ctypes.POINTER(ctype))


class EHMM:
    """Ensemble of HMMER3 profiles on the current CUDA device (reference: the hmmbuild.model.* files of
    witch_msa/gcmm/algorithm.py:463-470 and HMMSubset of gcmm/loader.py:17-58)."""

    def __init__(self, hmm_paths, cache=None):
        """cache: optional path of the serialised profile cache (e.g. <tree_decomp>/witch_b200.profiles): when it exists
        and no HMM file changed since it was written the text is not parsed again (`self.cache_hit`)."""
        lib = _lib.load()
        self.paths = [str(p) for p in hmm_paths]
        arr = (ctypes.c_char_p * len(self.paths))(*[p.encode() for p in self.paths])
        h = ctypes.c_void_p()
        self.cache_hit = False
        if cache is None:
            check(lib.witch_ehmm_create(len(self.paths), arr, ctypes.byref(h)))
        else:
            hit = ctypes.c_int(0)
            check(lib.witch_ehmm_create_cached(len(self.paths), arr, str(cache).encode(), ctypes.byref(hit), ctypes.byref(h)))
            self.cache_hit = bool(hit.value)
        self._h = h
        self.n = lib.witch_ehmm_count(h)
        self.M = np.zeros(self.n, dtype=np.int32)
        self.nseq = np.zeros(self.n, dtype=np.int32)
        check(lib.witch_ehmm_info(h, _ptr(self.M, ctypes.c_int32), _ptr(self.nseq, ctypes.c_int32)))
        self.alphabet = ALPHABETS[lib.witch_ehmm_alphabet(h)]

    def close(self):
        if getattr(self, "_h", None):
            try:
                _lib.load().witch_ehmm_destroy(self._h)
            except TypeError:   # interpreter shutdown: module globals are already gone
                pass
            self._h = None

    __del__ = close


class Queries:
    """Query sequences digitised and packed on the device."""

    def __init__(self, ehmm, seqs):
        lib = _lib.load()
        self.ehmm = ehmm
        seqs = [s if isinstance(s, str) else s.decode() for s in seqs]
        self.lengths = np.array([len(s) for s in seqs], dtype=np.int64)
        self.offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum(self.lengths, out=self.offsets[1:])
        blob = "".join(seqs).encode()
        h = ctypes.c_void_p()
        check(lib.witch_queries_create(ehmm._h, len(seqs), blob, _ptr(self.offsets, ctypes.c_int64), ctypes.byref(h)))
        self._h = h
        self.n = len(seqs)

    @classmethod
    def from_buffer(cls, ehmm, residues_ptr, offsets):
        """Queries from a caller-owned host buffer of concatenated ASCII residues (e.g. pinned memory): residues_ptr
        is the integer address, offsets the int64 [n+1] starts."""
        self = cls.__new__(cls)
        self.ehmm = ehmm
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.lengths = np.diff(self.offsets)
        self.n = len(self.lengths)
        h = ctypes.c_void_p()
        check(_lib.load().witch_queries_create(ehmm._h, self.n, ctypes.cast(residues_ptr, ctypes.c_char_p),
                                               _ptr(self.offsets, ctypes.c_int64), ctypes.byref(h)))
        self._h = h
        return self

    def close(self):
        if getattr(self, "_h", None):
            try:
                _lib.load().witch_queries_destroy(self._h)
            except TypeError:   # interpreter shutdown: module globals are already gone
                pass
            self._h = None

    __del__ = close


def score(ehmm, queries):
    """All queries x all HMMs. Returns (scores[n,H] float32 (NaN = unreported), reported[n,H] bool,
    pre[n,H] float32, flags[n,H] uint8)."""
    n, H = queries.n, ehmm.n
    scores = np.empty((n, H), dtype=np.float32)
    rep = np.zeros((n, H), dtype=np.uint8)
    pre = np.empty((n, H), dtype=np.float32)
    flags = np.zeros((n, H), dtype=np.uint8)
    check(_lib.load().witch_score(ehmm._h, queries._h, _ptr(scores, ctypes.c_float), _ptr(rep, ctypes.c_uint8),
                                  _ptr(pre, ctypes.c_float), _ptr(flags, ctypes.c_uint8)))
    return scores, rep.astype(bool), pre, flags


def score_dev(ehmm, queries, d_scores, d_reported, d_pre=0, d_flags=0, stream=0):
    check(_lib.load().witch_score_dev(ehmm._h, queries._h, d_scores, d_reported, d_pre or None, d_flags or None,
                                      stream or None))


def weights_topk(ehmm, scores, reported, k=10, round_decimals=1):
    """-> (idx[n,k] int32 (-1 pad), w[n,k] float64, count[n] int32)."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    rep = np.ascontiguousarray(reported, dtype=np.uint8)
    n = scores.shape[0]
    idx = np.full((n, k), -1, dtype=np.int32)
    w = np.zeros((n, k), dtype=np.float64)
    cnt = np.zeros(n, dtype=np.int32)
    check(_lib.load().witch_weights_topk(ehmm._h, _ptr(scores, ctypes.c_float), _ptr(rep, ctypes.c_uint8), n, k,
                                         round_decimals, _ptr(idx, ctypes.c_int32), _ptr(w, ctypes.c_double),
                                         _ptr(cnt, ctypes.c_int32)))
    return idx, w, cnt


def weights_topk_dev(ehmm, d_scores, d_reported, n, k, round_decimals, d_idx, d_w, d_count, stream=0):
    check(_lib.load().witch_weights_topk_dev(ehmm._h, d_scores, d_reported, n, k, round_decimals, d_idx, d_w, d_count,
                                             stream or None))


def align(ehmm, queries, qidx, hidx):
    """Optimal-accuracy column lists for the given (query, HMM) pairs. -> list of int32 arrays."""
    qidx = np.ascontiguousarray(qidx, dtype=np.int32)
    hidx = np.ascontiguousarray(hidx, dtype=np.int32)
    if len(qidx) != len(hidx) or (len(qidx) and (qidx.min() < 0 or qidx.max() >= queries.n
                                                 or hidx.min() < 0 or hidx.max() >= ehmm.n)):
        raise WitchError("witch_align: pair index out of range")
    lens = queries.lengths[qidx]
    off = np.zeros(len(qidx) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    cols = np.full(int(off[-1]) if len(qidx) else 0, -1, dtype=np.int32)
    if len(qidx):
        check(_lib.load().witch_align(ehmm._h, queries._h, len(qidx), _ptr(qidx, ctypes.c_int32),
                                      _ptr(hidx, ctypes.c_int32), _ptr(off, ctypes.c_int64),
                                      _ptr(cols, ctypes.c_int32)))
    return [cols[off[p]:off[p + 1]] for p in range(len(qidx))]


def graph_align(ehmm, seqs, pair_begin, pair_hmm, pair_w, cols_list, retained_columns, nongaps_per_column, backbone_length):
    """Weighted alignment-graph merge for many queries (gcmm/aligner.py:387-495 + compressInsertions).
    seqs: upper-case query strings; pair_begin[n+1]; pair_hmm/pair_w/cols_list per included pair (decreasing weight
    within a query); retained_columns / nongaps_per_column: per-subset int sequences. -> list of row strings
    ('' where the query has no included HMM)."""
    n = len(seqs)
    qlen = np.array([len(s) for s in seqs], dtype=np.int32)
    res_off = np.zeros(n, dtype=np.int64)
    if n:
        res_off[1:] = np.cumsum(qlen[:-1])
    blob = "".join(seqs).encode()
    pair_begin = np.ascontiguousarray(pair_begin, dtype=np.int32)
    pair_hmm = np.ascontiguousarray(pair_hmm, dtype=np.int32)
    pair_w = np.ascontiguousarray(pair_w, dtype=np.float64)
    col_off = np.zeros(max(len(cols_list), 1), dtype=np.int64)
    if cols_list:
        col_off[1:len(cols_list)] = np.cumsum([len(c) for c in cols_list[:-1]])
        cols = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.int32) for c in cols_list]), dtype=np.int32)
    else:
        cols = np.zeros(1, dtype=np.int32)
    H = len(retained_columns)
    hmm_off = np.zeros(H + 1, dtype=np.int64)
    np.cumsum([len(r) for r in retained_columns], out=hmm_off[1:])
    ret = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.int32) for r in retained_columns]), dtype=np.int32)
    ng = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.int32) for r in nongaps_per_column]), dtype=np.int32)
    cap = 2 * backbone_length + qlen.astype(np.int64) + 2
    row_off = np.zeros(n, dtype=np.int64)
    if n:
        row_off[1:] = np.cumsum(cap[:-1])
    rows = ctypes.create_string_buffer(int(cap.sum()) + 1)
    row_len = np.zeros(n, dtype=np.int32)
    if n:
        check(_lib.load().witch_graph_align(
            ehmm._h, n, _ptr(qlen, ctypes.c_int32), _ptr(res_off, ctypes.c_int64), blob, _ptr(pair_begin, ctypes.c_int32),
            _ptr(pair_hmm, ctypes.c_int32), _ptr(pair_w, ctypes.c_double), _ptr(col_off, ctypes.c_int64),
            _ptr(cols, ctypes.c_int32), H, _ptr(hmm_off, ctypes.c_int64), _ptr(ret, ctypes.c_int32), _ptr(ng, ctypes.c_int32),
            int(backbone_length), _ptr(row_off, ctypes.c_int64), rows, _ptr(row_len, ctypes.c_int32)))
    raw = rows.raw
    return [raw[row_off[q]:row_off[q] + row_len[q]].decode() for q in range(n)]


def debug_fwdbwd(ehmm, queries, qidx, hidx, multihit):
    qidx = np.ascontiguousarray(qidx, dtype=np.int32)
    hidx = np.ascontiguousarray(hidx, dtype=np.int32)
    f = np.zeros(len(qidx), dtype=np.float32)
    b = np.zeros(len(qidx), dtype=np.float32)
    check(_lib.load().witch_debug_fwdbwd(ehmm._h, queries._h, len(qidx), _ptr(qidx, ctypes.c_int32),
                                         _ptr(hidx, ctypes.c_int32), 1 if multihit else 0,
                                         _ptr(f, ctypes.c_float), _ptr(b, ctypes.c_float)))
    return f, b


def kernel_launches():
    return int(_lib.load().witch_kernel_launches())


def merge_rows(ehmm, rows, is_backbone, backbone_length, want_masked=True):
    """Final transitivity merge of aligned rows (gcmm/merger.py:69-78 + helpers/alignment_tools.py:1183-1316, 1140-1156).
    rows: strings with exactly backbone_length regular columns each; is_backbone[r]: the row is a backbone row.
    -> (merged rows, masked rows or None, gap widths)."""
    n = len(rows)
    row_len = np.array([len(r) for r in rows], dtype=np.int32)
    row_off = np.zeros(max(n, 1), dtype=np.int64)
    if n:
        row_off[1:n] = np.cumsum(row_len[:-1].astype(np.int64))
    blob = "".join(rows).encode()
    bb = np.ascontiguousarray(is_backbone, dtype=np.uint8)
    gw = np.zeros(backbone_length + 1, dtype=np.int32)
    ow = ctypes.c_int64(0)
    lib = _lib.load()
    check(lib.witch_merge_rows(ehmm._h, n, _ptr(row_off, ctypes.c_int64), _ptr(row_len, ctypes.c_int32), _ptr(bb, ctypes.c_uint8),
                               blob, int(backbone_length), _ptr(gw, ctypes.c_int32), ctypes.byref(ow), None, 0, None))
    width = int(ow.value)
    merged = ctypes.create_string_buffer(max(n * width, 1))
    masked = ctypes.create_string_buffer(max(n * backbone_length, 1)) if want_masked else None
    check(lib.witch_merge_rows(ehmm._h, n, _ptr(row_off, ctypes.c_int64), _ptr(row_len, ctypes.c_int32), _ptr(bb, ctypes.c_uint8),
                               blob, int(backbone_length), _ptr(gw, ctypes.c_int32), ctypes.byref(ow), merged, width, masked))
    raw = merged.raw
    out = [raw[r * width:(r + 1) * width].decode() for r in range(n)]
    msk = None
    if want_masked:
        rawm = masked.raw
        msk = [rawm[r * backbone_length:(r + 1) * backbone_length].decode() for r in range(n)]
    return out, msk, gw
