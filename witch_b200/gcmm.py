"""Host-side mirror of the reference's interface for the eHMM score + align path.

The function names, argument meaning and return types follow witch_msa/gcmm of the reference (WITCH v1.0.10) so
that its call sites (gcmm/gcmm.py:160-161, 219-222, 241-244) can switch with a few lines:

    reference                                               here
    SearchAlgorithm.search          (algorithm.py:273-336)  BatchedSearch.search      -> one batched device call
    evalHMMSearchOutput             (algorithm.py:579-605)  (no text to parse; the 1-decimal rounding is applied
                                                             where the reference parsed the text)
    rankBitscores                   (loader.py:299-376)     BatchedSearch.rankBitscores
    calculateWeights/writeWeights   (weighting.py:58-169)   BatchedSearch.writeWeights
    getBackbones(use_gcm=False)     (aligner.py:33-148)     BatchedSearch.getBackbones (all queries at once)
    writeWeightsToLocal             (weighting.py:174-178)  writeWeightsToLocal

All arithmetic runs in libwitch_b200.so on the GPU; this module only shapes inputs/outputs. torch is used to hold
the score / weight tensors on the device between the stages (plumbing). No CPU fallback.
"""
import math
import time

import numpy as np

from . import api


def adaptive_inclusion_counts(w, count, target=0.999):
    """Vectorised form of gcmm/aligner.py:58-63: per query, number of leading weights to keep so that their sum
    reaches `target` (at least one when count > 0). Sequential float64 accumulation like the reference."""
    n, k = w.shape
    keep = np.zeros(n, dtype=np.int32)
    cur = np.zeros(n, dtype=np.float64)
    active = count > 0
    for j in range(k):
        take = active & (j < count) & (cur < target)
        cur = np.where(take, cur + w[:, j], cur)
        keep += take.astype(np.int32)
    return keep


class BatchedSearch:
    """All-against-all query-HMM scoring, weighting and alignment on one GPU.

    index_to_hmm order == order of `hmm_paths`; taxon names are the (already renamed) query names of the reference's
    loadSubQueries (gcmm/loader.py:381-405)."""

    def __init__(self, hmm_paths, num_hmms=10, use_weight=True, runtime_path=None, profile_cache=None):
        # profile_cache: e.g. <outdir>/tree_decomp/witch_b200.profiles -- a `-p` re-run over an existing decomposition
        # (gcmm/gcmm.py:141-152) then skips parsing and configuring the hmmbuild.model.* text again
        t0 = time.time()
        self.ehmm = api.EHMM(hmm_paths, cache=profile_cache)
        self.num_hmms = int(num_hmms)
        self.use_weight = use_weight
        self.queries = None
        self.taxa = None
        self._dev = {}
        # <outdir>/runtime_breakdown.txt of the reference (configs.py:112-116, written through Configs.runtime): one
        # "(tag) Time to ... (s): <seconds>" line per stage, appended here in the same format when a path is given
        self.runtime_path = runtime_path
        self._runtime("gpu_load", "load the eHMM onto the GPU (profile cache %s)" % (
            "hit" if self.ehmm.cache_hit else "miss" if profile_cache else "off"), time.time() - t0)

    def _runtime(self, tag, what, seconds):
        if self.runtime_path:
            with open(self.runtime_path, "a") as f:
                f.write("({}) Time to {} (s): {}\n".format(tag, what, seconds))

    # ------------------------------------------------------------------ score
    def search(self, taxa, seqs):
        """SearchAlgorithm.search: returns nothing in the reference (results go to files); here the device score
        table is kept on `self` and also returned as numpy: (scores[n,H], reported[n,H])."""
        t0 = time.time()
        self.taxa = list(taxa)
        self._seqs = list(seqs)
        self.queries = api.Queries(self.ehmm, seqs)
        scores, rep, pre, flags = api.score(self.ehmm, self.queries)
        self.scores, self.reported, self.pre, self.flags = scores, rep, pre, flags
        self._runtime("gpu_score", "run all-against-all HMM scoring on the GPU", time.time() - t0)
        return scores, rep

    def hmmsearch_results(self, h, evalue=0.0):
        """The dict the reference writes per (HMM, chunk) file (algorithm.py:535-537): {taxon: (evalue, score)} with
        the 1-decimal printed score. E-values are not used downstream (loader.py:293) and are reported as `evalue`."""
        out = {}
        for q, t in enumerate(self.taxa):
            if self.reported[q, h]:
                out[t] = (evalue, float("%.1f" % self.scores[q, h]))
        return out

    def rankBitscores(self):
        """-> ranked_bitscores: {taxon: [(hmm_idx, score), ...]} sorted by score descending (loader.py:318-330)."""
        ranked = {}
        for q, t in enumerate(self.taxa):
            idx = np.nonzero(self.reported[q])[0]
            sc = [float("%.1f" % self.scores[q, h]) for h in idx]
            ranked[t] = sorted(zip((int(h) for h in idx), sc), key=lambda x: x[1], reverse=True)
        return ranked

    def writeWeights(self):
        """-> taxon_to_weights: {taxon: ((hmm_idx, weight), ...)} top num_hmms by weight (weighting.py:121-169).
        Queries without any reported HMM are absent, as in the reference."""
        t0 = time.time()
        idx, w, cnt = api.weights_topk(self.ehmm, self.scores, self.reported, self.num_hmms, 1)
        self.top_idx, self.top_w, self.top_count = idx, w, cnt
        self._runtime("gpu_weights", "obtain weights given bitscores", time.time() - t0)
        out = {}
        for q, t in enumerate(self.taxa):
            if cnt[q] > 0:
                out[t] = tuple((int(idx[q, j]), float(w[q, j])) for j in range(cnt[q]))
        return out

    # ------------------------------------------------------------------ align
    def getBackbones(self, taxon_to_weights=None):
        """Batched getBackbones(use_gcm=False) for every query: -> {taxon: (log_str, weights_map,
        subset_to_aligned_columns)} exactly as aligner.py:33-148 returns per query."""
        if taxon_to_weights is None:
            taxon_to_weights = self.writeWeights()
        name_to_q = {t: q for q, t in enumerate(self.taxa)}
        pq, ph, owner = [], [], []
        included = {}
        for t, sw in taxon_to_weights.items():
            if len(sw) == 0:
                continue
            if self.use_weight:
                cur, i = 0.0, 0
                while i < len(sw) and cur < 0.999:
                    cur += sw[i][1]
                    i += 1
                top = [(x[0], float(x[1])) for x in sw[:i]]
            else:
                top = [(x[0], 1) for x in sw]
            included[t] = top
            for h, _ in top:
                pq.append(name_to_q[t]); ph.append(h); owner.append(t)
        t0 = time.time()
        cols = api.align(self.ehmm, self.queries, pq, ph)
        self._runtime("gpu_align", "align queries to their weighted HMMs", time.time() - t0)
        out = {}
        for t, sw in taxon_to_weights.items():
            if len(sw) == 0:
                out[t] = ("N/A", None)
                continue
            top = included[t]
            log = "{}\tpassed to main pipeline with top {} weights: {}".format(t, len(top), top)
            out[t] = (log, {i: w for (i, w) in sw}, {})
        for t, h, c in zip(owner, ph, cols):
            out[t][2][h] = [int(x) for x in c]
        return out


    def alignSubQueriesNew(self, backbone_length, subset_to_retained_columns, subset_to_nongaps_per_column,
                           taxon_to_weights=None):
        """Batched alignSubQueriesNew (gcmm/aligner.py:350-538) for every query: adaptive inclusion, device alignment
        of the kept (query, HMM) pairs, device weighted alignment-graph DP + backtrace + compressInsertions.
        -> {taxon: row string} (upper = aligned to a backbone column, lower = insertion, '-' = gap); queries
        without weights are absent (the reference returns an empty ExtendedAlignment for them)."""
        bb = self.getBackbones(taxon_to_weights)
        seqs, pair_begin, pair_hmm, pair_w, cols, taxa = [], [0], [], [], [], []
        name_to_q = {t: q for q, t in enumerate(self.taxa)}
        for t, v in bb.items():
            if v[0] == "N/A":
                continue
            _, wmap, s2c = v
            taxa.append(t)
            seqs.append(self._seq_upper(name_to_q[t]))
            for h, c in s2c.items():        # insertion order == decreasing weight (getBackbones)
                pair_hmm.append(h); pair_w.append(wmap[h]); cols.append(c)
            pair_begin.append(len(pair_hmm))
        H = self.ehmm.n
        ret = [subset_to_retained_columns[h] for h in range(H)]
        ng = [subset_to_nongaps_per_column[h] for h in range(H)]
        t0 = time.time()
        rows = api.graph_align(self.ehmm, seqs, pair_begin, pair_hmm, pair_w, cols, ret, ng, backbone_length)
        self._runtime("gpu_graph", "merge per-HMM alignments of every query (graph DP)", time.time() - t0)
        return {t: r for t, r in zip(taxa, rows) if r}

    def _seq_upper(self, q):
        return self._seqs[q].upper()


def mergeAlignmentsCollapsed(ehmm, backbone_items, query_rows, backbone_length, outpath=None, renamed_taxa=None):
    """gcmm/merger.py:42-102 mirror: merge every query row (the dict alignSubQueriesNew returns) into the backbone
    alignment with transitivity, singleton insertions collapsed in lower case (UPP style); the masked version has
    the insertion columns removed. backbone_items: [(name, aligned row)]. Names renamed by the reference's loader
    (renamed_taxa: original -> renamed) are mapped back. When `outpath` is given both FASTA files are written with
    the reference's naming rule (<name>.masked.<suffix>). -> (merged dict, masked dict), backbone rows first, then the
    queries in insertion order -- the order merge_in's dict update produces."""
    names = [n for n, _ in backbone_items] + list(query_rows.keys())
    rows = [r for _, r in backbone_items] + list(query_rows.values())
    flags = [1] * len(backbone_items) + [0] * len(query_rows)
    merged, masked, _ = api.merge_rows(ehmm, rows, flags, backbone_length)
    back = {v: k for k, v in (renamed_taxa or {}).items()}
    names = [back.get(n, n) for n in names]
    full, mask = dict(zip(names, merged)), dict(zip(names, masked))
    if outpath:
        suffix = outpath.split(".")[-1]
        masked_outpath = ".".join(outpath.split(".")[:-1]) + ".masked." + suffix if suffix in ("fa", "fasta") else outpath + ".masked.fasta"
        for path, d in ((outpath, full), (masked_outpath, mask)):
            with open(path, "w") as f:
                for n, r in d.items():
                    f.write(">{}\n{}\n".format(n, r))
    return full, mask


def writeWeightsToLocal(taxon_to_weights, path):
    """weighting.py:174-178; plain floats (not numpy reprs) so that readWeightsFromLocal's eval works everywhere."""
    with open(path, "w") as f:
        for taxon, weights in taxon_to_weights.items():
            f.write("{}:{}\n".format(taxon, tuple((int(i), float(w)) for i, w in weights)))


# ---------------------------------------------------------------------------------------------------------------
class DevicePipeline:
    """The whole hot path with every intermediate kept in HBM (torch tensors): score -> weights/top-k -> adaptive
    inclusion -> align. Used by bench.py through the multi-GPU driver witch_b200.sharding.run_sharded."""

    def __init__(self, ehmm, k=10, device=None):
        import torch
        self.torch = torch
        self.ehmm = ehmm
        self.k = k
        self.device = device or torch.device("cuda", torch.cuda.current_device())

    def run(self, queries, stream=None):
        torch = self.torch
        n, H, k = queries.n, self.ehmm.n, self.k
        st = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        dev = self.device
        scores = torch.empty((n, H), dtype=torch.float32, device=dev)
        rep = torch.empty((n, H), dtype=torch.uint8, device=dev)
        api.score_dev(self.ehmm, queries, scores.data_ptr(), rep.data_ptr(), 0, 0, st)
        idx = torch.empty((n, k), dtype=torch.int32, device=dev)
        w = torch.empty((n, k), dtype=torch.float64, device=dev)
        cnt = torch.empty((n,), dtype=torch.int32, device=dev)
        api.weights_topk_dev(self.ehmm, scores.data_ptr(), rep.data_ptr(), n, k, 1, idx.data_ptr(), w.data_ptr(),
                             cnt.data_ptr(), st)
        # adaptive inclusion needs the (tiny) top-k table on the host to size the align batch
        idx_h, w_h, cnt_h = idx.cpu().numpy(), w.cpu().numpy(), cnt.cpu().numpy()
        keep = adaptive_inclusion_counts(w_h, cnt_h)
        qsel = np.repeat(np.arange(n, dtype=np.int32), keep)
        jsel = np.concatenate([np.arange(c, dtype=np.int32) for c in keep]) if n else np.zeros(0, np.int32)
        hsel = idx_h[qsel, jsel].astype(np.int32) if len(qsel) else np.zeros(0, np.int32)
        lens = queries.lengths[qsel] if len(qsel) else np.zeros(0, np.int64)
        off = np.zeros(len(qsel) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        cols = torch.full((max(int(off[-1]), 1),), -1, dtype=torch.int32, device=dev)
        if len(qsel):
            import ctypes
            from . import _lib
            _lib.check(_lib.load().witch_align_dev(
                self.ehmm._h, queries._h, len(qsel), qsel.ctypes.data_as(_lib.c_i32p), hsel.ctypes.data_as(_lib.c_i32p),
                off.ctypes.data_as(_lib.c_i64p), ctypes.c_void_p(cols.data_ptr()), ctypes.c_void_p(st)))
        cells_score = float(queries.lengths.sum()) * float(self.ehmm.M.sum())
        cells_align = float((lens * self.ehmm.M[hsel]).sum()) if len(qsel) else 0.0
        return dict(scores=scores, reported=rep, idx=idx, w=w, count=cnt, pair_q=qsel, pair_h=hsel, col_off=off,
                    cols=cols, keep=keep, cells_score=cells_score, cells_align=cells_align)
