/*
 * witch_b200.h -- C ABI of the B200-native eHMM score + align path for WITCH.
 *
 * The reference (c5shen/WITCH v1.0.10) has no FFI for this path: it shells out to HMMER 3.1b2 and talks through
 * text files. Each entry point below replaces one of those process/file boundaries; a maintainer binds them
 * with ctypes. Conventions: every function returns 0 on success and a negative code on
 * error (message via witch_last_error()); the caller owns every host buffer; the library owns device memory
 * behind opaque handles; a handle is bound to the CUDA device that was current when it was created; calls on
 * one handle must be serialised by the caller; no callbacks. There is NO CPU fallback: without a usable CUDA
 * device every compute entry point fails with WITCH_ERR_CUDA.
 */
#ifndef WITCH_B200_H
#define WITCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WITCH_OK 0
#define WITCH_ERR_ARG (-1)    /* bad argument */
#define WITCH_ERR_IO (-2)     /* cannot read / parse an HMM file */
#define WITCH_ERR_CUDA (-3)   /* CUDA runtime error, or no device */
#define WITCH_ERR_LIMIT (-4)  /* model or query exceeds a built-in limit */

#define WITCH_ALPH_DNA 0
#define WITCH_ALPH_RNA 1
#define WITCH_ALPH_AMINO 2

/* flag bits of witch_score's per-pair flags */
#define WITCH_FLAG_MULTIDOMAIN 1 /* a region failed HMMER's single-domain test and was resolved the way hmmsearch does:
                                    200 stochastic tracebacks over the region + clustering into envelopes (informational) */
#define WITCH_FLAG_SUMSCORE 2    /* the reconstruction ("sum") score overrode the per-sequence score */
#define WITCH_FLAG_ENVCAP 4      /* a built-in cap was hit (> 6 regions for the pair, or > 16 envelopes / > 4096 sampled
                                    domains in one multi-domain region); what fits was scored */

typedef struct witch_ehmm witch_ehmm;       /* an ensemble of profile HMMs resident on one GPU */
typedef struct witch_queries witch_queries; /* a digitised, packed query set resident on one GPU */

/* Thread-local description of the last error returned on this thread. */
const char *witch_last_error(void);

/* Version string "witch_b200 <semver> sm_100a". */
const char *witch_version(void);

/* Number of CUDA devices visible (0 when none / no driver). Never fails. */
int witch_device_count(void);

/*
 * Parse n_hmm HMMER3/f ASCII profiles (what `hmmbuild` wrote, reference witch_msa/gcmm/algorithm.py:463-470),
 * configure them in local mode (what hmmsearch / hmmalign do internally after reading the same files,
 * algorithm.py:526-532, aligner.py:98-100) and upload them to the current device.
 * All profiles must share one alphabet. NSEQ (reference gcmm/loader.py:40-58, HMMSubset.num_taxa) is kept.
 */
int witch_ehmm_create(int n_hmm, const char *const *hmm_paths, witch_ehmm **out);
/*
 * Same, through a serialised profile cache kept next to the HMM text (the reference re-reads and
 * re-configures the hmmbuild.model.* files of an existing tree_decomp directory on every `-p` run,
 * witch_msa/gcmm/loader.py:17-58 + the hmmsearch/hmmalign start-up of every job). cache_path: one binary file holding
 * the configured profiles and the size + modification time of every source file; when it is present and every source
 * file is unchanged the text is not parsed (*cache_hit = 1), otherwise the profiles are parsed and the file is
 * (re)written atomically (*cache_hit = 0; an unwritable path is not an error). cache_hit may be NULL.
 */
int witch_ehmm_create_cached(int n_hmm, const char *const *hmm_paths, const char *cache_path, int *cache_hit,
                             witch_ehmm **out);
void witch_ehmm_destroy(witch_ehmm *e);
int witch_ehmm_count(const witch_ehmm *e);
int witch_ehmm_alphabet(const witch_ehmm *e);
/* M[h] = model length (LENG), nseq[h] = NSEQ; either pointer may be NULL. */
int witch_ehmm_info(const witch_ehmm *e, int32_t *M, int32_t *nseq);

/*
 * Digitise and upload n query sequences. residues = all sequences concatenated (ASCII, either case, IUPAC
 * degenerate codes accepted), offsets[n+1] = start of each sequence in residues. Replaces the fragment-chunk
 * FASTA files of gcmm/algorithm.py:338-385 and the per-query c1.fasta of gcmm/aligner.py:356-362.
 */
int witch_queries_create(const witch_ehmm *e, int n, const char *residues, const int64_t *offsets,
                         witch_queries **out);
void witch_queries_destroy(witch_queries *q);
int witch_queries_count(const witch_queries *q);

/*
 * Score stage: every query against every HMM == all (HMM, chunk) `hmmsearch --cpu 1 --noali -E 99999999 --max`
 * jobs of SearchAlgorithm.search (gcmm/algorithm.py:273-336) plus the table parse (algorithm.py:579-605).
 * Outputs are row-major [n_queries][n_hmm] HOST arrays (any may be NULL except scores/reported):
 *   scores    per-sequence bit score, NOT rounded (hmmsearch prints it with one decimal; see witch_weights_topk)
 *   reported  1 if hmmsearch would list the pair (a domain envelope exists), else 0 (score is then NaN)
 *   pre       Forward bit score before the null2 correction (hmmsearch's score + bias)
 *   flags     WITCH_FLAG_* bits
 */
int witch_score(witch_ehmm *e, witch_queries *q, float *scores, uint8_t *reported, float *pre, uint8_t *flags);

/*
 * Same, results left on the device: d_scores/d_reported/(d_pre, d_flags may be NULL) are DEVICE pointers to
 * [n_queries][n_hmm] arrays (e.g. torch tensors' data_ptr()). Work is enqueued on `stream` (a cudaStream_t
 * passed as void*; NULL = default stream). The envelope work list is built on the device; the call synchronises
 * `stream` ONCE to learn its size (a 400-byte read-back after the parser pass) -- and a second and third time only
 * if some region took the multi-domain branch (that branch runs on a side stream of the handle, next to the
 * envelope pass of all other regions; the stream then waits for it). Scratch buffers belong to the handle and are
 * reused between calls: once they have grown to the workload's size a call neither allocates nor frees.
 * The final kernels are left running on `stream` when the call returns.
 * Limits are checked before anything is enqueued (WITCH_ERR_LIMIT): models <= 8192 nodes, and
 * (distinct query symbols) x (model nodes) <= ~50,000 (8192 nodes for plain DNA/RNA, ~2,000 for protein).
 */
int witch_score_dev(witch_ehmm *e, witch_queries *q, float *d_scores, uint8_t *d_reported, float *d_pre,
                    uint8_t *d_flags, void *stream);

/*
 * Adjusted-bitscore weights + top-k == rankBitscores + calculateWeights (gcmm/loader.py:299-332,
 * gcmm/weighting.py:58-74):  w_i = 1 / sum_j 2^((s_j - s_i) + log2(n_j / n_i)) over the reported HMMs j of the
 * query, sorted by weight descending, first k kept. If round_decimals >= 0 the scores are first rounded to that
 * many decimals (1 reproduces the %6.1f text round trip of algorithm.py:596-599); -1 uses them as they are.
 * HOST arrays: idx[n][k] (HMM index, -1 padding), w[n][k] (float64, 0 padding), count[n] (<= k).
 * Ties in weight are broken by ascending HMM index (the reference's order there is nondeterministic).
 */
int witch_weights_topk(const witch_ehmm *e, const float *scores, const uint8_t *reported, int n_queries, int k,
                       int round_decimals, int32_t *idx, double *w, int32_t *count);
/* Device-pointer variant (all pointers are device memory), asynchronous on `stream`. */
int witch_weights_topk_dev(const witch_ehmm *e, const float *d_scores, const uint8_t *d_reported, int n_queries,
                           int k, int round_decimals, int32_t *d_idx, double *d_w, int32_t *d_count, void *stream);

/*
 * Align stage: for each pair p align query qidx[p] to HMM hidx[p] == one `hmmalign -o OUT HMM c1.fasta` per
 * pair plus the Stockholm -> column-list conversion of getBackbones (gcmm/aligner.py:96-142).
 * cols (HOST, int32) receives, at col_offsets[p] .. col_offsets[p] + L(qidx[p]), one value per residue:
 * the 0-based match-state index, or -1 for an inserted / flanking residue.
 */
int witch_align(witch_ehmm *e, witch_queries *q, int n_pairs, const int32_t *qidx, const int32_t *hidx,
                const int64_t *col_offsets, int32_t *cols);
/* Device-output variant: h_qidx, h_hidx, h_col_offsets are HOST arrays (the pair list comes from the caller's
 * host-side inclusion logic), d_cols is device memory. Everything is enqueued on `stream`; the call does not
 * synchronise (except while the handle's scratch buffers are still growing). */
int witch_align_dev(witch_ehmm *e, witch_queries *q, int n_pairs, const int32_t *h_qidx, const int32_t *h_hidx,
                    const int64_t *h_col_offsets, int32_t *d_cols, void *stream);

/*
 * Weighted alignment-graph merge of one query's per-HMM alignments into a backbone row == the graph build, the
 * max-weight-trace DP, its backtrace and compressInsertions of alignSubQueriesNew (gcmm/aligner.py:387-495,
 * helpers/alignment_tools.py:1356-1384), for n_queries queries at once. All arrays are HOST memory.
 *   qlen[n], res_off[n]: length and offset of each query's residues in `residues` (ASCII, already upper-cased)
 *   pair_begin[n+1]: the included pairs of query q are pair_begin[q] .. pair_begin[q+1]-1, in decreasing weight
 *   pair_hmm[np], pair_w[np]: HMM index and weight of each pair; col_off[np], cols: its column list (witch_align)
 *   hmm_off[H+1], retained[], nongaps[]: per subset, the backbone column and the non-gap count of every retained
 *       column (subset_to_retained_columns / subset_to_nongaps_per_column, gcmm/algorithm.py:423-429)
 *   row_off[n]: where query q's row starts in `rows`; capacity needed per query: 2*backbone_length + qlen + 2
 *   rows, row_len[n]: output characters (upper = aligned, lower = insertion, '-' = gap) and their count
 *       (0 = the query has no included HMM; the reference returns an empty alignment for it)
 */
int witch_graph_align(witch_ehmm *e, int n_queries, const int32_t *qlen, const int64_t *res_off, const char *residues,
                      const int32_t *pair_begin, const int32_t *pair_hmm, const double *pair_w, const int64_t *col_off,
                      const int32_t *cols, int n_hmm, const int64_t *hmm_off, const int32_t *retained,
                      const int32_t *nongaps, int backbone_length, const int64_t *row_off, char *rows, int32_t *row_len);

/*
 * Final transitivity merge == ExtendedAlignment.merge_in applied to every query row in turn
 * (helpers/alignment_tools.py:1183-1316 driven by gcmm/merger.py:69-78) + remove_insertion_columns (:1140-1156).
 * Every input row has exactly `backbone_length` regular columns (upper case or '-') and lower-case insertion columns
 * between them (rows from witch_graph_align); rows flagged is_backbone[r] are taken verbatim (no insertion columns).
 * Insertion runs of different rows that sit in the same backbone gap are overlaid from the left. All arrays HOST.
 *   gap_width[backbone_length+1] (out, may be NULL): merged width of every gap's insertion block
 *   *out_width (out): merged alignment width = backbone_length + sum(gap_width)
 *   merged (may be NULL): nrows rows of merged_cap_width bytes each (>= *out_width; only the first *out_width are
 *       written); masked (may be NULL): nrows rows of backbone_length bytes (insertion columns removed).
 * Call once with merged = masked = NULL to learn *out_width, then again with buffers.
 */
int witch_merge_rows(witch_ehmm *e, int nrows, const int64_t *row_off, const int32_t *row_len, const uint8_t *is_backbone,
                     const char *rows, int backbone_length, int32_t *gap_width, int64_t *out_width, char *merged,
                     int64_t merged_cap_width, char *masked);

/*
 * Instrumentation for bench.py: number of kernels this library has launched so far in this process and the
 * accumulated device time (ms, CUDA events on the launching stream) of the dominant DP kernels since the last
 * witch_prof_reset(). Timing is only collected while witch_prof_enable(1).
 */
uint64_t witch_kernel_launches(void);
void witch_prof_enable(int on);
void witch_prof_reset(void);
/* which: 0 = multihit Forward/Backward parser kernel, 1 = envelope (unihit) kernels, 2 = align kernels,
 * 3 = multi-domain branch (runs concurrently with 1; cells not counted).
 * Returns accumulated ms and (via *cells) the DP cells those launches covered. Timed launches leave CUDA events
 * behind that are resolved here (this call synchronises on them; the stage calls do not). */
double witch_prof_get(int which, double *cells, uint64_t *launches);

/* Measured FP32 FMA throughput of the current device in TFLOP/s (2 flop per FMA lane-op): a dependent-chain-free
 * FFMA micro-benchmark run for roughly `ms` milliseconds. This is the ALU roofline denominator bench.py reports
 * (MEASURED_PEAKS.json carries no CUDA-core figure). Returns a negative value on error. */
double witch_measure_fp32_peak(double ms);

/* Debug / test hooks (stable, used by tests/): plain Forward and Backward scores in nats for n_pairs pairs.
 * mode: 1 = multihit local (hmmsearch parser), 0 = unihit local (hmmalign / envelope). HOST arrays. */
int witch_debug_fwdbwd(witch_ehmm *e, witch_queries *q, int n_pairs, const int32_t *qidx, const int32_t *hidx,
                       int mode, float *fwd_nats, float *bwd_nats);

#ifdef __cplusplus
}
#endif
#endif /* WITCH_B200_H */
