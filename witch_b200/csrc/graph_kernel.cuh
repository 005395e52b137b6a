// "Next" row: WITCH's weighted alignment-graph DP + backtrace + compressInsertions on the device
// (reference witch_msa/gcmm/aligner.py:387-495 and helpers/alignment_tools.py:1356-1384).
//
// One warp per query. The graph is sparse (<= one edge per residue and included HMM) but the reference resolves
// ties through a dense backtrace matrix, so the DP is done densely over (L+1) x (max_col-min_col+2) in FLOAT64 with
// exactly the reference's operation order: lane l owns one row of a 32-row block and sweeps the columns with a
// one-step skew (row above comes from lane l-1's previous step), 2 backtrace bits per cell are packed 16 to a word
// and stored coalesced; lane 0 then walks the trace and writes the row.
#pragma once
#include <cstdint>

namespace witch {

constexpr int GRAPH_KMAX = 16;  // included HMMs per query the kernel supports

struct GraphWork {
    int nq;
    const int *qlen;               // [nq]
    const long long *res_off;      // [nq] offset of the query's residues (ASCII) in `residues`
    const char *residues;
    const int *pair_begin;         // [nq+1] included pairs of query q, in decreasing weight order
    const int *pair_hmm;           // [np]
    const double *pair_w;          // [np]
    const long long *col_off;      // [np] offset of the pair's column list
    const int *cols;               // concatenated column lists (-1 = unaligned residue)
    const long long *hmm_off;      // [H+1] offsets into retained / nongaps
    const int *retained;           // backbone column of every retained column of every subset
    const int *nongaps;            // number of non-gap characters of that column in the subset
    int backbone_length;
    const long long *row_off;      // [nq] start of the query's output row (capacity 2*BL + L + 2)
    char *rows;
    int *row_len;                  // [nq] 0 = query has no included HMM (ignored by the reference)
    unsigned *counter;
    char *scratch;                 // per warp slot
    long long slot_bytes;
    int Lcap;
};

__global__ void __launch_bounds__(128) graph_dp_kernel(GraphWork G) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    char *slot = G.scratch + ((long long)blockIdx.x * 4 + w) * G.slot_bytes;
    const int BL = G.backbone_length;
    const int Wcap = BL + 3;
    // slot layout: rowbuf[Wcap] doubles | ent_j[Lcap][K] ints | ent_w[Lcap][K] doubles | ent_n[Lcap] | trace[Lcap+Wcap] | bt words
    double *rowbuf = (double *)slot;
    double *ent_w = rowbuf + Wcap;
    int *ent_j = (int *)(ent_w + (size_t)G.Lcap * GRAPH_KMAX);
    int *ent_n = ent_j + (size_t)G.Lcap * GRAPH_KMAX;
    char *trace = (char *)(ent_n + G.Lcap);
    unsigned *btw = (unsigned *)(trace + ((G.Lcap + Wcap + 15) / 16) * 16);
    for (;;) {
        int q = 0;
        if (lane == 0) q = (int)atomicAdd(G.counter, 1u);
        q = __shfl_sync(FULL, q, 0);
        if (q >= G.nq) break;
        const int L = G.qlen[q];
        const int p0 = G.pair_begin[q], np = min(G.pair_begin[q + 1] - p0, GRAPH_KMAX);
        char *out = G.rows + G.row_off[q];
        const char *res = G.residues + G.res_off[q];
        if (np <= 0 || L <= 0) {
            if (lane == 0) G.row_len[q] = 0;
            continue;
        }
        // ---- 1. sparse edge lists per residue, sorted by backbone column, duplicates summed in subset order ----
        int mn = BL + 1, mx = -1;
        for (int i = lane; i < L; i += 32) {
            int ej[GRAPH_KMAX];
            double ew[GRAPH_KMAX];
            int n = 0;
            for (int p = 0; p < np; p++) {
                const int c = G.cols[G.col_off[p0 + p] + i];
                if (c < 0) continue;
                const int h = G.pair_hmm[p0 + p];
                const long long o = G.hmm_off[h] + c;
                const int j = G.retained[o];
                const double wt = (double)G.nongaps[o] * G.pair_w[p0 + p];
                int z = 0;
                while (z < n && ej[z] != j) z++;
                if (z < n) ew[z] += wt;      // same (i, j): accumulate in subset order, like the reference's dict
                else { ej[n] = j; ew[n] = wt; n++; }
                mn = min(mn, j); mx = max(mx, j);
            }
            // insertion sort by column
            for (int a = 1; a < n; a++) {
                const int kj = ej[a]; const double kw = ew[a];
                int b = a - 1;
                while (b >= 0 && ej[b] > kj) { ej[b + 1] = ej[b]; ew[b + 1] = ew[b]; b--; }
                ej[b + 1] = kj; ew[b + 1] = kw;
            }
            ent_n[i] = n;
            for (int a = 0; a < n; a++) { ent_j[(size_t)i * GRAPH_KMAX + a] = ej[a]; ent_w[(size_t)i * GRAPH_KMAX + a] = ew[a]; }
        }
        for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(FULL, mn, o)); mx = max(mx, __shfl_xor_sync(FULL, mx, o)); }
        __syncwarp();
        const int min_col = mn, max_col = mx;
        const int W = max_col + 2 - min_col;  // columns min_col .. max_col+1  (<= 0 when nothing is aligned)
        int ntrace = 0;
        if (W >= 2) {
            // ---- 2. dense DP, 32 rows per block, skewed ----
            const int nsteps = (W - 1) + 31;
            const int NW16 = (nsteps >> 4) + 1;
            for (int z = lane; z < W; z += 32) rowbuf[z] = 0.0;
            __syncwarp();
            const int nblocks = (L + 31) >> 5;
            for (int b = 0; b < nblocks; b++) {
                const int i = (b << 5) + lane + 1;  // DP row (1-based), residue index i-1
                const bool rowok = i <= L;
                int ne = 0, pe = 0, nj = -1;
                double nw = 0.0;
                const int *ejp = ent_j + (size_t)(i - 1) * GRAPH_KMAX;
                const double *ewp = ent_w + (size_t)(i - 1) * GRAPH_KMAX;
                if (rowok) {
                    ne = ent_n[i - 1];
                    if (ne > 0) { nj = ejp[0] - min_col + 1; nw = ewp[0]; }  // DP column jj that uses edge (i-1, j): jj = j - min_col + 1
                }
                double left = 0.0, prev_up = 0.0, cur = 0.0;
                unsigned word = 0;
                unsigned *bw = btw + (size_t)b * NW16 * 32 + lane;
                for (int t = 1; t <= nsteps; t++) {
                    const int jj = t - lane;
                    double up = __shfl_up_sync(FULL, cur, 1);
                    if (lane == 0) up = (jj >= 1 && jj < W) ? rowbuf[jj] : 0.0;
                    const bool act = rowok && jj >= 1 && jj < W;
                    if (act) {
                        double cw = 0.0;
                        if (pe < ne && nj == jj) {
                            cw = nw;
                            pe++;
                            if (pe < ne) { nj = ejp[pe] - min_col + 1; nw = ewp[pe]; }
                        }
                        const double v0 = prev_up + cw;
                        double c = 0.0;
                        unsigned bt = 0;
                        if (cw <= 0.0) bt = 1; else if (v0 > c) { c = v0; bt = 0; }
                        if (up > c) { c = up; bt = 1; }
                        if (left > c) { c = left; bt = 2; }
                        cur = c; left = c; prev_up = up;
                        word |= bt << (2 * (t & 15));
                        if (lane == 31) rowbuf[jj] = c;  // last row of the block feeds lane 0 of the next one
                    }
                    if ((t & 15) == 15 || t == nsteps) { bw[(size_t)(t >> 4) * 32] = word; word = 0; }
                }
                __syncwarp();
            }
            // ---- 3. backtrace (lane 0) ----
            if (lane == 0) {
                int i = L, jj = W - 1;
                while (i > 0 && jj > 0) {
                    const int b = (i - 1) >> 5, l = (i - 1) & 31, t = jj + l;
                    const unsigned wd = btw[((size_t)b * NW16 + (t >> 4)) * 32 + l];
                    const unsigned bt = (wd >> (2 * (t & 15))) & 3u;
                    if (bt == 0) { trace[ntrace++] = res[i - 1]; i--; jj--; }
                    else if (bt == 1) { char ch = res[i - 1]; trace[ntrace++] = (ch >= 'A' && ch <= 'Z') ? (char)(ch + 32) : ch; i--; }
                    else { trace[ntrace++] = '-'; jj--; }
                }
                while (i > 0) { char ch = res[i - 1]; trace[ntrace++] = (ch >= 'A' && ch <= 'Z') ? (char)(ch + 32) : ch; i--; }
                while (jj > 0) { trace[ntrace++] = '-'; jj--; }
            }
        } else if (lane == 0) {
            for (int i = L; i > 0; i--) { char ch = res[i - 1]; trace[ntrace++] = (ch >= 'A' && ch <= 'Z') ? (char)(ch + 32) : ch; }
        }
        ntrace = __shfl_sync(FULL, ntrace, 0);
        __syncwarp();
        // ---- 4. row = '-' * min_col + reversed(trace) + '-' * (BL - max_col - 1), then compressInsertions ----
        const int tail = BL - max_col - 1;
        const int total = min_col + ntrace + tail;
        for (int z = lane; z < total; z += 32) {
            char ch = '-';
            if (z >= min_col && z < min_col + ntrace) ch = trace[ntrace - 1 - (z - min_col)];
            out[z] = ch;
        }
        __syncwarp();
        if (lane == 0) {
            int f_end = -1, b_start = -1;
            for (int z = 0; z < total; z++) if (out[z] >= 'A' && out[z] <= 'Z') { if (f_end < 0) f_end = z; b_start = z + 1; }
            if (f_end >= 0) {
                int wpos = 0;
                for (int z = 0; z < f_end; z++) if (out[z] != '-') out[wpos++] = out[z];  // letters first, then gaps
                for (int z = wpos; z < f_end; z++) out[z] = '-';
                wpos = total - 1;
                for (int z = total - 1; z >= b_start; z--) if (out[z] != '-') out[wpos--] = out[z];  // gaps first, then letters
                for (int z = wpos; z >= b_start; z--) out[z] = '-';
            }
            G.row_len[q] = total;
        }
        __syncwarp();
    }
}

}  // namespace witch
