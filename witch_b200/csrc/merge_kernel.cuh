// Final transitivity merge: WITCH merges every query row into the backbone alignment with
// ExtendedAlignment.merge_in (helpers/alignment_tools.py:1183-1316, driven by gcmm/merger.py:69-78), one query at a
// time on one core. For the rows WITCH produces (exactly `B` regular columns per row, lower-case insertion columns
// between them) the result has a closed form: insertion runs that sit in the same backbone gap are overlaid from the
// left, so gap g becomes a block as wide as the longest run any row has there. Three memory-bound passes:
//   merge_width_kernel : per row, run length per gap -> atomicMax into width[g]              (one warp per row)
//   (exclusive scan of width[g] + 1 over the B+1 gaps -> first output column of each gap block, cub-free, one CTA)
//   merge_scatter_kernel: per row, fill with '-' and scatter residues to their output columns (one warp per row)
// Rows flagged as backbone rows have no insertion columns: every character is a regular column.
#pragma once
#include "device_types.cuh"

namespace witch {

struct MergeWork {
    int nrows;
    const long long *row_off;  // [nrows] start of each input row in `rows`
    const int *row_len;        // [nrows]
    const uint8_t *is_backbone;  // [nrows] 1: every character is a regular column
    const char *rows;
    int B;                     // regular (backbone) columns per row
    int *width;                // [B+1] insertion-block width of each gap (zero-initialised)
    long long *gap_start;      // [B+2] first output column of each gap block; [B+1] = total width
    char *out;                 // [nrows * out_width]
    char *masked;              // [nrows * B] or nullptr
    long long out_width;
};

__device__ __forceinline__ bool merge_is_ins(int ch, bool bb) { return !bb && ch >= 'a' && ch <= 'z'; }

// One warp per row: 32 characters at a time. `g` = regular columns seen before a character, `k` = its rank inside
// the current insertion run; both come from a ballot over the block plus two warp-uniform carries.
template <bool SCATTER>
__global__ void __launch_bounds__(128) merge_rows_kernel(MergeWork W) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= W.nrows) return;
    const char *src = W.rows + W.row_off[row];
    const int len = W.row_len[row];
    const bool bb = W.is_backbone[row] != 0;
    char *dst = SCATTER ? W.out + row * W.out_width : nullptr;
    char *msk = (SCATTER && W.masked) ? W.masked + row * (long long)W.B : nullptr;
    if (SCATTER) {
        for (long long z = lane; z < W.out_width; z += 32) dst[z] = '-';
        __syncwarp();
    }
    int gbase = 0;     // regular columns before this block
    int runbase = 0;   // length of the insertion run that is open at the start of this block
    for (int base = 0; base < len; base += 32) {
        const int idx = base + lane;
        const int ch = idx < len ? (int)(unsigned char)src[idx] : 0;
        const bool valid = idx < len;
        const bool ins = valid && merge_is_ins(ch, bb);
        const unsigned mreg = __ballot_sync(0xffffffffu, valid && !ins);
        const unsigned mval = __ballot_sync(0xffffffffu, valid);
        const unsigned lt = (1u << lane) - 1u;
        const int g = gbase + __popc(mreg & lt);
        // rank in the run: characters since the last regular column before me (or since the block start + carry)
        const unsigned before = mreg & lt;
        const int k = before ? lane - (32 - __clz(before)) : runbase + lane;
        if (ins) {
            if (SCATTER) dst[W.gap_start[g] + k] = (char)ch;
            else {
                // the run's last character reports its length: next character is regular, or the row/block ends
                const bool last = (lane == 31) || !((mval >> (lane + 1)) & 1u) || ((mreg >> (lane + 1)) & 1u);
                if (last && g <= W.B) atomicMax(W.width + g, k + 1);
            }
        } else if (valid && SCATTER && g < W.B) {
            dst[W.gap_start[g] + W.width[g]] = (char)ch;
            if (msk) msk[g] = (char)ch;
        }
        // carries for the next block: regular columns seen, and the length of the run still open at the block's end
        gbase += __popc(mreg);
        if (mreg) runbase = __popc(mval) - (32 - __clz(mreg));
        else runbase += __popc(mval);
    }
}

// Exclusive scan of (width[g] + 1) over g = 0..B (one CTA; B+1 is a few thousand).
__global__ void __launch_bounds__(1024) merge_scan_kernel(const int *width, int B, long long *gap_start) {
    __shared__ long long part[1024];
    const int n = B + 1, T = blockDim.x, tid = threadIdx.x;
    const int per = (n + T - 1) / T, lo = tid * per, hi = min(lo + per, n);
    long long s = 0;
    for (int g = lo; g < hi; g++) s += (long long)width[g] + (g < B ? 1 : 0);
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        long long acc = 0;
        for (int t = 0; t < T; t++) { const long long v = part[t]; part[t] = acc; acc += v; }
        gap_start[n] = acc;   // total width
    }
    __syncthreads();
    long long acc = part[tid];
    for (int g = lo; g < hi; g++) { gap_start[g] = acc; acc += (long long)width[g] + (g < B ? 1 : 0); }
}

}  // namespace witch
