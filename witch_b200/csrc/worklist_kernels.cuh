// Device-side construction of the envelope work list of the score stage (no host round trip of the per-pair parser
// results): region counts -> exclusive scans -> multi-domain region list -> wave items with sort keys -> radix sort by
// (length class, model rank, envelope length descending) -> groups of up to WAVE_WARPS consecutive items of one HMM.
// Replaces the per-(HMM, chunk) job list of the reference's SearchAlgorithm.search (witch_msa/gcmm/algorithm.py:273-336).
#pragma once
#include "device_types.cuh"
#include "md_kernel.cuh"
#include "wave_kernels.cuh"

namespace witch {

constexpr int WL_BUCKETS = 12;           // 6 length classes x 2 model-size classes
constexpr int WL_SENTINEL_BUCKET = 15;   // unused slots sort behind every real bucket
// Length classes of the envelope work lists. WITCH_WL_FINE=0 merges everything up to 2,048 residues into one launch per
// list: measured on full c2 (round 2) the envelope kernels get 3.4 % faster (3,400 -> 3,286 ms per step: no small, badly
// filled launches), but the STEP gets 7 % slower (6.62 -> 7.08 s): one long persistent launch holds every SM's register
// file until it drains, so the multi-domain branch on the side stream (md_trace_kernel, ~1.1 s of dependent steps) can no
// longer slip in at a launch boundary and run next to the envelope pass. The finer classes stay the default; the align
// stage, which has no side stream next to it, merges its classes in run_wave_host_items.
#ifndef WITCH_WL_FINE
#define WITCH_WL_FINE 1
#endif
__host__ __device__ inline int wl_length_class(int Ls) {
    if (WITCH_WL_FINE) return Ls <= 256 ? 0 : Ls <= 512 ? 1 : Ls <= 1024 ? 2 : Ls <= 2048 ? 3 : Ls <= 4096 ? 4 : 5;
    return Ls <= 2048 ? 3 : Ls <= 4096 ? 4 : 5;
}
__host__ __device__ inline int wl_bucket(int Ls, int M) { return 2 * wl_length_class(Ls) + (M > 13 * 256 ? 1 : 0); }
__host__ __device__ inline unsigned long long wl_key(int bucket, int hrank, int Ls) {
    const int inv = 65535 - (Ls > 65535 ? 65535 : Ls);
    return ((unsigned long long)bucket << 32) | ((unsigned long long)(hrank & 0xffff) << 16) | (unsigned long long)inv;
}

struct WlDesc {   // filled by atomics on the device, read once by the host to size the launches
    int count[16];
    int maxLs[16];
    double cells[16];
    int totals[4];   // [0] single-domain regions (A slots), [1] multi-domain regions
};

// per pair: number of single-domain regions (cntA) and multi-domain regions (cntB); the single-domain regions are the
// envelope items of the first work list, so their bucket descriptor is accumulated here as well
__global__ void region_count_kernel(const PairParse *parse, long long np, int H, const int *M, int *cntA, int *cntB, WlDesc *desc) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const PairParse pp = parse[p];
    const int n = pp.nenv, md = (pp.flags >> 8) & ((1 << n) - 1);
    const int nb = __popc((unsigned)md);
    cntA[p] = n - nb;
    cntB[p] = nb;
    const int Mh = M[(int)(p % H)];
    for (int r = 0; r < n; r++)
        if (!(md >> r & 1)) {
            const int Ls = pp.env_j[r] - pp.env_i[r] + 1, b = wl_bucket(Ls, Mh);
            atomicAdd(&desc->count[b], 1);
            atomicMax(&desc->maxLs[b], Ls);
            atomicAdd(&desc->cells[b], (double)Ls * (double)Mh);
        }
}

// totals of the two scans (last base + last count)
__global__ void scan_totals_kernel(const int *cntA, const int *baseA, const int *cntB, const int *baseB, long long np, WlDesc *desc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        desc->totals[0] = baseA[np - 1] + cntA[np - 1];
        desc->totals[1] = baseB[np - 1] + cntB[np - 1];
    }
}

__global__ void md_list_kernel(const PairParse *parse, long long np, int H, const int *baseB, MdRegion *regs) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const PairParse pp = parse[p];
    int m = baseB[p];
    for (int r = 0; r < pp.nenv; r++)
        if (pp.flags >> (8 + r) & 1) {
            MdRegion R; R.q = (int)(p / H); R.h = (int)(p % H); R.i0 = pp.env_i[r]; R.j0 = pp.env_j[r];
            regs[m++] = R;
        }
}

__device__ __forceinline__ void wl_emit(WaveItem *items, unsigned long long *keys, WlDesc *desc, int slot, int q, int h, int i0, int Ls,
                                        const int *hrank, const int *M) {
    WaveItem it; it.q = q; it.h = h; it.i0 = i0; it.Ls = Ls; it.pair = slot;
    items[slot] = it;
    const int b = wl_bucket(Ls, M[h]);
    keys[slot] = wl_key(b, hrank[h], Ls);
    if (desc) {
        atomicAdd(&desc->count[b], 1);
        atomicMax(&desc->maxLs[b], Ls);
        atomicAdd(&desc->cells[b], (double)Ls * (double)M[h]);
    }
}

// single-domain regions -> items [0, nA) (output slot = item index)
__global__ void items_sd_kernel(const PairParse *parse, long long np, int H, const int *baseA, const int *hrank, const int *M,
                                WaveItem *items, unsigned long long *keys) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const PairParse pp = parse[p];
    int slot = baseA[p];
    for (int r = 0; r < pp.nenv; r++)
        if (!(pp.flags >> (8 + r) & 1))
            wl_emit(items, keys, nullptr, slot++, (int)(p / H), (int)(p % H), pp.env_i[r], pp.env_j[r] - pp.env_i[r] + 1, hrank, M);
}

// envelopes of the multi-domain regions -> the second work list, items [0, nmd*MD_MAXC) with output slots nA + index;
// unused entries get a sentinel key
__global__ void items_md_kernel(const MdRegion *regs, const MdOut *mdout, int nmd, int nA, const int *hrank, const int *M,
                                WaveItem *items, unsigned long long *keys, WlDesc *desc) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= nmd * MD_MAXC) return;
    const int m = z / MD_MAXC, c = z - m * MD_MAXC;
    if (c < mdout[m].nclust) {
        wl_emit(items, keys, desc, z, regs[m].q, regs[m].h, mdout[m].ci[c], mdout[m].cj[c] - mdout[m].ci[c] + 1, hrank, M);
        items[z].pair = nA + z;
    } else {
        WaveItem it; it.q = 0; it.h = 0; it.i0 = 1; it.Ls = 0; it.pair = nA + z;
        items[z] = it;
        keys[z] = wl_key(WL_SENTINEL_BUCKET, 0, 0);
    }
}

// groups over the SORTED items: a group = up to `gw` consecutive items of one (bucket, HMM) run
__global__ void group_head_kernel(const unsigned long long *keys, int n, int *runhead) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    runhead[j] = (j == 0 || (keys[j] >> 16) != (keys[j - 1] >> 16)) ? j : 0;
}
__global__ void group_flag_kernel(const unsigned long long *keys, const int *runstart, int n, int gw, int *gflag) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const bool valid = (int)(keys[j] >> 32) != WL_SENTINEL_BUCKET;
    gflag[j] = (valid && ((j - runstart[j]) % gw) == 0) ? 1 : 0;
}
// group_first[g] = item index; grange[2*b], grange[2*b+1] = first / end group of bucket b
__global__ void group_scatter_kernel(const unsigned long long *keys, const int *gflag, const int *gid, int n, int *group_first, int *grange) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int b = (int)(keys[j] >> 32);
    if (gflag[j]) group_first[gid[j]] = j;
    if (j == 0 || (int)(keys[j - 1] >> 32) != b) {
        grange[2 * b] = gid[j];
        if (j > 0) grange[2 * (int)(keys[j - 1] >> 32) + 1] = gid[j];
    }
    if (j == n - 1) grange[2 * b + 1] = gid[j] + gflag[j];
}

}  // namespace witch
