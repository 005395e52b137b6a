// Family D: the multi-domain branch of hmmsearch's domain definition (SURVEY.md 8(a) "Score semantics" item 6).
//
// When a region of the parser pass fails the single-domain test, hmmsearch 3.1b2 re-runs a multihit Forward over the
// region, samples 200 stochastic tracebacks from it (Easel's "fast" LCG, re-seeded with 42 for every region), derives a
// position-specific null2 from the traces and clusters the sampled domains (single linkage) into envelopes. Which path
// a trace takes is decided by comparisons `roll < cumulative probability`, so this kernel evaluates the region Forward
// exactly the way HMMER's SSE code does: same 4-way striped layout, one FP32 rounding per mul/add in the same order
// (__fmul_rn/__fadd_rn: no FMA contraction), the serial D->D passes with their early exit, specials evaluated in
// double and rounded per statement, sparse rescaling at E > 1e4, and the same striped parameter tables (hmm_profile.cpp).
// It affects only the few regions that are flagged (< 0.1 % of the pairs with hmmbuild-made profiles), so the mapping is
// the simplest one that keeps the arithmetic sequential where HMMER's is: ONE WARP PER REGION, lanes share the
// embarrassingly parallel parts (M/I cells of a row, the E-state choice, null2 accumulation, link tests), lanes 0..3 run
// the four stripe lanes of the serial D chain, lane 0 walks the traces. Scratch (the full Forward matrix) is in HBM.
#pragma once
#include "device_types.cuh"

namespace witch {

constexpr int MD_MAXC = 16;       // envelopes kept per multi-domain region
constexpr int MD_NSAMPLES = 200;  // hmmsearch's default number of sampled traces
constexpr int MD_MAXDOM = 64;     // domains of one trace
constexpr int MD_MAXSIG = 64;     // significant clusters of one region

struct MdRegion { int q, h, i0, j0; };   // region i0..j0 (1-based) of query q against HMM h
struct MdOut {
    int nclust;            // envelopes of the region (<= MD_MAXC)
    int flags;             // WITCH_FLAG_ENVCAP if more were found
    float regcorr;         // sum over the region of the position-specific ln null2 (from the traces)
    int ci[MD_MAXC], cj[MD_MAXC];
    float ccorr[MD_MAXC];  // sum of ln null2 over each envelope
    long long clk[4];      // device clock ticks spent in Forward / traces / clustering (WITCH_TIMING diagnostics), region size
};
// One batch of regions: region j of the batch is regions[order[begin + j]], its scratch slot starts at scratch +
// slot_off[begin + j] (slots are sized per region and packed; the host cuts batches that fit the scratch budget).
struct MdWork {
    const MdRegion *regions;
    const int *order;
    const long long *slot_off;
    int begin, end;
    unsigned *counter;
    char *scratch;
    int Qcap, nsp_cap;
    int spread;            // trace kernel: one walker every `spread` threads (32 = one region per warp, 1 = one per thread)
    MdOut *out;
};

constexpr int MD_EBLK = 8;        // striped vectors per block of the two-level E scan
__host__ __device__ inline int md_nblocks(int Q) { return (Q + MD_EBLK - 1) / MD_EBLK; }
struct MdLayout { long long hdr, dp, xmx, acc, sp, asg, epc, n2log, bsum, total; };
// slot of one region of Lcap residues against a model of Mcap nodes (Qcap striped vectors per row)
__host__ __device__ inline MdLayout md_layout(int Lcap, int Qcap, int Mcap, int nsp_cap) {
    MdLayout l;
    long long o = 0;
    l.hdr = o; o += 64;   // ints: [0] sampled domains, [1] overflow flag, [2..] clock ticks (diagnostics)
    l.dp = o; o += (long long)(Lcap + 1) * Qcap * 12 * 4;
    l.xmx = o; o += (long long)(Lcap + 1) * 8 * 4;
    l.acc = o; o += (long long)(Lcap + 4) * 4;
    l.sp = o; o += (long long)nsp_cap * 5 * 4;
    l.asg = o; o += (long long)nsp_cap * 4;
    l.epc = o; o += (long long)((Lcap > Mcap ? Lcap : Mcap) + 4) * 4;
    l.n2log = o; o += (long long)nsp_cap * 20 * 4;   // null2 odds of every sampled domain (K <= 20 floats each)
    o = (o + 7) / 8 * 8;
    l.bsum = o; o += (long long)(Lcap + 1) * md_nblocks(Qcap) * 8;   // per row: E-scan mass of every block of MD_EBLK striped vectors (double)
    l.total = (o + 255) / 256 * 256;
    return l;
}

#ifdef WITCH_HOST_SIM
static inline float md_mul(float a, float b) { return a * b; }
static inline float md_add(float a, float b) { return a + b; }
#else
__device__ __forceinline__ float md_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float md_add(float a, float b) { return __fadd_rn(a, b); }
#endif

#ifdef WITCH_HOST_SIM
static inline void md_prefetch(const void *) {}
static inline void md_prefetch_l2(const void *) {}
#else
__device__ __forceinline__ void md_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ void md_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
#endif

__device__ __forceinline__ unsigned md_mix3(unsigned a, unsigned b, unsigned c) {
    a -= b; a -= c; a ^= (c >> 13);
    b -= c; b -= a; b ^= (a << 8);
    c -= a; c -= b; c ^= (b >> 13);
    a -= b; a -= c; a ^= (c >> 12);
    b -= c; b -= a; b ^= (a << 16);
    c -= a; c -= b; c ^= (b >> 5);
    a -= b; a -= c; a ^= (c >> 3);
    b -= c; b -= a; b ^= (a << 10);
    c -= a; c -= b; c ^= (b >> 15);
    return c;
}
__device__ __forceinline__ double md_rand(unsigned &x) { x = x * 69069u + 1u; return (double)x * 2.3283064365386963e-10; }

// esl_vec_FNorm (n < 8) + esl_rnd_FChoose: float normalisation, double running sum against one roll
__device__ __forceinline__ int md_choose(unsigned &rng, float *p, int n) {
    float s = 0.f;
    for (int i = 0; i < n; i++) s = md_add(s, p[i]);
    if (s != 0.f) { for (int i = 0; i < n; i++) p[i] = p[i] / s; }
    else { for (int i = 0; i < n; i++) p[i] = 1.0f / (float)n; }
    const double roll = md_rand(rng);
    double norm = 0.0, sum = 0.0;
    for (int i = 0; i < n; i++) norm += (double)p[i];
    for (int i = 0; i < n; i++) {
        sum += (double)p[i];
        if (sum / norm > roll) return i;
    }
    return n - 1;
}

// members of a degenerate symbol as a bit mask over the canonical residues (Easel's alphabets)
__device__ __forceinline__ unsigned md_degen_mask(int Kp, int code) {
    if (Kp == 29) {  // amino: B=ND J=IL Z=QE O=K U=C X=all   (codes 21..26)
        const unsigned m[6] = {(1u << 11) | (1u << 2), (1u << 7) | (1u << 9), (1u << 13) | (1u << 3), 1u << 8, 1u << 1, 0xFFFFFu};
        return (code >= 21 && code <= 26) ? m[code - 21] : 0u;
    }
    const unsigned m[11] = {5, 10, 3, 12, 6, 9, 11, 14, 7, 13, 15};   // R Y M K S W H B V D N (codes 5..15)
    return (code >= 5 && code <= 15) ? m[code - 5] : 0u;
}

// link test of the sampled-domain clustering: >= 80 % overlap of the shorter one on the sequence and on the model (the
// model-side overlap is counted without the "+1", as the 3.1b2 binary does), and start OR end diagonals within 4
__device__ __forceinline__ bool md_link(int ai, int aj, int ak, int am, int bi, int bj, int bk, int bm) {
    int nov = min(aj, bj) - max(ai, bi) + 1;
    int n = min(aj - ai, bj - bi) + 1;
    if ((float)nov / (float)n < 0.8f) return false;
    nov = min(am, bm) - max(ak, bk);
    n = min(am - ak, bm - bk) + 1;
    if ((float)nov / (float)n < 0.8f) return false;
    if (abs((ai - ak) - (bi - bk)) <= 4) return true;
    return abs((aj - am) - (bj - bm)) <= 4;
}

constexpr int MD_WARPS = 4;   // warps per CTA of the warp-per-region kernels

// everything a kernel needs to know about region j of the batch
struct MdCtx {
    MdRegion R; MdOut *out;
    int M, Q, L, Lr, K;
    const float *tfv, *rfv;
    const uint8_t *rd;
    float pmove, ploop;
    size_t RW;
    float *dp, *xmx, *acc;
    int *hdr, *spi, *spj, *spk, *spm, *spt, *asg, *epc;
    float *n2log;
    double *bsum;
};
__device__ __forceinline__ MdCtx md_ctx(const DevEhmm &E, const DevQueries &Qs, const MdWork &W, int j) {
    MdCtx c;
    const int ridx = W.order[W.begin + j];
    c.R = W.regions[ridx];
    c.out = W.out + ridx;
    c.M = E.M[c.R.h]; c.Q = E.oQ[c.R.h];
    c.tfv = E.otfv + E.otoff[c.R.h];
    c.rfv = E.orfv + E.oroff[c.R.h];
    c.L = Qs.len[c.R.q]; c.Lr = c.R.j0 - c.R.i0 + 1;
    c.K = (E.Kp == 29) ? 20 : 4;
    c.rd = Qs.dsq + Qs.off[c.R.q] + (c.R.i0 - 1);   // rd[i-1] = dense code of residue i of the region
    c.pmove = 3.0f / ((float)c.L + 3.0f); c.ploop = 1.0f - c.pmove;
    c.RW = (size_t)c.Q * 12;                       // floats per row: [q][M|D|I][4]
    const MdLayout lay = md_layout(c.Lr, c.Q, c.M, W.nsp_cap);
    char *slot = W.scratch + W.slot_off[W.begin + j];
    c.hdr = (int *)(slot + lay.hdr);
    c.dp = (float *)(slot + lay.dp);
    c.xmx = (float *)(slot + lay.xmx);           // 8 floats per row: E N J B C SCALE - -
    c.acc = (float *)(slot + lay.acc);
    c.spi = (int *)(slot + lay.sp); c.spj = c.spi + W.nsp_cap; c.spk = c.spj + W.nsp_cap; c.spm = c.spk + W.nsp_cap; c.spt = c.spm + W.nsp_cap;
    c.asg = (int *)(slot + lay.asg);
    c.epc = (int *)(slot + lay.epc);
    c.n2log = (float *)(slot + lay.n2log);
    c.bsum = (double *)(slot + lay.bsum);
    return c;
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 1: the region's Forward matrix, HMMER's arithmetic operation by operation. One warp per region.
__global__ void __launch_bounds__(MD_WARPS * 32) md_forward_kernel(DevEhmm E, DevQueries Qs, MdWork W) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    WITCH_DYN_SMEM(float, md_smem);   // per warp: the current row's M and D vectors, [Qcap][4] each
    float *sMv = md_smem + (size_t)w * 8 * W.Qcap, *sDv = sMv + (size_t)4 * W.Qcap;
    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(W.counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= W.end - W.begin) break;
        const MdCtx cx = md_ctx(E, Qs, W, item);
        const MdRegion R = cx.R;
        const int M = cx.M, Q = cx.Q, Lr = cx.Lr;
        const float *tfv = cx.tfv, *rfv = cx.rfv;
        const uint8_t *rd = cx.rd;
        const float pmove = cx.pmove, ploop = cx.ploop;
        const size_t RW = cx.RW;
        float *dp = cx.dp, *xmx = cx.xmx;
        (void)R;
        const long long clk0 = clock64();
        // ====================== Forward over the region (multihit, length model of the whole sequence) ======================
        // A row's M and D vectors are staged in shared memory ([q][4] each) for the serial part; the global matrix row
        // receives the final values with coalesced stores.
        for (int c = lane; c < Q * 12; c += 32) dp[c] = 0.f;
        if (lane == 0) { xmx[0] = 0.f; xmx[1] = 1.f; xmx[2] = 0.f; xmx[3] = pmove; xmx[4] = 0.f; xmx[5] = 1.f; }
        float fN = 1.0f, fB = pmove, fJ = 0.f, fC = 0.f;
        __syncwarp();
        for (int i = 1; i <= Lr; i++) {
            const float *__restrict__ rp = rfv + (size_t)Qs.symrow[rd[i - 1]] * Q * 4;
            const float *__restrict__ prev = dp + (size_t)(i - 1) * RW;
            float *__restrict__ cur = dp + (size_t)i * RW;
            // M and I cells: independent across (q, z)
            for (int c = lane; c < Q * 4; c += 32) {
                const int q = c >> 2, z = c & 3;
                float mpv, dpv, ipv;
                if (q > 0) { const float *v = prev + (size_t)(q - 1) * 12 + z; mpv = v[0]; dpv = v[4]; ipv = v[8]; }
                else if (z > 0) { const float *v = prev + (size_t)(Q - 1) * 12 + z - 1; mpv = v[0]; dpv = v[4]; ipv = v[8]; }
                else { mpv = 0.f; dpv = 0.f; ipv = 0.f; }
                const float *__restrict__ tp = tfv + (size_t)q * 28 + z;
                const float mp2 = prev[(size_t)q * 12 + z], ip2 = prev[(size_t)q * 12 + 8 + z];
                float sv = md_mul(fB, tp[0]);
                sv = md_add(sv, md_mul(mpv, tp[4]));
                sv = md_add(sv, md_mul(ipv, tp[8]));
                sv = md_add(sv, md_mul(dpv, tp[12]));
                sv = md_mul(sv, rp[c]);
                sMv[c] = sv;
                const float dc = md_mul(sv, tp[16]);   // M->D into the next column
                if (q + 1 < Q) sDv[c + 4] = dc;
                else if (z < 3) sDv[z + 1] = dc;       // wraps into the next stripe lane of vector 0
                if (c == 0) sDv[0] = 0.f;
                cur[(size_t)q * 12 + 8 + z] = md_add(md_mul(mp2, tp[20]), md_mul(ip2, tp[24]));
            }
            __syncwarp();
            // D->D paths and the E sum: serial chains per stripe lane, exactly in HMMER's order. Lanes 0-3 run the D chains,
            // lanes 4-7 the running sum of the M cells (HMMER adds all M cells first, then the D cells, lane by lane).
            const int z4 = lane & 3;
            const float *__restrict__ td = tfv + (size_t)Q * 28 + z4;
            float d = 0.f, xe = 0.f;
            if (lane < 4) {
                for (int q0 = 0; q0 < Q; q0 += 8) {
                    float in[8], t8[8];
                    const int nq8 = min(8, Q - q0);
#pragma unroll
                    for (int u = 0; u < 8; u++) if (u < nq8) { in[u] = sDv[(q0 + u) * 4 + z4]; t8[u] = td[(size_t)(q0 + u) * 4]; }
#pragma unroll
                    for (int u = 0; u < 8; u++) if (u < nq8) { in[u] = md_add(d, in[u]); d = md_mul(in[u], t8[u]); }
#pragma unroll
                    for (int u = 0; u < 8; u++) if (u < nq8) sDv[(q0 + u) * 4 + z4] = in[u];
                }
            } else if (lane < 8) {
                for (int q0 = 0; q0 < Q; q0 += 8) {
                    float in[8];
                    const int nq8 = min(8, Q - q0);
#pragma unroll
                    for (int u = 0; u < 8; u++) if (u < nq8) in[u] = sMv[(q0 + u) * 4 + z4];
#pragma unroll
                    for (int u = 0; u < 8; u++) if (u < nq8) xe = md_add(xe, in[u]);
                }
            }
            for (int j = 1; j < 4; j++) {
                float din = __shfl_up_sync(FULL, d, 1);
                if (lane == 0) din = 0.f;
                int changed = 0;
                if (lane < 4) {
                    d = din;
                    // (once the carried term is exactly 0 the rest of the pass changes nothing: leave it)
                    for (int q0 = 0; q0 < Q && d != 0.f; q0 += 8) {
                        float in[8], t8[8];
                        const int nq8 = min(8, Q - q0);
#pragma unroll
                        for (int u = 0; u < 8; u++) if (u < nq8) { in[u] = sDv[(q0 + u) * 4 + z4]; t8[u] = td[(size_t)(q0 + u) * 4]; }
#pragma unroll
                        for (int u = 0; u < 8; u++) if (u < nq8) {
                            const float v = md_add(d, in[u]);
                            if (in[u] < v) changed = 1;
                            in[u] = v;
                            d = md_mul(d, t8[u]);
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++) if (u < nq8) sDv[(q0 + u) * 4 + z4] = in[u];
                    }
                }
                const unsigned any = __ballot_sync(FULL, changed != 0);
                if (M >= 100 && any == 0u) break;
            }
            xe = __shfl_sync(FULL, xe, 4 + z4);   // lanes 0-3 continue the sum of their stripe lane with the D cells
            if (lane < 4) {
                for (int q0 = 0; q0 < Q; q0 += 8) {
                    float in[8];
                    const int nq8 = min(8, Q - q0);
#pragma unroll
                    for (int u = 0; u < 8; u++) if (u < nq8) in[u] = sDv[(q0 + u) * 4 + z4];
#pragma unroll
                    for (int u = 0; u < 8; u++) if (u < nq8) xe = md_add(in[u], xe);
                }
            }
            const float x1 = __shfl_sync(FULL, xe, 1), x2 = __shfl_sync(FULL, xe, 2), x3 = __shfl_sync(FULL, xe, 3);
            float fE = md_add(md_add(__shfl_sync(FULL, xe, 0), x1), md_add(x2, x3));
            // specials: double evaluation, one float rounding per statement
            fJ = (float)((double)ploop * (double)fJ + 0.5 * (double)fE);
            fN = (float)((double)ploop * (double)fN);
            fC = (float)((double)ploop * (double)fC + 0.5 * (double)fE);
            fB = (float)((double)pmove * (double)fN + (double)pmove * (double)fJ);
            float scale = 1.0f, inv = 1.0f;
            const bool resc = (double)fE > 1.0e4;   // sparse rescaling (warp-uniform decision)
            if (resc) {
                const double e = (double)fE;
                fN = (float)((double)fN / e); fC = (float)((double)fC / e); fJ = (float)((double)fJ / e); fB = (float)((double)fB / e);
                inv = (float)(1.0 / e);
                scale = fE;
                fE = 1.0f;
            }
            __syncwarp();
            // final M and D cells of the row -> global matrix (the I cells are there already)
            for (int c = lane; c < Q * 4; c += 32) {
                const int q = c >> 2, z = c & 3;
                float m = sMv[c], dd = sDv[c];
                if (resc) { m = md_mul(m, inv); dd = md_mul(dd, inv); cur[(size_t)q * 12 + 8 + z] = md_mul(cur[(size_t)q * 12 + 8 + z], inv); }
                cur[(size_t)q * 12 + z] = m;
                cur[(size_t)q * 12 + 4 + z] = dd;
            }
            if (lane == 0) {
                float *x = xmx + (size_t)i * 8;
                x[0] = fE; x[1] = fN; x[2] = fJ; x[3] = fB; x[4] = fC; x[5] = scale;
            }
            {   // E-scan mass of every block of MD_EBLK vectors of the stored row (what a trace's choice of the E state's
                // cell adds up: (double)(cell * 1/E), M cells then D cells of each vector), for the two-level scan
                const float enorm = 1.0f / fE;
                const int nblk = md_nblocks(Q);
                double *bs = cx.bsum + (size_t)i * nblk;
                for (int b = lane; b < nblk; b += 32) {
                    double sum = 0.0;
                    const int q1 = min(Q, (b + 1) * MD_EBLK);
                    for (int q = b * MD_EBLK; q < q1; q++) {
#pragma unroll
                        for (int r = 0; r < 8; r++) {
                            float v = (r < 4) ? sMv[q * 4 + r] : sDv[q * 4 + r - 4];
                            if (resc) v = md_mul(v, inv);
                            sum += (double)md_mul(v, enorm);
                        }
                    }
                    bs[b] = sum;
                }
            }
            __syncwarp();
        }

        if (lane == 0) { cx.hdr[2] = (int)((clock64() - clk0) >> 10); }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 2: the 200 stochastic traces of a region and its sampled domains. The walk is sequential by nature (one
// random-number stream per region, every choice depends on the previous one), so ONE THREAD per region: the parallelism
// is across the regions of the batch (ordered by size, so the lanes of a warp walk matrices of similar shape), and what
// bounds a walk is the latency of one dependent memory access per step. The step is therefore written so that
//  * the three cell states (M, D, I) share ONE branch-free body: cell / transition addresses are selected by state, all
//    loads of a step are independent of each other (one memory latency per step whatever mix of states a warp holds),
//    the choice among 4 or 2 paths is the same code with zero-padded paths;
//  * the emission odds of a step (null2 sums) are loaded one step late, together with the next step's cell;
//  * the cell the walk needs MD_PF steps later if it stays on the diagonal (M -> M, by far the most likely path) is
//    prefetched into L2 every step;
//  * N -> N steps (no random number, no effect) are skipped: reaching N ends the trace;
//  * the choice of an E state's cell scans its row MD_ESCAN striped vectors per iteration (a state of its own), so a lane
//    that scans does not hold up the lanes that walk.
// Per sampled domain the kernel logs its coordinates and its null2 odds; the per-residue accumulation happens in the
// clustering kernel.
constexpr int MD_ESCAN = 4;   // striped vectors examined per iteration of the E scan's second level
#ifndef WITCH_MD_PF
#define WITCH_MD_PF 6
#endif
constexpr int MD_PF = WITCH_MD_PF;   // diagonal prefetch distance (steps)

// esl_vec_FNorm + esl_rnd_FChoose for n = 2 or 4 paths (p2 = p3 = 0 when n = 2: adding an exact zero changes no rounding)
__device__ __forceinline__ int md_choose4(unsigned &rng, float p0, float p1, float p2, float p3, int n) {
    const float s = md_add(md_add(md_add(p0, p1), p2), p3);
    if (s != 0.f) { p0 = p0 / s; p1 = p1 / s; p2 = p2 / s; p3 = p3 / s; }
    else { const float u = 1.0f / (float)n; p0 = u; p1 = u; p2 = (n == 4) ? u : 0.f; p3 = p2; }
    const double roll = md_rand(rng);
    const double d0 = (double)p0, d1 = d0 + (double)p1, d2 = d1 + (double)p2, norm = d2 + (double)p3;
    int c = n - 1;
    if (n == 4 && d2 / norm > roll) c = 2;
    if (d1 / norm > roll) c = 1;
    if (d0 / norm > roll) c = 0;
    return c;
}

template <int K>
__global__ void __launch_bounds__(128) md_trace_kernel(DevEhmm E, DevQueries Qs, MdWork W) {
    // A batch of few regions gives every walker a warp of its own (no serialisation of the three step bodies across
    // regions, a private L1 footprint); a batch of many regions packs up to 32 walkers into a warp.
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = gt / W.spread;
    if (gt % W.spread != 0 || j >= W.end - W.begin) return;
    const MdCtx cx = md_ctx(E, Qs, W, j);
    const int Q = cx.Q, Lr = cx.Lr;
    const float *__restrict__ dp = cx.dp, *__restrict__ xmx = cx.xmx;
    const float *__restrict__ nt8 = E.ont8 + E.onoff[cx.R.h] * 8, *__restrict__ nem = E.onem + E.onoff[cx.R.h] * K;
    const float pmove = cx.pmove, ploop = cx.ploop;
    const size_t RW = cx.RW;
    float *n2log = cx.n2log;
    const long long clk0 = clock64();
    enum { tM = 1, tD = 2, tI = 3, tS = 4, tN = 5, tB = 6, tE = 7, tC = 8, tJ = 10, tESCAN = 11, tEBLK = 12 };
    unsigned rng = md_mix3(42u, 87654321u, 12345678u);
    if (rng == 0u) rng = 42u;
    int nsp = 0, oflow = 0, t = 0;
    int i = Lr, k = 0, s0 = tC, ndom = 0;
    int sqto = 0, sqfrom = 0, hto = 0, hfrom = 0, Ld = 0;
    int cq = 0, cr = 0;        // striped position of node k: (k-1) % Q, (k-1) / Q
    int eq = 0;                // E scan: next striped vector
    double eroll = 0.0, esum = 0.0;
    float enorm = 0.f;
    bool pend = false;         // an emitting step whose odds are not in sums[] yet
    int pek = 0;               // its node
    double sums[K];            // sum over the running domain's emitting states of the match odds of every residue
#pragma unroll
    for (int x = 0; x < K; x++) sums[x] = 0.0;
    for (;;) {
        int s1;
        if (s0 == tM || s0 == tD || s0 == tI) {
            // ---- one step out of a cell state: every load below is independent of the others
            const bool isM = s0 == tM, isD = s0 == tD, isI = s0 == tI;
            int pq = cq - 1, pr = cr;                  // position of node k-1
            if (pq < 0) { pq += Q; pr--; }
            const bool hasprev = pr >= 0;
            const int lq = isI ? cq : (hasprev ? pq : 0), lr = isI ? cr : (hasprev ? pr : 0);   // cell to look at
            const int lrow = isD ? i : i - 1;
            const float *cell = dp + (size_t)lrow * RW + (size_t)lq * 12 + lr;
            // transitions: {BM,MM,IM,DM into k} for M, {MD,DD} of node k-1 for D, {MI,II} of node k for I -- one 128-bit load
            const float4 T = __ldg(reinterpret_cast<const float4 *>(nt8 + (size_t)(isD ? (hasprev ? k - 1 : 0) : k) * 8 + (isM ? 0 : 4)));
            float vM = cell[0], vD = cell[4], vI = cell[8];
            const float xB = xmx[(size_t)(i - 1) * 8 + 3];
            float em[K];
#pragma unroll
            for (int x = 0; x < K; x += 4) {
                const float4 e4 = __ldg(reinterpret_cast<const float4 *>(nem + (size_t)pek * K + x));
                em[x] = e4.x; em[x + 1] = e4.y; em[x + 2] = e4.z; em[x + 3] = e4.w;
            }
            {   // the cell MD_PF steps down the diagonal
                int fq = cq - 1 - MD_PF, fr = cr;
                if (fq < 0) { fq += Q; fr--; }
                if (fq < 0) { fq += Q; fr--; }
                const int frow = i - 1 - MD_PF;
                if (frow >= 0 && fr >= 0 && fq >= 0) md_prefetch(dp + (size_t)frow * RW + (size_t)fq * 12 + fr);
            }
            const float t0 = isI ? T.z : T.x, t1 = isI ? T.w : T.y, t2 = T.z, t3 = T.w;
            if (!hasprev && !isI) { vM = 0.f; vD = 0.f; vI = 0.f; }
            if (pend) {
#pragma unroll
                for (int x = 0; x < K; x++) sums[x] += (double)em[x];
                pend = false;
            }
            const float p0 = md_mul(isM ? xB : vM, t0);   // (node 1 has no predecessor cell: those paths are exactly 0)
            const float p1 = md_mul(isM ? vM : isD ? vD : vI, t1);
            const float p2 = isM ? md_mul(vI, t2) : 0.f;
            const float p3 = isM ? md_mul(vD, t3) : 0.f;
            const int c = md_choose4(rng, p0, p1, p2, p3, isM ? 4 : 2);
            // M: B M I D | D: M D | I: M I
            s1 = isM ? ((c == 0) ? tB : (c == 1) ? tM : (c == 2) ? tI : tD) : (c == 0) ? tM : (isD ? tD : tI);
            if (!isI) { k--; cq = pq; cr = pr; }
            if (!isD) i--;
        } else if (s0 == tEBLK) {
            // E state, level 1: which block of MD_EBLK vectors holds the roll (running sum of the blocks' masses)
            const int nblk = md_nblocks(Q);
            const double *bs = cx.bsum + (size_t)i * nblk;
            const int b1 = min(eq + 16, nblk);
            double m16[16];
#pragma unroll
            for (int z = 0; z < 16; z++) m16[z] = (eq + z < b1) ? bs[eq + z] : 0.0;
            bool hit = false;
#pragma unroll
            for (int z = 0; z < 16; z++) {
                if (!hit && eq + z < b1) {
                    if (esum + m16[z] > eroll) hit = true; else { esum += m16[z]; }
                    if (hit) eq += z;
                }
            }
            if (!hit) {
                eq = b1;
                if (eq < nblk) continue;
                eq = Q; s0 = tESCAN;   // (the roll lies beyond the row's whole mass, i.e. within rounding of 1: level 2 resolves it)
                continue;
            }
            eq *= MD_EBLK;          // first vector of the block; esum = mass of everything before it
            s0 = tESCAN;
            continue;
        } else if (s0 == tESCAN) {
            // E state, level 2: one roll against the running sum over the M/D cells of row i from vector eq on, in HMMER's
            // striped order (the running sum differs from HMMER's strictly sequential one by the association of the
            // block masses: a choice can differ only when the roll lies within ~1e-13 of a cumulative probability)
            const float *row = dp + (size_t)i * RW;
            bool found = false;
            int kk = 1, ss = tM;
            const int qe = min(eq + MD_ESCAN, Q);
#pragma unroll
            for (int q = eq; q < qe; q++) {
                const float4 m4 = *reinterpret_cast<const float4 *>(row + (size_t)q * 12), d4 = *reinterpret_cast<const float4 *>(row + (size_t)q * 12 + 4);
                const float v8[8] = {m4.x, m4.y, m4.z, m4.w, d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    esum += (double)md_mul(v8[r], enorm);
                    if (!found && esum > eroll) { kk = (r & 3) * Q + q + 1; ss = (r < 4) ? tM : tD; found = true; }
                }
            }
            eq = qe;
            if (!found && eq < Q) continue;   // (rounding at a block boundary: carry on into the next block)
            // (not found at all: the roll sits within rounding of 1; HMMER would wrap around -- first cell)
            k = kk; s1 = ss;
            cq = (kk - 1) % Q; cr = (kk - 1) / Q;
        } else {
            // ---- special states C, J, B (N never becomes the current state: reaching it ends the trace)
            const bool isB = s0 == tB, isC = s0 == tC;
            const float *x1 = xmx + (size_t)i * 8, *x0 = xmx + (size_t)(i > 0 ? i - 1 : 0) * 8;
            const float a0 = isB ? x1[1] : isC ? x0[4] : x0[2];
            const float b0 = x1[0], b2 = x1[2], b5 = x1[5];
            const float p0 = md_mul(isB ? pmove : ploop, a0);
            const float p1 = isB ? md_mul(pmove, b2) : md_mul(md_mul(0.5f, b0), b5);
            const int c = md_choose4(rng, p0, p1, 0.f, 0.f, 2);
            s1 = isB ? (c == 0 ? tN : tJ) : (c == 0 ? s0 : tE);
            if (s1 == s0) i--;   // C -> C, J -> J
        }
        // ---- consequences of the step
        if (s1 == tE) {   // a new domain starts (seen from its end): draw the roll, then scan row i over the next iterations
            eroll = md_rand(rng);
            enorm = 1.0f / xmx[(size_t)i * 8];
            esum = 0.0; eq = 0;
            sqto = 0; sqfrom = 0; hto = 0; hfrom = 0; Ld = 0;
#pragma unroll
            for (int x = 0; x < K; x++) sums[x] = 0.0;
            s0 = tEBLK;
            continue;
        }
        if (s1 == tM || s1 == tI) {   // (3.1b2 counts a residue emitted by I_k in the MATCH cell of node k)
            if (s1 == tM) { if (sqto == 0) { sqto = i; hto = k; } sqfrom = i; hfrom = k; }
            Ld++;
            pend = true; pek = k;
        } else if (s1 == tB) {
            // domain sqfrom..sqto complete: log it with its null2 odds (accumulated per residue by the clustering kernel)
            if (ndom < MD_MAXDOM && nsp + ndom < W.nsp_cap) {
                const int z = nsp + ndom;
                cx.spi[z] = sqfrom + cx.R.i0 - 1; cx.spj[z] = sqto + cx.R.i0 - 1; cx.spk[z] = hfrom; cx.spm[z] = hto; cx.spt[z] = t;
                const float nrm = (float)(1.0 / (double)(float)Ld);
#pragma unroll
                for (int x = 0; x < K; x++) n2log[(size_t)z * K + x] = (float)(sums[x] * (double)nrm);
                ndom++;
            } else oflow = 1;
        }
        s0 = s1;
        if (s1 == tN) {
            // end of trace t (N -> ... -> S consumes no random number): its domains enter the ensemble in sequence order
            // (they were found last to first)
            for (int a = nsp, b = nsp + ndom - 1; a < b; a++, b--) {
                int tmp;
                tmp = cx.spi[a]; cx.spi[a] = cx.spi[b]; cx.spi[b] = tmp; tmp = cx.spj[a]; cx.spj[a] = cx.spj[b]; cx.spj[b] = tmp;
                tmp = cx.spk[a]; cx.spk[a] = cx.spk[b]; cx.spk[b] = tmp; tmp = cx.spm[a]; cx.spm[a] = cx.spm[b]; cx.spm[b] = tmp;
                for (int x = 0; x < K; x++) { const float f = n2log[(size_t)a * K + x]; n2log[(size_t)a * K + x] = n2log[(size_t)b * K + x]; n2log[(size_t)b * K + x] = f; }
            }
            nsp += ndom;
            if (++t >= MD_NSAMPLES) break;
            i = Lr; k = 0; s0 = tC; ndom = 0; cq = 0; cr = 0;
        }
    }
    cx.hdr[0] = nsp; cx.hdr[1] = oflow; cx.hdr[3] = (int)((clock64() - clk0) >> 10);
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 3: ln of the mean null2 odds, single-linkage clustering of the sampled domains, envelopes. One warp per region.
__global__ void __launch_bounds__(MD_WARPS * 32) md_cluster_kernel(DevEhmm E, DevQueries Qs, MdWork W) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    __shared__ int s_sig[MD_WARPS][MD_MAXSIG][6];     // i, j, k, m, count, cluster order
    __shared__ unsigned s_bits[MD_WARPS][8];
    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(W.counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= W.end - W.begin) break;
        const MdCtx cx = md_ctx(E, Qs, W, item);
        const MdRegion R = cx.R;
        MdOut *out = cx.out;
        const int M = cx.M, Lr = cx.Lr;
        float *acc = cx.acc;
        int *spi = cx.spi, *spj = cx.spj, *spk = cx.spk, *spm = cx.spm, *spt = cx.spt, *asg = cx.asg, *epc = cx.epc;
        const int nsp = cx.hdr[0];
        int oflow = cx.hdr[1];
        const long long clk1 = 0, clk0 = 0, clk2 = clock64();
        (void)clk1; (void)clk0;
        // per-residue null2: every trace adds the odds of the domain that covers the residue (HMMER gives a domain's first
        // residue the neutral 1.0 as well), or 1.0; float additions in trace order, as hmmsearch accumulates them
        {
            const int K = cx.K;
            const float *n2log = cx.n2log;
            for (int p = 1 + lane; p <= Lr; p += 32) {
                const int pa = p + R.i0 - 1;   // absolute position
                const int code = Qs.symrow[cx.rd[p - 1]];
                const unsigned mask = code < K ? 0u : md_degen_mask(E.Kp, code);
                float a = 0.f;
                int d = 0;
                for (int t = 0; t < MD_NSAMPLES; t++) {
                    float v = 1.0f;
                    for (; d < nsp && spt[d] == t; d++)
                        if (spi[d] < pa && pa <= spj[d]) {
                            const float *n2 = n2log + (size_t)d * K;
                            if (code < K) v = n2[code];
                            else {
                                float sg = 0.f; int n = 0;
                                for (int x = 0; x < K; x++) if (mask >> x & 1u) { sg += n2[x]; n++; }
                                v = n ? sg / (float)n : 1.0f;
                            }
                        }
                    a = md_add(a, v);
                }
                acc[p] = a;
            }
        }
        __syncwarp();
        // ln of the mean null2 odds per residue -> acc[]; sum over the region
        float regc = 0.f;
        for (int p = 1 + lane; p <= Lr; p += 32) { const float v = logf(acc[p] / (float)MD_NSAMPLES); acc[p] = v; regc += v; }
        for (int o = 16; o > 0; o >>= 1) regc += __shfl_xor_sync(FULL, regc, o);
        __syncwarp();

        // ====================== single-linkage clustering of the sampled domains ======================
        for (int a = lane; a < nsp; a += 32) asg[a] = a;
        __syncwarp();
        for (;;) {   // label = largest vertex index of the component (Easel numbers clusters from the last vertex down)
            int changed = 0;
            for (int a = lane; a < nsp; a += 32) {
                const int ai = spi[a], aj = spj[a], ak = spk[a], am = spm[a];
                int lab = asg[a];
                for (int b = 0; b < nsp; b++) {
                    const int lb = asg[b];
                    if (lb > lab && md_link(ai, aj, ak, am, spi[b], spj[b], spk[b], spm[b])) { lab = lb; changed = 1; }
                }
                asg[a] = lab;
            }
            __syncwarp();
            if (__ballot_sync(FULL, changed != 0) == 0u) break;
        }
        // clusters in Easel's order (descending representative), statistics of each
        int nsig = 0, ncl = 0;
        for (int rep = nsp - 1; rep >= 0; rep--) {
            if (asg[rep] != rep) continue;
            const int corder = ncl++;
            // members, distinct traces, extents
            if (lane < 8) s_bits[w][lane] = 0u;
            __syncwarp();
            int imin = 0x7fffffff, imax = -1, jmin = 0x7fffffff, jmax = -1, kmin = 0x7fffffff, kmax = -1, mmin = 0x7fffffff, mmax = -1;
            for (int a = lane; a < nsp; a += 32)
                if (asg[a] == rep) {
                    atomicOr(&s_bits[w][spt[a] >> 5], 1u << (spt[a] & 31));
                    imin = min(imin, spi[a]); imax = max(imax, spi[a]); jmin = min(jmin, spj[a]); jmax = max(jmax, spj[a]);
                    kmin = min(kmin, spk[a]); kmax = max(kmax, spk[a]); mmin = min(mmin, spm[a]); mmax = max(mmax, spm[a]);
                }
            __syncwarp();
            int ninc = (lane < 8) ? __popc(s_bits[w][lane]) : 0;
            for (int o = 16; o > 0; o >>= 1) {
                ninc += __shfl_xor_sync(FULL, ninc, o);
                imin = min(imin, __shfl_xor_sync(FULL, imin, o)); imax = max(imax, __shfl_xor_sync(FULL, imax, o));
                jmin = min(jmin, __shfl_xor_sync(FULL, jmin, o)); jmax = max(jmax, __shfl_xor_sync(FULL, jmax, o));
                kmin = min(kmin, __shfl_xor_sync(FULL, kmin, o)); kmax = max(kmax, __shfl_xor_sync(FULL, kmax, o));
                mmin = min(mmin, __shfl_xor_sync(FULL, mmin, o)); mmax = max(mmax, __shfl_xor_sync(FULL, mmax, o));
            }
            if ((float)ninc / (float)MD_NSAMPLES < 0.25f) continue;
            const int thr = (int)ceilf((float)ninc * 0.02f);
            int best[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {   // i (leftmost), k (leftmost), j (rightmost), m (rightmost)
                const int *arr = (e == 0) ? spi : (e == 1) ? spk : (e == 2) ? spj : spm;
                const int lo = (e == 0) ? imin : (e == 1) ? kmin : (e == 2) ? jmin : mmin;
                const int hi2 = (e == 0) ? imax : (e == 1) ? kmax : (e == 2) ? jmax : mmax;
                for (int u = lane; u <= hi2 - lo; u += 32) epc[u] = 0;
                __syncwarp();
                for (int a = lane; a < nsp; a += 32) if (asg[a] == rep) atomicAdd(&epc[arr[a] - lo], 1);
                __syncwarp();
                int b = -1;
                if (lane == 0) {
                    if (e < 2) { for (int u = 0; u <= hi2 - lo; u++) if (epc[u] >= thr) { b = lo + u; break; } }
                    else { for (int u = hi2 - lo; u >= 0; u--) if (epc[u] >= thr) { b = lo + u; break; } }
                    if (b < 0) { int am = 0; for (int u = 1; u <= hi2 - lo; u++) if (epc[u] > epc[am]) am = u; b = lo + am; }
                }
                best[e] = __shfl_sync(FULL, b, 0);
                __syncwarp();
            }
            if (best[0] > best[2] || best[1] > best[3]) continue;
            if (nsig < MD_MAXSIG) {
                if (lane == 0) { int *s = s_sig[w][nsig]; s[0] = best[0]; s[1] = best[2]; s[2] = best[1]; s[3] = best[3]; s[4] = ninc; s[5] = corder; }
                nsig++;
            } else oflow = 1;
        }
        __syncwarp();
        // order by start (stable: ties keep Easel's cluster order), drop dominated envelopes, envelope corrections
        if (lane == 0) {
            for (int a = 1; a < nsig; a++) {
                int key[6];
                for (int z = 0; z < 6; z++) key[z] = s_sig[w][a][z];
                int b = a - 1;
                while (b >= 0 && s_sig[w][b][0] > key[0]) { for (int z = 0; z < 6; z++) s_sig[w][b + 1][z] = s_sig[w][b][z]; b--; }
                for (int z = 0; z < 6; z++) s_sig[w][b + 1][z] = key[z];
            }
            unsigned long long dominated = 0ull;
            for (int d = 0; d < nsig; d++)
                for (int d2 = d + 1; d2 < nsig; d2++) {
                    const int *A = s_sig[w][d], *B = s_sig[w][d2];
                    const int nov = min(A[1], B[1]) - max(A[0], B[0]) + 1;
                    if (nov == 0) break;
                    const int n = min(A[1] - A[0] + 1, B[1] - B[0] + 1);
                    if ((float)nov / (float)n >= 0.8f) {
                        if (A[4] > B[4]) dominated |= 1ull << d2; else dominated |= 1ull << d;
                    }
                }
            int nout = 0, fl = oflow ? 4 : 0;
            for (int d = 0; d < nsig; d++) {
                if (dominated >> d & 1ull) continue;
                if (nout < MD_MAXC) {
                    const int i2 = s_sig[w][d][0], j2 = s_sig[w][d][1];
                    float corr = 0.f;
                    for (int p = i2 - R.i0 + 1; p <= j2 - R.i0 + 1; p++) corr += acc[p];
                    out->ci[nout] = i2; out->cj[nout] = j2; out->ccorr[nout] = corr;
                    nout++;
                } else fl = 4;
            }
            out->nclust = nout; out->flags = fl; out->regcorr = regc;
            out->clk[0] = (long long)cx.hdr[2] << 10; out->clk[1] = (long long)cx.hdr[3] << 10; out->clk[2] = clock64() - clk2; out->clk[3] = ((long long)Lr << 32) | (unsigned)M;
        }
        __syncwarp();
    }
}

}  // namespace witch
