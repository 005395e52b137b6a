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
    MdOut *out;
};

struct MdLayout { long long hdr, dp, xmx, acc, sp, asg, epc, tkb, total; };
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
    l.tkb = o; o += (long long)(Lcap + 4) * 4;
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
#else
__device__ __forceinline__ void md_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
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
    int *hdr, *spi, *spj, *spk, *spm, *spt, *asg, *epc, *tkb;
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
    c.tkb = (int *)(slot + lay.tkb);
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
            __syncwarp();
        }

        if (lane == 0) { cx.hdr[2] = (int)((clock64() - clk0) >> 10); }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 2: the 200 stochastic traces of a region, its position-specific null2 and its sampled domains. The walk is
// sequential by nature (one random-number stream per region), so ONE THREAD per region: the parallelism is across the
// regions of the batch (ordered by size, so the lanes of a warp walk matrices of similar shape).
__global__ void __launch_bounds__(128) md_trace_kernel(DevEhmm E, DevQueries Qs, MdWork W) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W.end - W.begin) return;
    const MdCtx cx = md_ctx(E, Qs, W, j);
    const int Q = cx.Q, Lr = cx.Lr, K = cx.K;
    const float *tfv = cx.tfv, *rfv = cx.rfv, *dp = cx.dp, *xmx = cx.xmx;
    const uint8_t *rd = cx.rd;
    const float pmove = cx.pmove, ploop = cx.ploop;
    const size_t RW = cx.RW;
    float *acc = cx.acc;
    const long long clk0 = clock64();
    enum { tM = 1, tD = 2, tI = 3, tS = 4, tN = 5, tB = 6, tE = 7, tC = 8, tJ = 10 };
    for (int p = 0; p <= Lr + 1; p++) acc[p] = 0.f;
    unsigned rng = md_mix3(42u, 87654321u, 12345678u);
    if (rng == 0u) rng = 42u;
    int nsp = 0, oflow = 0;
    for (int t = 0; t < MD_NSAMPLES; t++) {
        int i = Lr, k = 0, s0 = tC, ndom = 0, hi = Lr;
        int sqto = 0, sqfrom = 0, hto = 0, hfrom = 0, Ld = 0;
        int cq = 0, cr = 0;   // (k-1) % Q and (k-1) / Q of the current node, kept incrementally
        double sums[20];      // sum over the running domain's emitting states of the match odds of every residue
        while (s0 != tS) {
            int s1;
            const float *x1 = xmx + (size_t)i * 8, *x0 = xmx + (size_t)(i > 0 ? i - 1 : 0) * 8;
            float path[4];
            if (s0 == tM) {
                k--;
                const int q = cq, r = cr;   // = k % Q, k / Q of the decremented k
                if (--cq < 0) { cq += Q; cr--; }
                const float *tp = tfv + (size_t)q * 28 + r;
                const float *pr = dp + (size_t)(i - 1) * RW;
                float mp = 0.f, dd = 0.f, ip = 0.f;
                if (q > 0) { const float *v = pr + (size_t)(q - 1) * 12 + r; mp = v[0]; dd = v[4]; ip = v[8]; }
                else if (r > 0) { const float *v = pr + (size_t)(Q - 1) * 12 + r - 1; mp = v[0]; dd = v[4]; ip = v[8]; }
                path[0] = md_mul(x0[3], tp[0]); path[1] = md_mul(mp, tp[4]); path[2] = md_mul(ip, tp[8]); path[3] = md_mul(dd, tp[12]);
                const int c = md_choose(rng, path, 4);
                s1 = (c == 0) ? tB : (c == 1) ? tM : (c == 2) ? tI : tD;
                i--;
            } else if (s0 == tD) {
                k--;
                const int q = cq, r = cr;
                if (--cq < 0) { cq += Q; cr--; }
                const float *crow = dp + (size_t)i * RW;
                float mp = 0.f, dd = 0.f, tmd = 0.f, tdd = 0.f;
                if (q > 0) {
                    mp = crow[(size_t)(q - 1) * 12 + r]; dd = crow[(size_t)(q - 1) * 12 + 4 + r];
                    tmd = tfv[(size_t)(q - 1) * 28 + 16 + r]; tdd = tfv[(size_t)Q * 28 + (size_t)(q - 1) * 4 + r];
                } else if (r > 0) {
                    mp = crow[(size_t)(Q - 1) * 12 + r - 1]; dd = crow[(size_t)(Q - 1) * 12 + 4 + r - 1];
                    tmd = tfv[(size_t)(Q - 1) * 28 + 16 + r - 1]; tdd = tfv[(size_t)Q * 28 + (size_t)(Q - 1) * 4 + r - 1];
                }
                path[0] = md_mul(mp, tmd); path[1] = md_mul(dd, tdd);
                s1 = md_choose(rng, path, 2) == 0 ? tM : tD;
            } else if (s0 == tI) {
                const int q = cq, r = cr;
                const float *pr = dp + (size_t)(i - 1) * RW + (size_t)q * 12 + r;
                path[0] = md_mul(pr[0], tfv[(size_t)q * 28 + 20 + r]);
                path[1] = md_mul(pr[8], tfv[(size_t)q * 28 + 24 + r]);
                s1 = md_choose(rng, path, 2) == 0 ? tM : tI;
                i--;
            } else if (s0 == tN) {
                s1 = (i == 0) ? tS : tN;
            } else if (s0 == tC) {
                path[0] = md_mul(ploop, x0[4]);
                path[1] = md_mul(md_mul(0.5f, x1[0]), x1[5]);
                s1 = md_choose(rng, path, 2) == 0 ? tC : tE;
            } else if (s0 == tJ) {
                path[0] = md_mul(ploop, x0[2]);
                path[1] = md_mul(md_mul(0.5f, x1[0]), x1[5]);
                s1 = md_choose(rng, path, 2) == 0 ? tJ : tE;
            } else if (s0 == tB) {
                path[0] = md_mul(pmove, x1[1]);
                path[1] = md_mul(pmove, x1[2]);
                s1 = md_choose(rng, path, 2) == 0 ? tN : tJ;
            } else {   // E: one roll against the running sum over all M/D cells of row i, in HMMER's striped order
                const double roll = md_rand(rng);
                const float norm = 1.0f / x1[0];
                const float *row = dp + (size_t)i * RW;
                double sum = 0.0;
                int kk = 1;
                s1 = tM;
                bool found = false;
                for (int q = 0; q < Q && !found; q++) {
                    const float4 m4 = *reinterpret_cast<const float4 *>(row + (size_t)q * 12), d4 = *reinterpret_cast<const float4 *>(row + (size_t)q * 12 + 4);
                    const float v8[8] = {m4.x, m4.y, m4.z, m4.w, d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        sum += (double)md_mul(v8[r], norm);
                        if (!found && sum > roll) { kk = (r & 3) * Q + q + 1; s1 = (r < 4) ? tM : tD; found = true; }
                    }
                }
                k = kk;
                cq = (kk - 1) % Q; cr = (kk - 1) / Q;
                sqto = 0; sqfrom = 0; hto = 0; hfrom = 0; Ld = 0;   // a new domain starts (seen from its end)
                for (int x = 0; x < K; x++) sums[x] = 0.0;
            }
            if (s1 == tM || s1 == tI) {   // (3.1b2 counts a residue emitted by I_k in the MATCH cell of node k)
                if (s1 == tM) { if (sqto == 0) { sqto = i; hto = k; } sqfrom = i; hfrom = k; }
                Ld++;
                for (int x = 0; x < K; x++) sums[x] += (double)rfv[((size_t)x * Q + cq) * 4 + cr];
            } else if (s1 == tB) {
                // domain sqfrom..sqto complete: null2 odds of its states, accumulated per residue
                float null2[20];
                const float nrm = (float)(1.0 / (double)(float)Ld);
                for (int x = 0; x < K; x++) null2[x] = (float)(sums[x] * (double)nrm);
                for (int p = sqto + 1; p <= hi; p++) acc[p] = md_add(acc[p], 1.0f);
                for (int p = sqfrom + 1; p <= sqto; p++) {
                    const int code = Qs.symrow[rd[p - 1]];
                    float v;
                    if (code < K) v = null2[code];
                    else {
                        const unsigned mask = md_degen_mask(E.Kp, code);
                        float sg = 0.f; int n = 0;
                        for (int x = 0; x < K; x++) if (mask >> x & 1u) { sg += null2[x]; n++; }
                        v = n ? sg / (float)n : 1.0f;
                    }
                    acc[p] = md_add(acc[p], v);
                }
                hi = sqfrom;   // (HMMER gives residue sqfrom the neutral 1.0 as well)
                if (ndom < MD_MAXDOM && nsp + ndom < W.nsp_cap) {
                    const int z = nsp + ndom;
                    cx.spi[z] = sqfrom + cx.R.i0 - 1; cx.spj[z] = sqto + cx.R.i0 - 1; cx.spk[z] = hfrom; cx.spm[z] = hto; cx.spt[z] = t;
                    ndom++;
                } else oflow = 1;
            }
            if ((s1 == tN || s1 == tJ || s1 == tC) && s1 == s0) i--;
            s0 = s1;
        }
        for (int p = 1; p <= hi; p++) acc[p] = md_add(acc[p], 1.0f);
        // the trace's domains enter the ensemble in sequence order (they were found last to first): reverse the segment
        for (int a = nsp, b = nsp + ndom - 1; a < b; a++, b--) {
            int tmp;
            tmp = cx.spi[a]; cx.spi[a] = cx.spi[b]; cx.spi[b] = tmp; tmp = cx.spj[a]; cx.spj[a] = cx.spj[b]; cx.spj[b] = tmp;
            tmp = cx.spk[a]; cx.spk[a] = cx.spk[b]; cx.spk[b] = tmp; tmp = cx.spm[a]; cx.spm[a] = cx.spm[b]; cx.spm[b] = tmp;
        }
        nsp += ndom;
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
