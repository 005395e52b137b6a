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
struct MdWork {
    const MdRegion *regions;
    int nregions;
    unsigned *counter;
    char *scratch;
    long long slot_bytes;
    int Lcap, Qcap, Mcap, nsp_cap;
    MdOut *out;
};

struct MdLayout { long long dp, xmx, acc, sp, asg, epc, tkb, total; };
__host__ __device__ inline MdLayout md_layout(int Lcap, int Qcap, int Mcap, int nsp_cap) {
    MdLayout l;
    long long o = 0;
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

constexpr int MD_WARPS = 4;   // warps per CTA

__global__ void __launch_bounds__(MD_WARPS * 32) md_region_kernel(DevEhmm E, DevQueries Qs, MdWork W) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    __shared__ float s_null2[MD_WARPS][32];
    __shared__ int s_dom[MD_WARPS][MD_MAXDOM][4];
    __shared__ int s_sig[MD_WARPS][MD_MAXSIG][6];     // i, j, k, m, count, cluster order
    __shared__ unsigned s_bits[MD_WARPS][8];
    const MdLayout lay = md_layout(W.Lcap, W.Qcap, W.Mcap, W.nsp_cap);
    char *slot = W.scratch + ((long long)blockIdx.x * MD_WARPS + w) * W.slot_bytes;
    float *dp = (float *)(slot + lay.dp);
    float *xmx = (float *)(slot + lay.xmx);     // 8 floats per row: E N J B C SCALE - -
    float *acc = (float *)(slot + lay.acc);
    int *spi = (int *)(slot + lay.sp), *spj = spi + W.nsp_cap, *spk = spj + W.nsp_cap, *spm = spk + W.nsp_cap, *spt = spm + W.nsp_cap;
    int *asg = (int *)(slot + lay.asg);
    int *epc = (int *)(slot + lay.epc);
    const int K = (E.Kp == 29) ? 20 : 4;
    WITCH_DYN_SMEM(float, md_smem);   // per warp: the current row's M and D vectors, [Qcap][4] each
    float *sMv = md_smem + (size_t)w * 8 * W.Qcap, *sDv = sMv + (size_t)4 * W.Qcap;
    int *tkb = (int *)(slot + lay.tkb);   // model nodes of the emitting states of the running domain

    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(W.counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= W.nregions) break;
        const MdRegion R = W.regions[item];
        MdOut *out = W.out + item;
        const int M = E.M[R.h], Q = E.oQ[R.h];
        const float *tfv = E.otfv + E.otoff[R.h];
        const float *rfv = E.orfv + E.oroff[R.h];
        const int L = Qs.len[R.q], Lr = R.j0 - R.i0 + 1;
        const uint8_t *rd = Qs.dsq + Qs.off[R.q] + (R.i0 - 1);   // rd[i-1] = dense code of residue i of the region
        const float pmove = 3.0f / ((float)L + 3.0f), ploop = 1.0f - pmove;
        const size_t RW = (size_t)Q * 12;   // floats per row: [q][M|D|I][4]

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

        const long long clk1 = clock64();
        // ====================== 200 stochastic traces, null2 by trace, sampled domains ======================
        for (int p = lane; p <= Lr + 1; p += 32) acc[p] = 0.f;
        unsigned rng = md_mix3(42u, 87654321u, 12345678u);
        if (rng == 0u) rng = 42u;
        int nsp = 0, oflow = 0;
        __syncwarp();
        enum { tM = 1, tD = 2, tI = 3, tS = 4, tN = 5, tB = 6, tE = 7, tC = 8, tJ = 10 };
        for (int t = 0; t < MD_NSAMPLES; t++) {
            // Lane 0 walks the trace backwards on its own and calls the warp in at three kinds of events: an E state to
            // resolve (a choice among all M/D cells of a row), a finished domain (null2 of its states, per-residue
            // accumulation), the end of the trace. Its running domain: sqfrom..sqto, hfrom..hto, Ld emitting states whose
            // model nodes are listed in tkb[].
            int i = Lr, k = 0, s0 = tC, ndom = 0, hi = Lr;
            int sqto = 0, sqfrom = 0, hto = 0, hfrom = 0, Ld = 0;
            int cq = 0, cr = 0;   // (k-1) % Q and (k-1) / Q of the current node, kept incrementally (no divisions per step)
            for (;;) {
                int ev = 0;
                if (lane == 0) {
                    while (ev == 0) {
                        if (s0 == tE) { ev = 1; break; }
                        if (s0 == tS) { ev = 3; break; }
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
                            const float xb = x0[3], t0 = tp[0], t1 = tp[4], t2 = tp[8], t3 = tp[12];
                            // the path most likely continues on the diagonal: pull the cells of the next steps towards L1
                            if (i >= 5) {
#pragma unroll
                                for (int dstep = 2; dstep <= 4; dstep += 2) {
                                    int qd = q - 1 - dstep, rd2 = r;
                                    if (qd < 0) { qd += Q; rd2--; }
                                    if (rd2 >= 0) {
                                        md_prefetch(dp + (size_t)(i - 1 - dstep) * RW + (size_t)qd * 12 + rd2);
                                        md_prefetch(tfv + (size_t)(qd + 1 < Q ? qd + 1 : 0) * 28);
                                    }
                                }
                            }
                            path[0] = md_mul(xb, t0); path[1] = md_mul(mp, t1); path[2] = md_mul(ip, t2); path[3] = md_mul(dd, t3);
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
                        } else {   // B
                            path[0] = md_mul(pmove, x1[1]);
                            path[1] = md_mul(pmove, x1[2]);
                            s1 = md_choose(rng, path, 2) == 0 ? tN : tJ;
                        }
                        if (s1 == tM || s1 == tI) {   // (3.1b2 counts a residue emitted by I_k in the MATCH cell of node k)
                            if (s1 == tM) { if (sqto == 0) { sqto = i; hto = k; } sqfrom = i; hfrom = k; }
                            tkb[Ld++] = k;
                        }
                        if ((s1 == tN || s1 == tJ || s1 == tC) && s1 == s0) i--;
                        s0 = s1;
                        if (s1 == tB) ev = 2;
                    }
                }
                ev = __shfl_sync(FULL, ev, 0);
                if (ev == 3) break;
                if (ev == 1) {
                    // choice among all M/D cells of row i in striped order, cooperatively: each lane sums a contiguous
                    // range of vectors, a prefix scan locates the lane that crosses the roll, that lane finds the cell
                    i = __shfl_sync(FULL, i, 0);
                    double roll = 0.0;
                    if (lane == 0) roll = md_rand(rng);
                    roll = __shfl_sync(FULL, roll, 0);
                    const float norm = 1.0f / xmx[(size_t)i * 8];
                    const float *row = dp + (size_t)i * RW;
                    const int per = (Q + 31) / 32, q0 = min(lane * per, Q), q1 = min(q0 + per, Q);
                    double part = 0.0;
                    for (int q = q0; q < q1; q++)
#pragma unroll
                        for (int r = 0; r < 8; r++) part += (double)md_mul(row[(size_t)q * 12 + r], norm);
                    double incl = part;
                    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += u; }
                    const unsigned hit = __ballot_sync(FULL, incl > roll);
                    int kk = 1, ss = tM;
                    if (hit != 0u) {
                        const int src = __ffs((int)hit) - 1;
                        if (lane == src) {
                            double sum = incl - part;
                            bool found = false;
                            for (int q = q0; q < q1 && !found; q++)
                                for (int r = 0; r < 8; r++) {
                                    sum += (double)md_mul(row[(size_t)q * 12 + r], norm);
                                    if (sum > roll) { kk = (r & 3) * Q + q + 1; ss = (r < 4) ? tM : tD; found = true; break; }
                                }
                            if (!found) { kk = (q1 - 1) + 1; ss = tM; }
                        }
                        kk = __shfl_sync(FULL, kk, src); ss = __shfl_sync(FULL, ss, src);
                    }
                    // a new domain starts (seen from its end)
                    k = kk; s0 = ss;
                    cq = (kk - 1) % Q; cr = (kk - 1) / Q;
                    sqto = 0; sqfrom = 0; hto = 0; hfrom = 0; Ld = 0;
                    if (ss == tM) { sqto = i; hto = kk; sqfrom = i; hfrom = kk; if (lane == 0) tkb[0] = kk; Ld = 1; }
                    continue;
                }
                // ev == 2: domain sqfrom..sqto complete: null2 odds of its states, accumulated per residue
                sqfrom = __shfl_sync(FULL, sqfrom, 0); sqto = __shfl_sync(FULL, sqto, 0);
                hfrom = __shfl_sync(FULL, hfrom, 0); hto = __shfl_sync(FULL, hto, 0); Ld = __shfl_sync(FULL, Ld, 0);
                __syncwarp();
                {
                    const float nrm = (float)(1.0 / (double)(float)Ld);
                    for (int x = 0; x < K; x++) {
                        double sx = 0.0;
                        for (int z = lane; z < Ld; z += 32) {
                            const int kz = tkb[z] - 1;
                            sx += (double)rfv[((size_t)x * Q + (kz % Q)) * 4 + kz / Q];
                        }
                        for (int o = 16; o > 0; o >>= 1) sx += __shfl_xor_sync(FULL, sx, o);
                        if (lane == 0) s_null2[w][x] = (float)(sx * (double)nrm);
                    }
                }
                __syncwarp();
                for (int p = sqto + 1 + lane; p <= hi; p += 32) acc[p] = md_add(acc[p], 1.0f);
                for (int p = sqfrom + 1 + lane; p <= sqto; p += 32) {
                    const int code = Qs.symrow[rd[p - 1]];
                    float v;
                    if (code < K) v = s_null2[w][code];
                    else {
                        const unsigned mask = md_degen_mask(E.Kp, code);
                        float sg = 0.f; int n = 0;
                        for (int x = 0; x < K; x++) if (mask >> x & 1u) { sg += s_null2[w][x]; n++; }
                        v = n ? sg / (float)n : 1.0f;
                    }
                    acc[p] = md_add(acc[p], v);
                }
                hi = sqfrom;   // (HMMER gives residue sqfrom the neutral 1.0 as well)
                if (ndom < MD_MAXDOM && lane == 0) { s_dom[w][ndom][0] = sqfrom; s_dom[w][ndom][1] = sqto; s_dom[w][ndom][2] = hfrom; s_dom[w][ndom][3] = hto; }
                ndom++;
                sqto = 0; sqfrom = 0; hto = 0; hfrom = 0; Ld = 0;
                __syncwarp();
            }
            for (int p = 1 + lane; p <= hi; p += 32) acc[p] = md_add(acc[p], 1.0f);
            // the trace's domains enter the ensemble in sequence order (they were found last to first)
            if (ndom > MD_MAXDOM) { oflow = 1; ndom = MD_MAXDOM; }
            __syncwarp();
            for (int d = lane; d < ndom; d += 32) {
                const int z = nsp + d;
                if (z < W.nsp_cap) {
                    const int *v = s_dom[w][ndom - 1 - d];
                    spi[z] = v[0] + R.i0 - 1; spj[z] = v[1] + R.i0 - 1; spk[z] = v[2]; spm[z] = v[3]; spt[z] = t;
                }
            }
            nsp += ndom;
            if (nsp > W.nsp_cap) { oflow = 1; nsp = W.nsp_cap; }
            __syncwarp();
        }
        // ln of the mean null2 odds per residue -> acc[]; sum over the region
        float regc = 0.f;
        for (int p = 1 + lane; p <= Lr; p += 32) { const float v = logf(acc[p] / (float)MD_NSAMPLES); acc[p] = v; regc += v; }
        for (int o = 16; o > 0; o >>= 1) regc += __shfl_xor_sync(FULL, regc, o);
        __syncwarp();

        const long long clk2 = clock64();
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
            out->clk[0] = clk1 - clk0; out->clk[1] = clk2 - clk1; out->clk[2] = clock64() - clk2; out->clk[3] = ((long long)Lr << 32) | (unsigned)M;
        }
        __syncwarp();
    }
}

}  // namespace witch
