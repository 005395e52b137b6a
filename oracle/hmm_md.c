/*
 * hmm_md.c -- TEST INFRASTRUCTURE ONLY (part of oracle/liboracle.so; never linked or called by the product path).
 *
 * CPU restatement of the branch of HMMER 3.1b2's domain definition that hmmsearch takes when a region fails the
 * single-domain test: p7_domaindef_ByPosteriorHeuristics -> region_trace_ensemble (200 stochastic tracebacks over
 * the region's multihit Forward matrix, position-specific null2 from the traces, single-linkage clustering of the
 * sampled domains into envelopes). Reference call site: witch_msa/gcmm/algorithm.py:526-532 (`hmmsearch --max`).
 * HMMER is shipped in the reference only as x86-64 binaries (witch_msa/tools/magus/tools/hmmer/hmmsearch, not
 * stripped); what is restated here was checked against that binary's own code (objdump) and is PINNED by
 * tests/golden/md_golden.json: cluster lists captured from the binary with tools/hmmer_probe (qsort interposer)
 * and its printed per-domain / per-sequence scores.
 *
 * Why float and this exact operation order: the sampled traces depend on comparisons `roll < cumulative
 * probability`, so the region Forward matrix is computed the way HMMER's SSE code computes it (4-way striped
 * vectors, FP32 mul/add in the same order, sparse rescaling at xE > 1e4, specials evaluated in double and rounded
 * to float per statement, Easel's own polynomial expf for the profile tables, the "fast" LCG random number
 * generator re-seeded with 42 for every region). With that, the 200 traces reproduce hmmsearch's, except where a
 * roll falls within an ulp-level difference of a cumulative probability.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction: SSE mulps/addps semantics).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { tBM = 0, tMM, tIM, tDM, tMD, tMI, tII, NT7 };   /* order of the striped transition vectors (p7o_tsc_e) */
enum { xE = 0, xN, xJ, xB, xC, xSCALE, NXC };            /* p7X_* special cells of a row */
enum { hMM = 0, hMI, hMD, hIM, hII, hDM, hDD };           /* column order of the HMM file / core model */
enum { sM = 1, sD = 2, sI = 3, sS = 4, sN = 5, sB = 6, sE = 7, sC = 8, sT = 9, sJ = 10 };  /* p7t_statetype_e */

typedef struct {
    int M, Q, K, Kp;
    float *tfv; /* [(7*Q + Q)][4]: BM,MM,IM,DM,MD,MI,II per q, then Q DD vectors */
    float *rfv; /* [Kp][Q][4] match emission odds */
    float *msc; /* [Kp][M+1] match scores (for null2: odds = rfv) */
} oprof_t;

/* Easel's esl_sse_expf (Cephes polynomial), one lane: constants read from the binary's .rodata */
static float sse_expf1(float x) {
    const float maxlogf = 88.3762588501f, minlogf = -88.3762588501f;
    union { float f; uint32_t u; } c;
    const float x0 = x;
    float fx = x * 1.44269502f;
    fx = fx + 0.5f;
    int k = (int)fx;                 /* cvttps2dq: truncation */
    float tmp = (float)k;
    if (fx < tmp) tmp = tmp - 1.0f;  /* floor */
    fx = tmp;
    k = (int)fx;
    {
        const float a = fx * 0.693359375f, z0 = fx * -2.12194440e-4f;
        x = x - a;
        x = x - z0;
    }
    const float z = x * x;
    c.u = 961571175u;  float y = c.f * x;            /* 1.9875691500E-4 */
    c.u = 985088974u;  y = y + c.f; y = y * x;       /* 1.3981999507E-3 */
    c.u = 1007192328u; y = y + c.f; y = y * x;       /* 8.3334519073E-3 */
    c.u = 1026206145u; y = y + c.f; y = y * x;       /* 4.1665795894E-2 */
    c.u = 1042983594u; y = y + c.f; y = y * x;       /* 1.6666665459E-1 */
    y = y + 0.5f;
    y = y * z;
    y = y + x;
    y = y + 1.0f;
    c.u = (uint32_t)(k + 127) << 23;
    y = y * c.f;
    if (x0 > maxlogf) return INFINITY;
    if (x0 <= minlogf) return 0.0f;   /* also -inf */
    return y;
}

void md_oprofile_free(oprof_t *om) {
    if (!om) return;
    free(om->tfv); free(om->rfv); free(om->msc); free(om);
}

/*
 * raw_t [(M+1)*7], raw_mat [(M+1)*K]: the numbers of the HMMER3/f text file (-ln p; +inf for '*'), node 0..M.
 * Follows read_asc30hmm (p = expf(-x)), p7_hmm_CalculateOccupancy, p7_ProfileConfig (local mode) and
 * p7_oprofile_Convert's fb_conversion.
 */
oprof_t *md_oprofile_create(int M, int K, int Kp, const double *raw_t, const double *raw_mat, const float *bg,
                            const int *degen_n, const int *degen_set) {
    oprof_t *om = (oprof_t *)calloc(1, sizeof(oprof_t));
    const int Q = (M - 1) / 4 + 1 < 2 ? 2 : (M - 1) / 4 + 1;
    om->M = M; om->Q = Q; om->K = K; om->Kp = Kp;
    float *t = (float *)calloc((size_t)(M + 1) * 7, sizeof(float));
    float *mat = (float *)calloc((size_t)(M + 1) * K, sizeof(float));
    for (int k = 0; k <= M; k++) {
        for (int x = 0; x < 7; x++) {
            const double v = raw_t[(size_t)k * 7 + x];
            t[k * 7 + x] = isinf(v) ? 0.0f : expf((float)(-1.0 * v));
        }
        for (int x = 0; x < K && k >= 1; x++) {
            const double v = raw_mat[(size_t)k * K + x];
            mat[(size_t)k * K + x] = isinf(v) ? 0.0f : expf((float)(-1.0 * v));
        }
    }
    /* occupancy */
    float *occ = (float *)calloc(M + 2, sizeof(float));
    occ[0] = 0.f;
    occ[1] = t[0 * 7 + hMI] + t[0 * 7 + hMM];
    for (int k = 2; k <= M; k++) {
        const float a = occ[k - 1] * (t[(k - 1) * 7 + hMM] + t[(k - 1) * 7 + hMI]);
        occ[k] = (float)((double)a + (1.0 - (double)occ[k - 1]) * (double)t[(k - 1) * 7 + hDM]);
    }
    float Z = 0.f;
    for (int k = 1; k <= M; k++) Z += occ[k] * (float)(M - k + 1);
    /* scores (nats) */
    float *tsc = (float *)malloc((size_t)(M + 1) * 8 * sizeof(float));   /* [k][MM,MI,MD,IM,II,DM,DD,BM] */
    for (int k = 0; k <= M; k++)
        for (int x = 0; x < 8; x++) tsc[k * 8 + x] = -INFINITY;
    for (int k = 1; k <= M; k++) tsc[(k - 1) * 8 + 7] = (float)log((double)(occ[k] / Z));
    for (int k = 1; k < M; k++)
        for (int x = 0; x < 7; x++) tsc[k * 8 + x] = (float)log((double)t[k * 7 + x]);
    om->msc = (float *)malloc((size_t)Kp * (M + 1) * sizeof(float));
    for (int x = 0; x < Kp; x++) om->msc[(size_t)x * (M + 1)] = -INFINITY;
    for (int k = 1; k <= M; k++) {
        float sc[64];
        for (int x = 0; x < Kp; x++) sc[x] = -INFINITY;
        for (int x = 0; x < K; x++) sc[x] = (float)log((double)mat[(size_t)k * K + x] / bg[x]);
        for (int x = K + 1; x <= Kp - 3; x++) {   /* esl_abc_FExpectScVec */
            float result = 0.f, denom = 0.f;
            for (int a = 0; a < degen_n[x]; a++) {
                const int i = degen_set[x * K + a];
                result += sc[i] * bg[i]; denom += bg[i];
            }
            sc[x] = degen_n[x] > 0 ? result / denom : -INFINITY;
        }
        for (int x = 0; x < Kp; x++) om->msc[(size_t)x * (M + 1) + k] = sc[x];
    }
    /* striped probability-space tables */
    om->rfv = (float *)calloc((size_t)Kp * Q * 4, sizeof(float));
    for (int x = 0; x < Kp; x++)
        for (int q = 0; q < Q; q++)
            for (int z = 0; z < 4; z++) {
                const int k = q + 1 + z * Q;
                om->rfv[((size_t)x * Q + q) * 4 + z] = sse_expf1(k <= M ? om->msc[(size_t)x * (M + 1) + k] : -INFINITY);
            }
    om->tfv = (float *)calloc((size_t)8 * Q * 4, sizeof(float));
    static const int src[7] = { 7, hMM, hIM, hDM, hMD, hMI, hII };
    for (int q = 0; q < Q; q++) {
        const int k = q + 1;
        for (int tt = 0; tt < 7; tt++) {
            const int kb0 = (tt <= tDM) ? k - 1 : k;
            for (int z = 0; z < 4; z++) {
                const int kb = kb0 + z * Q;
                om->tfv[((size_t)(7 * q + tt)) * 4 + z] = sse_expf1(kb < M ? tsc[kb * 8 + src[tt]] : -INFINITY);
            }
        }
        for (int z = 0; z < 4; z++) {
            const int kb = k + z * Q;
            om->tfv[((size_t)(7 * Q + q)) * 4 + z] = sse_expf1(kb < M ? tsc[kb * 8 + hDD] : -INFINITY);
        }
    }
    free(t); free(mat); free(occ); free(tsc);
    return om;
}

/* ---- Easel's "fast" generator (esl_randomness_CreateFast): Knuth LCG seeded through Jenkins' mix3 ---- */
typedef struct { uint32_t x; } rng_t;
static uint32_t mix3(uint32_t a, uint32_t b, uint32_t c) {
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
static void rng_init(rng_t *r, uint32_t seed) { r->x = mix3(seed, 87654321u, 12345678u); if (r->x == 0) r->x = 42; }
static double rng_next(rng_t *r) { r->x = r->x * 69069u + 1u; return (double)r->x * 2.3283064365386963e-10; }

/* esl_vec_FNorm (n < 8: sequential float sum, float divisions) + esl_rnd_FChoose (double running sum / double norm) */
static int fchoose(rng_t *r, float *p, int n) {
    float s = 0.f;
    for (int i = 0; i < n; i++) s += p[i];
    if (s != 0.f) for (int i = 0; i < n; i++) p[i] = p[i] / s;
    else for (int i = 0; i < n; i++) p[i] = 1.0f / (float)n;
    const double roll = rng_next(r);
    double norm = 0.0, sum = 0.0;
    for (int i = 0; i < n; i++) norm += (double)p[i];
    for (int i = 0; i < n; i++) {
        sum += (double)p[i];
        if (sum / norm > roll) return i;
    }
    return n - 1;   /* (esl_fatal in Easel; unreachable for normalised p) */
}

typedef struct {
    int L, Q;
    float *dp;   /* [(L+1)][Q][3][4]: M, D, I vectors per q */
    float *xmx;  /* [(L+1)][6] */
} fmx_t;
#define DPV(f, i, q, s) ((f)->dp + ((((size_t)(i) * (f)->Q + (q)) * 3 + (s)) * 4))

/* p7_Forward (impl_sse/fwdback.c forward_engine, do_full) on dsq[0..L-1], multihit, length model Lm */
static void md_forward(const oprof_t *om, const uint8_t *dsq, int L, int Lm, fmx_t *fx) {
    const int Q = om->Q, M = om->M;
    fx->L = L; fx->Q = Q;
    fx->dp = (float *)calloc((size_t)(L + 1) * Q * 12, sizeof(float));
    fx->xmx = (float *)calloc((size_t)(L + 1) * NXC, sizeof(float));
    const float nj = 1.0f;
    const float pmove = (2.0f + nj) / ((float)Lm + 2.0f + nj), ploop = 1.0f - pmove;
    const float eLoop = 0.5f, eMove = 0.5f;
    float fN = 1.0f, fB = pmove, fJ = 0.f, fC = 0.f, fE = 0.f;
    fx->xmx[xE] = 0.f; fx->xmx[xN] = 1.f; fx->xmx[xJ] = 0.f; fx->xmx[xB] = pmove; fx->xmx[xC] = 0.f; fx->xmx[xSCALE] = 1.f;
    float *dcv = (float *)malloc(4 * sizeof(float));
    for (int i = 1; i <= L; i++) {
        const float *rp = om->rfv + (size_t)dsq[i - 1] * Q * 4;
        float xEv[4] = { 0, 0, 0, 0 }, mpv[4], dpv[4], ipv[4], sv[4], dc[4] = { 0, 0, 0, 0 };
        /* rightshift of the last vectors of row i-1 */
        for (int z = 3; z >= 1; z--) {
            mpv[z] = DPV(fx, i - 1, Q - 1, 0)[z - 1]; dpv[z] = DPV(fx, i - 1, Q - 1, 1)[z - 1]; ipv[z] = DPV(fx, i - 1, Q - 1, 2)[z - 1];
        }
        mpv[0] = dpv[0] = ipv[0] = 0.f;
        for (int q = 0; q < Q; q++) {
            const float *tp = om->tfv + (size_t)7 * q * 4;
            float *cM = DPV(fx, i, q, 0), *cD = DPV(fx, i, q, 1), *cI = DPV(fx, i, q, 2);
            const float *pM = DPV(fx, i - 1, q, 0), *pD = DPV(fx, i - 1, q, 1), *pI = DPV(fx, i - 1, q, 2);
            for (int z = 0; z < 4; z++) {
                float s = fB * tp[tBM * 4 + z];
                s = s + mpv[z] * tp[tMM * 4 + z];
                s = s + ipv[z] * tp[tIM * 4 + z];
                s = s + dpv[z] * tp[tDM * 4 + z];
                s = s * rp[q * 4 + z];
                sv[z] = s;
                xEv[z] = xEv[z] + s;
            }
            for (int z = 0; z < 4; z++) {
                mpv[z] = pM[z]; dpv[z] = pD[z]; ipv[z] = pI[z];
                cM[z] = sv[z];
                cD[z] = dc[z];
                dc[z] = sv[z] * tp[tMD * 4 + z];
                cI[z] = mpv[z] * tp[tMI * 4 + z] + ipv[z] * tp[tII * 4 + z];
            }
        }
        /* DD paths */
        {
            const float *td = om->tfv + (size_t)7 * Q * 4;
            float d[4] = { 0.f, dc[0], dc[1], dc[2] };
            for (int z = 0; z < 4; z++) DPV(fx, i, 0, 1)[z] = 0.f;
            for (int q = 0; q < Q; q++) {
                float *cD = DPV(fx, i, q, 1);
                for (int z = 0; z < 4; z++) { cD[z] = d[z] + cD[z]; d[z] = cD[z] * td[q * 4 + z]; }
            }
            if (M < 100) {
                for (int j = 1; j < 4; j++) {
                    float e[4] = { 0.f, d[0], d[1], d[2] };
                    memcpy(d, e, sizeof(e));
                    for (int q = 0; q < Q; q++) {
                        float *cD = DPV(fx, i, q, 1);
                        for (int z = 0; z < 4; z++) { cD[z] = d[z] + cD[z]; d[z] = d[z] * td[q * 4 + z]; }
                    }
                }
            } else {
                for (int j = 1; j < 4; j++) {
                    float e[4] = { 0.f, d[0], d[1], d[2] };
                    memcpy(d, e, sizeof(e));
                    int changed = 0;
                    for (int q = 0; q < Q; q++) {
                        float *cD = DPV(fx, i, q, 1);
                        for (int z = 0; z < 4; z++) {
                            const float s = d[z] + cD[z];
                            if (cD[z] < s) changed = 1;
                            cD[z] = s;
                            d[z] = d[z] * td[q * 4 + z];
                        }
                    }
                    if (!changed) break;
                }
            }
            for (int q = 0; q < Q; q++) {
                const float *cD = DPV(fx, i, q, 1);
                for (int z = 0; z < 4; z++) xEv[z] = cD[z] + xEv[z];
            }
        }
        {   /* horizontal sum: (a0+a1) + (a2+a3) */
            const float s01 = xEv[0] + xEv[1], s23 = xEv[2] + xEv[3];
            fE = s01 + s23;
        }
        /* specials: evaluated in double, rounded to float per statement (what the binary does) */
        fJ = (float)((double)ploop * (double)fJ + (double)eLoop * (double)fE);
        fN = (float)((double)ploop * (double)fN);
        fC = (float)((double)ploop * (double)fC + (double)eMove * (double)fE);
        fB = (float)((double)pmove * (double)fN + (double)pmove * (double)fJ);
        float scale = 1.0f;
        if ((double)fE > 1.0e4) {
            const double e = (double)fE;
            fN = (float)((double)fN / e); fC = (float)((double)fC / e); fJ = (float)((double)fJ / e); fB = (float)((double)fB / e);
            const float inv = (float)(1.0 / e);
            for (int q = 0; q < Q; q++)
                for (int s = 0; s < 3; s++) {
                    float *v = DPV(fx, i, q, s);
                    for (int z = 0; z < 4; z++) v[z] = v[z] * inv;
                }
            scale = fE;
            fE = 1.0f;
        }
        float *x = fx->xmx + (size_t)i * NXC;
        x[xE] = fE; x[xN] = fN; x[xJ] = fJ; x[xB] = fB; x[xC] = fC; x[xSCALE] = scale;
    }
    free(dcv);
}

/* one sampled domain of one trace */
typedef struct { int idx, i, j, k, m; float prob; } spc_t;

typedef struct {
    int n, cap;
    spc_t *v;
} splist_t;
static void sp_add(splist_t *s, int idx, int i, int j, int k, int m) {
    if (s->n == s->cap) { s->cap = s->cap ? 2 * s->cap : 256; s->v = (spc_t *)realloc(s->v, (size_t)s->cap * sizeof(spc_t)); }
    spc_t c = { idx, i, j, k, m, 0.f };
    s->v[s->n++] = c;
}

/* link test of p7_spensemble.c:link_spsamples (min_overlap 0.8 of the smaller, max_diagdiff 4; the model-overlap
 * numerator has no "+1" in 3.1b2; EITHER the start or the end diagonals being close links the two) */
static int sp_link(const spc_t *a, const spc_t *b) {
    int nov = (a->j < b->j ? a->j : b->j) - (a->i > b->i ? a->i : b->i) + 1;
    int n = (a->j - a->i < b->j - b->i ? a->j - a->i : b->j - b->i) + 1;
    if ((float)nov / (float)n < 0.8f) return 0;
    nov = (a->m < b->m ? a->m : b->m) - (a->k > b->k ? a->k : b->k);
    n = (a->m - a->k < b->m - b->k ? a->m - a->k : b->m - b->k) + 1;
    if ((float)nov / (float)n < 0.8f) return 0;
    if (abs((a->i - a->k) - (b->i - b->k)) <= 4) return 1;
    if (abs((a->j - a->m) - (b->j - b->m)) <= 4) return 1;
    return 0;
}

static int cmp_spc_i(const void *a, const void *b) {
    const spc_t *x = (const spc_t *)a, *y = (const spc_t *)b;
    return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}

/*
 * The multi-domain branch for region ireg..jreg (1-based) of dsq[0..L-1].
 * Out: n2sc[ireg..jreg] (array indexed 1..L) = ln of the mean null2 odds over the traces; clusters (ienv, jenv,
 * kenv, menv, count) sorted by start, dominated ones removed. Returns the number of envelopes (<= max_out).
 * nsamples = 200, seed = 42 in hmmsearch.
 */
int md_region(const oprof_t *om, const int *degen_n, const int *degen_set, const uint8_t *dsq, int L, int ireg, int jreg,
              int nsamples, uint32_t seed, double *n2sc, int *out_env /* [max_out][5] */, int max_out, int *out_nsampled,
              int *out_sig /* optional [max_sig][5]: every significant cluster before the dominated filter */, int max_sig, int *n_sig) {
    const int Lr = jreg - ireg + 1, Q = om->Q, M = om->M, K = om->K, Kp = om->Kp;
    const uint8_t *rd = dsq + (ireg - 1);   /* rd[i-1] = residue i of the region */
    fmx_t fx;
    md_forward(om, rd, Lr, L, &fx);
    const float nj = 1.0f;
    const float pmove = (2.0f + nj) / ((float)L + 2.0f + nj), ploop = 1.0f - pmove;
    float *acc = (float *)calloc(Lr + 2, sizeof(float));   /* n2sc accumulators, 1..Lr */
    rng_t rng;
    rng_init(&rng, seed);
    splist_t sp = { 0, 0, NULL };
    /* trace storage (backwards) */
    int cap = 2 * (Lr + M) + 16;
    int8_t *st = (int8_t *)malloc(cap);
    int *tk = (int *)malloc(cap * sizeof(int)), *ti = (int *)malloc(cap * sizeof(int));
    float *cntM = (float *)calloc(M + 2, sizeof(float)), *cntI = (float *)calloc(M + 2, sizeof(float));
    for (int t = 0; t < nsamples; t++) {
        int n = 0, i = Lr, k = 0, s0 = sC, s1;
        st[n] = sT; tk[n] = 0; ti[n] = i; n++;
        st[n] = sC; tk[n] = 0; ti[n] = i; n++;
        while (s0 != sS) {
            const float *x1 = fx.xmx + (size_t)i * NXC, *x0 = i > 0 ? fx.xmx + (size_t)(i - 1) * NXC : x1;
            float path[4];
            switch (s0) {
            case sM: {
                k--;   /* cell M(i,k_old): predecessors on row i-1 at column k_old-1 = k */
                const int q = k % Q, r = k / Q;   /* (k_old-1) % Q, (k_old-1) / Q */
                const float *tp = om->tfv + (size_t)7 * q * 4;
                float mp, dp_, ip;
                if (q > 0) { mp = DPV(&fx, i - 1, q - 1, 0)[r]; dp_ = DPV(&fx, i - 1, q - 1, 1)[r]; ip = DPV(&fx, i - 1, q - 1, 2)[r]; }
                else if (r > 0) { mp = DPV(&fx, i - 1, Q - 1, 0)[r - 1]; dp_ = DPV(&fx, i - 1, Q - 1, 1)[r - 1]; ip = DPV(&fx, i - 1, Q - 1, 2)[r - 1]; }
                else { mp = dp_ = ip = 0.f; }
                path[0] = x0[xB] * tp[tBM * 4 + r];
                path[1] = mp * tp[tMM * 4 + r];
                path[2] = ip * tp[tIM * 4 + r];
                path[3] = dp_ * tp[tDM * 4 + r];
                static const int stt[4] = { sB, sM, sI, sD };
                s1 = stt[fchoose(&rng, path, 4)];
                i--;
                break;
            }
            case sD: {
                k--;
                const int q = k % Q, r = k / Q;
                float mp, dp_, tmd, tdd;
                if (q > 0) {
                    mp = DPV(&fx, i, q - 1, 0)[r]; dp_ = DPV(&fx, i, q - 1, 1)[r];
                    tmd = om->tfv[((size_t)7 * (q - 1) + tMD) * 4 + r]; tdd = om->tfv[((size_t)7 * Q + q - 1) * 4 + r];
                } else if (r > 0) {
                    mp = DPV(&fx, i, Q - 1, 0)[r - 1]; dp_ = DPV(&fx, i, Q - 1, 1)[r - 1];
                    tmd = om->tfv[((size_t)7 * (Q - 1) + tMD) * 4 + r - 1]; tdd = om->tfv[((size_t)7 * Q + Q - 1) * 4 + r - 1];
                } else { mp = dp_ = tmd = tdd = 0.f; }
                path[0] = mp * tmd;
                path[1] = dp_ * tdd;
                s1 = fchoose(&rng, path, 2) == 0 ? sM : sD;
                break;
            }
            case sI: {
                const int q = (k - 1) % Q, r = (k - 1) / Q;
                path[0] = DPV(&fx, i - 1, q, 0)[r] * om->tfv[((size_t)7 * q + tMI) * 4 + r];
                path[1] = DPV(&fx, i - 1, q, 2)[r] * om->tfv[((size_t)7 * q + tII) * 4 + r];
                s1 = fchoose(&rng, path, 2) == 0 ? sM : sI;
                i--;
                break;
            }
            case sN: s1 = (i == 0) ? sS : sN; break;
            case sC:
                path[0] = ploop * x0[xC];
                path[1] = (0.5f * x1[xE]) * x1[xSCALE];
                s1 = fchoose(&rng, path, 2) == 0 ? sC : sE;
                break;
            case sJ:
                path[0] = ploop * x0[xJ];
                path[1] = (0.5f * x1[xE]) * x1[xSCALE];
                s1 = fchoose(&rng, path, 2) == 0 ? sJ : sE;
                break;
            case sE: {
                const double roll = rng_next(&rng);
                const float norm = 1.0f / x1[xE];
                double sum = 0.0;
                s1 = -1;
                for (int rep = 0; rep < 2 && s1 < 0; rep++)
                    for (int q = 0; q < Q && s1 < 0; q++) {
                        const float *vm = DPV(&fx, i, q, 0), *vd = DPV(&fx, i, q, 1);
                        for (int r = 0; r < 4; r++) { sum += (double)(vm[r] * norm); if (sum > roll) { k = r * Q + q + 1; s1 = sM; break; } }
                        if (s1 >= 0) break;
                        for (int r = 0; r < 4; r++) { sum += (double)(vd[r] * norm); if (sum > roll) { k = r * Q + q + 1; s1 = sD; break; } }
                    }
                if (s1 < 0) { s1 = sM; k = 1; }   /* (HMMER would loop / raise; unreachable for a valid matrix) */
                break;
            }
            case sB:
                path[0] = pmove * x1[xN];
                path[1] = pmove * x1[xJ];
                s1 = fchoose(&rng, path, 2) == 0 ? sN : sJ;
                break;
            default: s1 = sS; break;
            }
            if (n + 2 >= cap) { cap *= 2; st = (int8_t *)realloc(st, cap); tk = (int *)realloc(tk, cap * sizeof(int)); ti = (int *)realloc(ti, cap * sizeof(int)); }
            st[n] = (int8_t)s1; tk[n] = k; ti[n] = i; n++;
            if ((s1 == sN || s1 == sJ || s1 == sC) && s1 == s0) i--;
            s0 = s1;
        }
        /* walk the (reversed) trace forward: domains B..E */
        int pos = 1;
        int z = n - 1;
        while (z >= 0) {
            if (st[z] != sB) { z--; continue; }
            int sqfrom = 0, sqto = 0, hfrom = 0, hto = 0, Ld = 0;
            int zz = z - 1;
            memset(cntM, 0, (M + 2) * sizeof(float)); memset(cntI, 0, (M + 2) * sizeof(float));
            for (; zz >= 0 && st[zz] != sE; zz--) {
                if (st[zz] == sM) {
                    if (sqfrom == 0) { sqfrom = ti[zz]; hfrom = tk[zz]; }
                    sqto = ti[zz]; hto = tk[zz];
                    cntM[tk[zz]] += 1.0f; Ld++;
                } else if (st[zz] == sI) { cntM[tk[zz]] += 1.0f; Ld++; }   /* 3.1b2 (binary checked): inserts are counted in the MATCH cell of their node */
            }
            z = zz;
            sp_add(&sp, t, sqfrom + ireg - 1, sqto + ireg - 1, hfrom, hto);
            /* p7_Null2_ByTrace: null2[x] = sum_k fM(k) e_k(x) + sum_k fI(k), striped accumulation order */
            float null2[64];
            const float norm = (float)(1.0 / (double)(float)Ld);
            for (int x = 0; x < K; x++) {
                float sv[4] = { 0, 0, 0, 0 };
                const float *rp = om->rfv + (size_t)x * Q * 4;
                for (int q = 0; q < Q; q++)
                    for (int r = 0; r < 4; r++) {
                        const int kk = r * Q + q + 1;
                        const float fm = kk <= M ? cntM[kk] * norm : 0.f, fi = kk <= M ? cntI[kk] * norm : 0.f;
                        sv[r] = sv[r] + fm * rp[q * 4 + r];
                        sv[r] = sv[r] + fi;
                    }
                null2[x] = (sv[0] + sv[1]) + (sv[2] + sv[3]);
            }
            for (int x = K; x < Kp; x++) null2[x] = 1.0f;
            for (int x = K + 1; x <= Kp - 3; x++) {   /* esl_abc_FAvgScVec: unweighted mean over the degeneracy set */
                float s = 0.f;
                for (int a = 0; a < degen_n[x]; a++) s += null2[degen_set[x * K + a]];
                if (degen_n[x] > 0) null2[x] = s / (float)degen_n[x];
            }
            for (; pos <= sqfrom; pos++) acc[pos] += 1.0f;            /* HMMER: "<=": residue sqfrom gets 1.0 too */
            for (; pos <= sqto; pos++) acc[pos] += null2[rd[pos - 1]];
        }
        for (; pos <= Lr; pos++) acc[pos] += 1.0f;
    }
    for (int p = 1; p <= Lr; p++) n2sc[ireg + p - 1] = (double)logf(acc[p] / (float)nsamples);
    if (out_nsampled) *out_nsampled = sp.n;

    /* ---- p7_spensemble_Cluster: single linkage, then per-cluster consensus endpoints ---- */
    int *asg = (int *)malloc((size_t)(sp.n > 0 ? sp.n : 1) * sizeof(int));
    int nc = 0;
    for (int a = 0; a < sp.n; a++) asg[a] = -1;
    int *stack = (int *)malloc((size_t)(sp.n > 0 ? sp.n : 1) * sizeof(int));
    for (int a = sp.n - 1; a >= 0; a--) {   /* Easel numbers the clusters from the last vertex down */
        if (asg[a] >= 0) continue;
        int top = 0;
        stack[top++] = a; asg[a] = nc;
        while (top > 0) {
            const int v = stack[--top];
            for (int b = 0; b < sp.n; b++)
                if (asg[b] < 0 && sp_link(&sp.v[v], &sp.v[b])) { asg[b] = nc; stack[top++] = b; }
        }
        nc++;
    }
    spc_t *sig = (spc_t *)malloc((size_t)(nc > 0 ? nc : 1) * sizeof(spc_t));
    int nsig = 0;
    int *epc = (int *)malloc((size_t)(Lr + M + 8) * sizeof(int));
    for (int c = 0; c < nc; c++) {
        int ninc = 0, last = -1;
        for (int h = 0; h < sp.n; h++)
            if (asg[h] == c) { if (sp.v[h].idx != last) ninc++; last = sp.v[h].idx; }
        if ((float)ninc / (float)nsamples < 0.25f) continue;
        int imin = 0, imax = 0, jmin = 0, jmax = 0, kmin = 0, kmax = 0, mmin = 0, mmax = 0;
        for (int h = 0; h < sp.n; h++)
            if (asg[h] == c) {
                const spc_t *e = &sp.v[h];
                if (imin == 0) { imin = imax = e->i; jmin = jmax = e->j; kmin = kmax = e->k; mmin = mmax = e->m; }
                else {
                    if (e->i < imin) imin = e->i; if (e->i > imax) imax = e->i;
                    if (e->j < jmin) jmin = e->j; if (e->j > jmax) jmax = e->j;
                    if (e->k < kmin) kmin = e->k; if (e->k > kmax) kmax = e->k;
                    if (e->m < mmin) mmin = e->m; if (e->m > mmax) mmax = e->m;
                }
            }
        const int thr = (int)ceilf((float)ninc * 0.02f);
        int best_i, best_j, best_k, best_m;
#define ENDPOINT(field, lo, hi, leftmost, best)                                                       \
        do {                                                                                          \
            for (int u = 0; u <= (hi) - (lo); u++) epc[u] = 0;                                        \
            for (int h = 0; h < sp.n; h++) if (asg[h] == c) epc[sp.v[h].field - (lo)]++;              \
            int found = 0;                                                                            \
            if (leftmost) { for (best = (lo); best <= (hi); best++) if (epc[best - (lo)] >= thr) { found = 1; break; } } \
            else { for (best = (hi); best >= (lo); best--) if (epc[best - (lo)] >= thr) { found = 1; break; } }           \
            if (!found) { int am = 0; for (int u = 1; u <= (hi) - (lo); u++) if (epc[u] > epc[am]) am = u; best = (lo) + am; } \
        } while (0)
        ENDPOINT(i, imin, imax, 1, best_i);
        ENDPOINT(k, kmin, kmax, 1, best_k);
        ENDPOINT(j, jmin, jmax, 0, best_j);
        ENDPOINT(m, mmin, mmax, 0, best_m);
        if (best_i > best_j || best_k > best_m) continue;
        spc_t s = { c, best_i, best_j, best_k, best_m, (float)ninc / (float)nsamples };
        sig[nsig++] = s;
    }
    for (int a = 1; a < nsig; a++) {   /* stable sort by start (glibc's qsort is a merge sort for small arrays) */
        const spc_t key = sig[a];
        int b = a - 1;
        while (b >= 0 && cmp_spc_i(&sig[b], &key) > 0) { sig[b + 1] = sig[b]; b--; }
        sig[b + 1] = key;
    }
    if (out_sig && n_sig)
        for (int d = 0; d < nsig; d++) {
            if (*n_sig < max_sig) {
                int *o = out_sig + 5 * (*n_sig);
                o[0] = sig[d].i; o[1] = sig[d].j; o[2] = sig[d].k; o[3] = sig[d].m; o[4] = (int)lrintf(sig[d].prob * (float)nsamples);
            }
            (*n_sig)++;
        }
    /* dominated domains (region_trace_ensemble) */
    int *dom = (int *)calloc(nsig > 0 ? nsig : 1, sizeof(int));
    for (int d = 0; d < nsig; d++)
        for (int d2 = d + 1; d2 < nsig; d2++) {
            const int nov = (sig[d].j < sig[d2].j ? sig[d].j : sig[d2].j) - (sig[d].i > sig[d2].i ? sig[d].i : sig[d2].i) + 1;
            if (nov == 0) break;
            const int la = sig[d].j - sig[d].i + 1, lb = sig[d2].j - sig[d2].i + 1;
            const int n = la < lb ? la : lb;
            if ((float)nov / (float)n >= 0.8f) {
                if (sig[d].prob > sig[d2].prob) dom[d2] = 1; else dom[d] = 1;
            }
        }
    int nout = 0;
    for (int d = 0; d < nsig; d++) {
        if (dom[d]) continue;
        if (nout < max_out) {
            out_env[nout * 5 + 0] = sig[d].i; out_env[nout * 5 + 1] = sig[d].j; out_env[nout * 5 + 2] = sig[d].k;
            out_env[nout * 5 + 3] = sig[d].m; out_env[nout * 5 + 4] = (int)lrintf(sig[d].prob * (float)nsamples);
        }
        nout++;
    }
    free(dom); free(epc); free(sig); free(stack); free(asg); free(sp.v);
    free(st); free(tk); free(ti); free(cntM); free(cntI); free(acc); free(fx.dp); free(fx.xmx);
    return nout;
}

