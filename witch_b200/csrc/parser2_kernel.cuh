// Family S, second generation: the multihit Forward + Backward parser of parser_kernel.cuh for TWO queries of one HMM per
// CTA, the two queries riding in the two halves of packed f32x2 registers (fma.rn.f32x2 / mul / add: one issue slot for
// both). Same mapping (thread j owns C consecutive columns, one barrier per Forward row, two per Backward row, the D->D
// chain as an affine scan), same arithmetic per query -- the transition parameters and every scan coefficient belong to
// the HMM and are shared by the pair, so they stay single registers (packed instructions take a scalar broadcast
// operand) or live in shared memory next to the emission rows (PMODE: 0 = all nine parameter sets in registers, 1 = all in
// shared memory, 2 = the six on the row's dependency chain in registers, MI / II / entry in shared memory).
// The shape that pays on a B200: 128-thread CTAs, 13 columns per thread (C % 4 != 0: shared-memory rows are
// [C/4][T][4] + a [T][C%4] tail), PMODE 2, compile-time CTA size (FIXT: every shared-memory offset is an immediate) --
// 245 registers, no spills, 2 CTAs/SM, half the warp instructions per cell of the one-query kernel.
// The queries of a pair are aligned by row index; A is the longer one. B's Forward runs on past its own end (its total is
// captured at row L_B), B's Backward starts when row L_B is reached (its state is exactly zero before). A lone last query
// is paired with itself (second result dropped).
#pragma once
#include "parser_kernel.cuh"

namespace witch {

#ifdef WITCH_HOST_SIM
static inline float2 p2_fma(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
static inline float2 p2_mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 p2_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
#else
__device__ __forceinline__ float2 p2_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 p2_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 p2_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
#endif
#ifndef WITCH_HOST_SIM
__device__ __forceinline__ float4 p2_lds4v(unsigned a) {   // volatile: the parameter rows are rewritten for the Backward pass
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
#else
static inline float4 p2_lds4v(unsigned a) { return lds_f4(a); }
#endif
__device__ __forceinline__ float2 p2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 p2b(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 p2_up(float2 v, int d) { return make_float2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d)); }
__device__ __forceinline__ float2 p2_down(float2 v, int d) { return make_float2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d)); }
__device__ __forceinline__ float2 p2_xor(float2 v, int d) { return make_float2(__shfl_xor_sync(0xffffffffu, v.x, d), __shfl_xor_sync(0xffffffffu, v.y, d)); }

// posterior decoding of the special states + region detection of ONE query from its Forward / Backward special rows
// (identical to the tail of mh_parser_kernel; all threads of the CTA call it)
__device__ __forceinline__ void parser_decode_regions(float *Fs, float *Bs, int Lr, int L, int sT, float invT, float ploop, float fwd_bits,
                                                      PairParse *res, int T, int tid) {
    float *dPB = Bs + 8 * Lr, *dPE = dPB + Lr, *dMO = dPE + Lr, *dBT = dMO + Lr, *dET = dBT + Lr;
    const int lane = tid & 31, w = tid >> 5;
    __syncthreads();
    for (int i = tid; i <= L; i += T) {
        const float4 f1a = reinterpret_cast<const float4 *>(Fs + 8 * i)[0], f1b = reinterpret_cast<const float4 *>(Fs + 8 * i)[1];
        const float4 b1a = reinterpret_cast<const float4 *>(Bs + 8 * i)[0], b1b = reinterpret_cast<const float4 *>(Bs + 8 * i)[1];
        const float fii = exp2f(f1b.y + b1b.y - (float)sT) * invT;
        dPB[i] = f1a.y * b1a.x * fii;
        dPE[i] = f1a.z * b1a.y * fii;
        float mo = 0.f;
        if (i > 0) {
            const float4 f0a = reinterpret_cast<const float4 *>(Fs + 8 * (i - 1))[0], f0b = reinterpret_cast<const float4 *>(Fs + 8 * (i - 1))[1];
            const float fpi = exp2f(f0b.y + b1b.y - (float)sT) * invT * ploop;
            mo = 1.0f - (f0a.x * b1a.z + f0a.w * b1a.w + f0b.x * b1b.x) * fpi;
        }
        dMO[i] = mo;
    }
    __syncthreads();
    if (w == 0) {
        const float rt1 = 0.25f, rt2 = 0.10f, rt3 = 0.20f;
        float cb = 0.f, ce = 0.f;
        for (int base = 0; base <= L; base += 32) {
            int idx = base + lane;
            float vb = (idx >= 1 && idx <= L) ? dPB[idx - 1] : 0.f;
            float ve = (idx >= 1 && idx <= L) ? dPE[idx] : 0.f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                float ub = __shfl_up_sync(0xffffffffu, vb, o), ue = __shfl_up_sync(0xffffffffu, ve, o);
                if (lane >= o) { vb += ub; ve += ue; }
            }
            vb += cb; ve += ce;
            if (idx <= L) { dBT[idx] = vb; dET[idx] = ve; }
            cb = __shfl_sync(0xffffffffu, vb, 31); ce = __shfl_sync(0xffffffffu, ve, 31);
        }
        __syncwarp();
        int nenv = 0, flags = 0, i0 = -1, trig = 0;
        for (int base = 1; base <= L; base += 32) {
            const int idx = base + lane;
            float mo = 0.f, db = 0.f, de = 0.f;
            if (idx <= L) { mo = dMO[idx]; db = dPB[idx - 1]; de = dPE[idx]; }
            const int lim = min(32, L - base + 1);
            for (int z = 0; z < lim; z++) {
                const float m = __shfl_sync(0xffffffffu, mo, z), b = __shfl_sync(0xffffffffu, db, z), ee = __shfl_sync(0xffffffffu, de, z);
                const int j = base + z;
                if (!trig) {
                    if (m - b < rt2) i0 = j; else if (i0 == -1) i0 = j;
                    if (m >= rt1) trig = 1;
                } else if (m - ee < rt2) {
                    float mx = -1.f;
                    const float eb = dET[i0 - 1], bj = dBT[j];
                    for (int zz = i0 + lane; zz <= j; zz += 32) mx = fmaxf(mx, fminf(dET[zz] - eb, bj - dBT[zz - 1]));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    if (mx >= rt3) flags |= 1 | (nenv < MAX_ENV ? (256 << nenv) : 0);
                    if (nenv < MAX_ENV && lane == 0) { res->env_i[nenv] = i0; res->env_j[nenv] = j; }
                    nenv++;
                    i0 = -1; trig = 0;
                }
            }
        }
        if (lane == 0) {
            res->fwd_bits = fwd_bits;
            res->nenv = nenv < MAX_ENV ? nenv : MAX_ENV;
            res->flags = flags | (nenv > MAX_ENV ? 4 : 0);
        }
    }
}

// shared-memory parameter rows, each laid out like an emission row: [C/4][T][4] floats + [T][C%4] tail. PMODE selects where
// the 9 per-column parameter sets of a thread live: 0 = registers (loaded once per item), 1 = all in shared memory (one
// LDS.128 per 4 columns and use), 2 = hybrid: the six that sit on the row's dependency chain (MM, IM, DM, MD, DD, the DD
// products) in registers, the three that do not (MI, II, entry) in shared memory.
enum { P2_A = 0, P2_B, P2_G, P2_MD, P2_DD, P2_MI, P2_II, P2_EN, P2_PDD, P2_NROWS };
__host__ __device__ constexpr int p2_smem_rows(int pmode) { return pmode == 1 ? (int)P2_NROWS : pmode == 2 ? 3 : 0; }
__host__ __device__ constexpr int p2_slot(int pmode, int row) { return pmode == 1 ? row : row - P2_MI; }   // hybrid: MI, II, EN -> 0, 1, 2
// element (thread j, column cc) of a row in the vector + tail layout
template <int C>
__device__ __forceinline__ int p2_index(int T, int j, int cc) {
    constexpr int V = C / 4, R = C % 4;
    return cc < 4 * V ? (cc >> 2) * (T * 4) + j * 4 + (cc & 3) : V * T * 4 + j * R + (cc - 4 * V);
}

// FIXT: the CTA always has MAXT threads, so every shared-memory offset of the row loops is an immediate
template <int C, int MAXT, int MINB, int PMODE, bool FIXT = false>
__global__ void __launch_bounds__(MAXT, MINB) mh_parser2_kernel(DevEhmm E, DevQueries Q, ParserWork Wk) {
    static_assert(C >= 4, "packed parser: at least one vector group of columns");
    constexpr bool PSMEM = PMODE == 1;     // hot rows (A, B, G, MD, DD, PDD) in shared memory
    constexpr bool PCOLD = PMODE >= 1;     // cold rows (MI, II, EN) in shared memory
    constexpr int V = C / 4, R = C % 4;
    WITCH_DYN_SMEM(float, smem);
    int T = FIXT ? MAXT : (int)blockDim.x, tid = threadIdx.x;
    if (!FIXT) PIN32(T);
    PIN32(tid);
    int lane = tid & 31, w = tid >> 5, NW = T >> 5;
    PIN32(lane); PIN32(w); PIN32(NW);
    const int TC = T * C;
    float *emis_s = smem;                                        // [nsym][TC]
    float *par_s = emis_s + (size_t)Q.nsym * TC;                 // [9][TC] (PSMEM only)
    float *s_red = par_s + (size_t)p2_smem_rows(PMODE) * TC;     // reduction area
    float2 *r2 = reinterpret_cast<float2 *>(s_red);              // per-row exchange, float2 per warp, double-buffered by row parity
    float *rc = s_red + 2 * 160;                                 // HMM constants per warp: PW[16], KW[16], AC[16]
#define R_TOT(par, x) r2[(par) * 16 + (x)]
#define R_ES(par, x) r2[32 + (par) * 16 + (x)]
#define R_BM(par, x) r2[64 + (par) * 16 + (x)]
#define R_BI(par, x) r2[96 + (par) * 16 + (x)]
#define R_BV(par, x) r2[128 + (par) * 16 + (x)]
#define R_PW(x) rc[(x)]
#define R_KW(x) rc[16 + (x)]
#define R_AC(x) rc[32 + (x)]
    __shared__ int s_item;
    const unsigned emis_ta = smem_u32(emis_s) + tid * 16;                      // vector part of this thread's columns
    const unsigned emis_tt = smem_u32(emis_s) + V * T * 16 + tid * (R * 4);     // tail part (C % 4 columns)
    const unsigned par_ta = smem_u32(par_s) + tid * 16, par_tt = smem_u32(par_s) + V * T * 16 + tid * (R * 4);
    const unsigned erow_b = TC * 4, estep = T * 16;

    const int Lr = (Wk.Lcap + 4) & ~3;
    const int stride_scr = PARSER_SCRATCH_ROWS * Lr;
    float *FsA = Wk.scratch + (size_t)blockIdx.x * 2 * stride_scr, *BsA = FsA + 8 * Lr;
    float *FsB = FsA + stride_scr, *BsB = FsB + 8 * Lr;

    for (int z = tid; z < 2 * 160 + 48; z += T) s_red[z] = 0.f;
    const int npairs = (Wk.nq + 1) >> 1;
    const long long nitems = (long long)Wk.nh * npairs;
    int loaded_h = -1;
    // parameter access: registers (loaded once per item) or shared memory (one LDS.128 per 4 columns and use)
    float ra[PSMEM ? 1 : C], rb[PSMEM ? 1 : C], rg[PSMEM ? 1 : C], rmd[PSMEM ? 1 : C], rdd[PSMEM ? 1 : C], rmi[PCOLD ? 1 : C],
        rii[PCOLD ? 1 : C], ren[PCOLD ? 1 : C], rpDD[PSMEM ? 1 : C];
    auto ldp = [&](const int row, float (&dst)[C]) {   // fetch one shared-memory parameter row of this thread's columns
        const unsigned ro = p2_slot(PMODE, row) * erow_b;
#pragma unroll
        for (int v = 0; v < V; v++) {
            const float4 t4 = p2_lds4v(par_ta + ro + v * estep);
            dst[4 * v] = t4.x; dst[4 * v + 1] = t4.y; dst[4 * v + 2] = t4.z; dst[4 * v + 3] = t4.w;
        }
#pragma unroll
        for (int r = 0; r < R; r++) dst[4 * V + r] = lds_f1v(par_tt + ro + 4 * r);
    };
    // emission odds of this thread's columns for the residue codes xa / xb of the two queries
    auto lde = [&](const int xa, const int xb, float (&da)[C], float (&db)[C]) {
        const unsigned a0 = emis_ta + xa * erow_b, b0 = emis_ta + xb * erow_b;
#pragma unroll
        for (int v = 0; v < V; v++) {
            const float4 t4 = lds_f4(a0 + v * estep), u4 = lds_f4(b0 + v * estep);
            da[4 * v] = t4.x; da[4 * v + 1] = t4.y; da[4 * v + 2] = t4.z; da[4 * v + 3] = t4.w;
            db[4 * v] = u4.x; db[4 * v + 1] = u4.y; db[4 * v + 2] = u4.z; db[4 * v + 3] = u4.w;
        }
        const unsigned a1 = emis_tt + xa * erow_b, b1 = emis_tt + xb * erow_b;
#pragma unroll
        for (int r = 0; r < R; r++) { da[4 * V + r] = lds_f1(a1 + 4 * r); db[4 * V + r] = lds_f1(b1 + 4 * r); }
    };
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (int)atomicAdd(Wk.counter, 1u);
        __syncthreads();
        const long long item = s_item;
        if (item >= nitems) break;
        const int h = Wk.hmms[item % Wk.nh];
        const int pidx = (int)(item / Wk.nh);
        const int qA = Wk.qorder[2 * pidx];
        const bool hasB = 2 * pidx + 1 < Wk.nq;
        const int qB = hasB ? Wk.qorder[2 * pidx + 1] : qA;   // a lone last query is paired with itself (second result dropped)
        const int LA = Q.len[qA], LB = Q.len[qB];            // queries are ordered longest first: LA >= LB
        const uint8_t *__restrict__ qdA = Q.dsq + Q.off[qA], *__restrict__ qdB = Q.dsq + Q.off[qB];
        const long long po = E.poff[h];
        PairParse *resA = Wk.out + (size_t)qA * E.H + h, *resB = Wk.out + (size_t)qB * E.H + h;
        if (LB <= 0) {   // empty queries (sorted last): nothing to score; a non-empty partner is scored as a pair with itself
            if (tid == 0) { resB->fwd_bits = 0.f; resB->nenv = 0; resB->flags = 0; if (LA <= 0) { resA->fwd_bits = 0.f; resA->nenv = 0; resA->flags = 0; } }
            if (LA <= 0) continue;
        }
        const int L = LA;                       // rows of the pair
        const int LBe = LB > 0 ? LB : LA;       // B's own length (B = A when B is empty)
        const uint8_t *qdBe = LB > 0 ? qdB : qdA;
        if (h != loaded_h) {
            const int st = E.stride[h];
            const float *eg = E.emis + E.eoff[h];
            for (int idx = tid; idx < Q.nsym * TC; idx += T) {
                int x = idx / TC, col = idx - x * TC;
                int j = col / C, cc = col - j * C;
                float v = (col < st - 1) ? __ldg(eg + (size_t)Q.symrow[x] * st + 1 + col) : 0.f;
                emis_s[(size_t)x * TC + p2_index<C>(T, j, cc)] = v;
            }
            loaded_h = h;
        }
        const int k0 = tid * C;
        const float2 pmove = p2(3.0f / ((float)LA + 3.0f), 3.0f / ((float)LBe + 3.0f));
        const float2 ploop = p2(1.0f - pmove.x, 1.0f - pmove.y);
        const float2 HALF = p2b(0.5f);

        // =========================== Forward ===========================
        float pa[C], pb[C], pg[C], pmd[C], pdd[C], pmi[C], pii[C], pen[C], pDD[C];
        load_cols<C>(E.tMM + po, k0, pa);
        load_cols<C>(E.tIM + po, k0, pb);
        load_cols<C>(E.tDM + po, k0, pg);
        load_cols<C>(E.tMD + po, k0, pmd);
        load_cols<C>(E.tDD + po, k0, pdd);
        load_cols<C>(E.tMI + po, k0 + 1, pmi);
        load_cols<C>(E.tII + po, k0 + 1, pii);
        load_cols<C>(E.entry + po, k0 + 1, pen);
        const float mdo = __ldg(E.tMD + po + k0 + C), ddo = __ldg(E.tDD + po + k0 + C);
        pDD[0] = 1.f;
#pragma unroll
        for (int c = 1; c < C; c++) pDD[c] = pDD[c - 1] * pdd[c];
        float coef[5], Cexcl, Rt, KKl, czl, czl2, Aprev;
        constexpr int CW = (MAXT <= 128) ? 4 : (MAXT <= 256) ? 8 : 16;
        const int lw = lane & (CW - 1);
        {
            float SP = 0.f;
#pragma unroll
            for (int c = 0; c < C; c++) SP += pDD[c];
            const float Pt = pDD[C - 1] * ddo;
            float Pc = Pt;
#pragma unroll
            for (int s = 0; s < 5; s++) {
                float up = __shfl_up_sync(0xffffffffu, Pc, 1 << s);
                coef[s] = (lane >= (1 << s)) ? Pc : 0.f;
                if (lane >= (1 << s)) Pc *= up;
            }
            Cexcl = __shfl_up_sync(0xffffffffu, Pc, 1);
            if (lane == 0) Cexcl = 1.f;
            Rt = 0.f;
            for (int it = 0; it < 31; it++) {
                const float dn = __shfl_down_sync(0xffffffffu, fmaf(Pt, Rt, SP), 1);
                Rt = (lane < 31) ? dn : 0.f;
            }
            float kw = SP * Cexcl;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) kw += __shfl_xor_sync(0xffffffffu, kw, o);
            __syncthreads();   // (previous item's readers of the constants are done)
            if (lane == 31) { R_PW(w) = Pc; R_AC(w) = pDD[C - 1] * Cexcl; }
            if (lane == 0) R_KW(w) = kw;
        }
#pragma unroll
        for (int c = 0; c < C; c++) {   // forward parameter rows -> shared memory / registers
            const int ix = p2_index<C>(T, tid, c);
            if (PSMEM) {
                par_s[p2_slot(PMODE, P2_A) * TC + ix] = pa[c]; par_s[p2_slot(PMODE, P2_B) * TC + ix] = pb[c]; par_s[p2_slot(PMODE, P2_G) * TC + ix] = pg[c];
                par_s[p2_slot(PMODE, P2_MD) * TC + ix] = pmd[c]; par_s[p2_slot(PMODE, P2_DD) * TC + ix] = pdd[c]; par_s[p2_slot(PMODE, P2_PDD) * TC + ix] = pDD[c];
            } else { ra[c] = pa[c]; rb[c] = pb[c]; rg[c] = pg[c]; rmd[c] = pmd[c]; rdd[c] = pdd[c]; rpDD[c] = pDD[c]; }
            if (PCOLD) { par_s[p2_slot(PMODE, P2_MI) * TC + ix] = pmi[c]; par_s[p2_slot(PMODE, P2_II) * TC + ix] = pii[c]; par_s[p2_slot(PMODE, P2_EN) * TC + ix] = pen[c]; }
            else { rmi[c] = pmi[c]; rii[c] = pii[c]; ren[c] = pen[c]; }
        }
        const float pDDlast = pDD[C - 1];
        __syncthreads();
        {
            czl = 0.f; czl2 = 0.f; KKl = 0.f;
            if (lw < w) { czl = 1.f; for (int ww = lw + 1; ww < w; ww++) czl *= R_PW(ww); }
            if (lw < w - 1) { czl2 = 1.f; for (int ww = lw + 1; ww < w - 1; ww++) czl2 *= R_PW(ww); }
            if (lw < NW) {
                float pr = 1.f;
                for (int w2 = lw + 1; w2 < NW; w2++) { KKl = fmaf(R_KW(w2), pr, KKl); pr *= R_PW(w2); }
            }
            Aprev = (w > 0) ? R_AC(w - 1) : 0.f;
        }
        float2 sM[C], sI[C], sD[C], dl[C];
#pragma unroll
        for (int c = 0; c < C; c++) { sM[c] = p2b(0.f); sI[c] = p2b(0.f); sD[c] = p2b(0.f); dl[c] = p2b(0.f); }
        float2 xN = p2b(1.f), xB = pmove, xE = p2b(0.f), xJ = p2b(0.f), xC = p2b(0.f), yex = p2b(0.f), dL = p2b(0.f);
        int sFA = 0, sFB = 0;
        float capC = 0.f; int capS = 0;   // B's C(L_B) and exponent, captured when its last row is finished
        if (tid == 0) {
            FsA[0] = 1.f; FsA[1] = pmove.x; FsA[2] = 0.f; FsA[3] = 0.f; FsA[4] = 0.f; FsA[5] = 0.f;
            FsB[0] = 1.f; FsB[1] = pmove.y; FsB[2] = 0.f; FsB[3] = 0.f; FsB[4] = 0.f; FsB[5] = 0.f;
        }
        __syncthreads();
        int xrA = qdA[0], xrB = qdBe[0];
        // finish row r (after its barrier): D values, E(r), specials; returns the rescale factors
        auto finish_row = [&](const int r) -> float2 {
            const int par = r & 1;
            const float2 tv = R_TOT(par, lw), ev = R_ES(par, lw);
            float2 z = p2_mul(tv, p2b(czl)), z2 = p2_mul(tv, p2b(czl2)), et = p2_fma(tv, p2b(KKl), ev);
#pragma unroll
            for (int o = CW / 2; o > 0; o >>= 1) { z = p2_add(z, p2_xor(z, o)); z2 = p2_add(z2, p2_xor(z2, o)); et = p2_add(et, p2_xor(et, o)); }
            const float2 X = p2_fma(p2b(Cexcl), z, yex);
            if (PSMEM) { float t_[C]; ldp(P2_PDD, t_);
#pragma unroll
                for (int c = 0; c < C; c++) sD[c] = p2_fma(p2b(t_[c]), X, dl[c]);
            } else {
#pragma unroll
                for (int c = 0; c < C; c++) sD[c] = p2_fma(p2b(rpDD[c]), X, dl[c]);
            }
            dL = p2_up(sD[C - 1], 1);
            if (lane == 0) dL = (w > 0) ? p2_fma(p2b(Aprev), z2, R_BV(par, w - 1)) : p2b(0.f);
            xE = et;
            xJ = p2_fma(xJ, ploop, p2_mul(et, HALF)); xC = p2_fma(xC, ploop, p2_mul(et, HALF)); xN = p2_mul(xN, ploop); xB = p2_mul(p2_add(xN, xJ), pmove);
            float2 scl = p2b(1.f);
            if (et.x > 1.0e12f || et.y > 1.0e12f) {
                if (et.x > 1.0e12f) { const int e = fexp(et.x); scl.x = pow2i(-e); sFA += e; }
                if (et.y > 1.0e12f) { const int e = fexp(et.y); scl.y = pow2i(-e); sFB += e; }
                xE = p2_mul(xE, scl); xJ = p2_mul(xJ, scl); xC = p2_mul(xC, scl); xN = p2_mul(xN, scl); xB = p2_mul(xB, scl); dL = p2_mul(dL, scl);
#pragma unroll
                for (int c = 0; c < C; c++) { sM[c] = p2_mul(sM[c], scl); sI[c] = p2_mul(sI[c], scl); sD[c] = p2_mul(sD[c], scl); }
            }
            return scl;
        };
        float4 *fsA = reinterpret_cast<float4 *>(FsA + 8), *fsB = reinterpret_cast<float4 *>(FsB + 8);
        for (int i = 1; i <= L; i++) {
            float2 scl = p2b(1.f);
            if (i > 1) {
                scl = finish_row(i - 1);
                if (tid == 0) {
                    fsA[0] = make_float4(xN.x, xB.x, xE.x, xJ.x); fsA[1] = make_float4(xC.x, (float)sFA, 0.f, 0.f);
                    if (i - 1 <= LBe) { fsB[0] = make_float4(xN.y, xB.y, xE.y, xJ.y); fsB[1] = make_float4(xC.y, (float)sFB, 0.f, 0.f); }
                }
                fsA += 2; fsB += 2;
                if (i - 1 == LBe) { capC = xC.y; capS = sFB; }
            }
            float ea[C], eb[C];
            lde(xrA, xrB, ea, eb);
            if (i < L) { xrA = qdA[i]; xrB = qdBe[min(i, LBe - 1)]; }
            float2 mL = p2_up(sM[C - 1], 1), iL = p2_up(sI[C - 1], 1);
            if (lane == 0) {
                if (w > 0 && i > 1) { const int par = (i - 1) & 1; mL = p2_mul(R_BM(par, w - 1), scl); iL = p2_mul(R_BI(par, w - 1), scl); }
                else { mL = p2b(0.f); iL = p2b(0.f); }
            }
            float2 nM[C];
            {
                float qa_[C], qb_[C], qg_[C], qmi_[C], qii_[C], qen_[C];
                if (PSMEM) { ldp(P2_A, qa_); ldp(P2_B, qb_); ldp(P2_G, qg_); }
                if (PCOLD) { ldp(P2_MI, qmi_); ldp(P2_II, qii_); ldp(P2_EN, qen_); }
#pragma unroll
                for (int c = C - 1; c >= 0; c--) {
                    const float2 pm = c > 0 ? sM[c - 1] : mL, pi = c > 0 ? sI[c - 1] : iL, pd = c > 0 ? sD[c - 1] : dL;
                    const float tmi = PCOLD ? qmi_[c] : rmi[c], tii = PCOLD ? qii_[c] : rii[c], ten = PCOLD ? qen_[c] : ren[c];
                    const float ta_ = PSMEM ? qa_[c] : ra[c], tb_ = PSMEM ? qb_[c] : rb[c], tg_ = PSMEM ? qg_[c] : rg[c];
                    sI[c] = p2_fma(sM[c], p2b(tmi), p2_mul(sI[c], p2b(tii)));
                    float2 acc = p2_mul(xB, p2b(ten));
                    acc = p2_fma(pm, p2b(ta_), acc); acc = p2_fma(pi, p2b(tb_), acc); acc = p2_fma(pd, p2b(tg_), acc);
                    nM[c] = p2(acc.x * ea[c], acc.y * eb[c]);
                }
            }
            {
                float qmd_[C], qdd_[C];
                if (PSMEM) { ldp(P2_MD, qmd_); ldp(P2_DD, qdd_); }
                dl[0] = p2b(0.f);
#pragma unroll
                for (int c = 1; c < C; c++) dl[c] = p2_fma(nM[c - 1], p2b(PSMEM ? qmd_[c] : rmd[c]), p2_mul(dl[c - 1], p2b(PSMEM ? qdd_[c] : rdd[c])));
            }
            const float2 yl = p2_fma(nM[C - 1], p2b(mdo), p2_mul(dl[C - 1], p2b(ddo)));
            float2 at4[4] = {p2b(0.f), p2b(0.f), p2b(0.f), p2b(0.f)};   // four partial sums: no 13-deep chain of dependent adds
#pragma unroll
            for (int c = 0; c < C; c++) { sM[c] = nM[c]; at4[c & 3] = p2_add(at4[c & 3], p2_add(nM[c], dl[c])); }
            const float2 at = p2_add(p2_add(at4[0], at4[1]), p2_add(at4[2], at4[3]));
            float2 y = yl, v = p2_fma(yl, p2b(Rt), at);
#pragma unroll
            for (int s = 0; s < 5; s++) {
                const float2 up = p2_up(y, 1 << s);
                v = p2_add(v, p2_xor(v, 1 << s));
                y = p2_fma(p2b(coef[s]), up, y);
            }
            yex = p2_up(y, 1);
            if (lane == 0) yex = p2b(0.f);
            {
                const int par = i & 1;
                if (lane == 31) {
                    R_TOT(par, w) = y; R_BM(par, w) = nM[C - 1]; R_BI(par, w) = sI[C - 1];
                    R_BV(par, w) = p2_fma(p2b(pDDlast), yex, dl[C - 1]);
                }
                if (lane == 0) R_ES(par, w) = v;
            }
            __syncthreads();
        }
        finish_row(L);
        if (tid == 0) {
            float4 *r = reinterpret_cast<float4 *>(FsA + 8 * L);
            r[0] = make_float4(xN.x, xB.x, xE.x, xJ.x); r[1] = make_float4(xC.x, (float)sFA, 0.f, 0.f);
            if (L == LBe) { float4 *rB = reinterpret_cast<float4 *>(FsB + 8 * L); rB[0] = make_float4(xN.y, xB.y, xE.y, xJ.y); rB[1] = make_float4(xC.y, (float)sFB, 0.f, 0.f); }
        }
        if (L == LBe) { capC = xC.y; capS = sFB; }
        const float2 Tm = p2(xC.x * pmove.x, capC * pmove.y);
        const int sTA = sFA, sTB = capS;
        const float fwdA = log2f(Tm.x) + (float)sTA, fwdB = log2f(Tm.y) + (float)sTB;

        // =========================== Backward ===========================
        load_cols<C>(E.tMM + po, k0 + 1, pa);
        load_cols<C>(E.tIM + po, k0 + 1, pb);
        load_cols<C>(E.tDM + po, k0 + 1, pg);
        load_cols<C>(E.tMD + po, k0 + 1, pmd);
        load_cols<C>(E.tDD + po, k0 + 1, pdd);
        const float GX = __ldg(E.gD + po + k0 + C + 1);
        __syncthreads();
        {
            float Pc = pdd[0];
#pragma unroll
            for (int c = 1; c < C; c++) Pc *= pdd[c];
#pragma unroll
            for (int s = 0; s < 5; s++) {
                float dn = __shfl_down_sync(0xffffffffu, Pc, 1 << s);
                coef[s] = (lane + (1 << s) < 32) ? Pc : 0.f;
                if (lane + (1 << s) < 32) Pc *= dn;
            }
            Cexcl = __shfl_down_sync(0xffffffffu, Pc, 1);
            if (lane == 31) Cexcl = 1.f;
            if (lane == 0) R_PW(w) = Pc;
        }
        if (PSMEM) {   // backward rows: MM, IM, DM, MD, DD of the owned columns (MI, II, entry are already in place)
#pragma unroll
            for (int c = 0; c < C; c++) {
                const int ix = p2_index<C>(T, tid, c);
                par_s[P2_A * TC + ix] = pa[c]; par_s[P2_B * TC + ix] = pb[c]; par_s[P2_G * TC + ix] = pg[c]; par_s[P2_MD * TC + ix] = pmd[c];
                par_s[P2_DD * TC + ix] = pdd[c];
            }
        } else {
#pragma unroll
            for (int c = 0; c < C; c++) { ra[c] = pa[c]; rb[c] = pb[c]; rg[c] = pg[c]; rmd[c] = pmd[c]; rdd[c] = pdd[c]; }
        }
        __syncthreads();
        czl = 0.f;
        if (lw > w && lw < NW) { czl = 1.f; for (int ww = w + 1; ww < lw; ww++) czl *= R_PW(ww); }
#pragma unroll
        for (int c = 0; c < C; c++) { sM[c] = p2b(0.f); sI[c] = p2b(0.f); sD[c] = p2b(0.f); }
        float2 bN = p2b(0.f), bJ = p2b(0.f), bC = p2b(0.f), bE = p2b(0.f);
        int sBA = 0, sBB = 0;
        const float2 invT = p2(1.0f / Tm.x, 1.0f / Tm.y);
        float4 *bsA = reinterpret_cast<float4 *>(BsA + 8 * L), *bsB = reinterpret_cast<float4 *>(BsB + 8 * L);
        for (int i = L; i >= 0; i--) {
            const int par = i & 1;
            float2 mn[C], mnR[C];
            float2 eR = p2b(0.f);
            if (i < L) {
                const int xa = qdA[i], xb = qdBe[min(i, LBe - 1)];
                const unsigned a0 = emis_ta + xa * erow_b, b0 = emis_ta + xb * erow_b;
                float ea[C], eb[C];
                lde(xa, xb, ea, eb);
#pragma unroll
                for (int c = 0; c < C; c++) mn[c] = p2(sM[c].x * ea[c], sM[c].y * eb[c]);
                if (tid + 1 < T) eR = p2(lds_f1(a0 + 16), lds_f1(b0 + 16));
            } else {
#pragma unroll
                for (int c = 0; c < C; c++) mn[c] = p2b(0.f);
            }
            float2 nb = p2_down(mn[0], 1);
            if (lane == 31) nb = (w + 1 < NW && i < L) ? p2_mul(R_BM((i + 1) & 1, w + 1), eR) : p2b(0.f);
            float2 bp4[4] = {p2b(0.f), p2b(0.f), p2b(0.f), p2b(0.f)};
            float2 Mp[C], nI[C], tm[C];
            {
                float qen_[C], qa_[C], qb_[C], qg_[C], qmi_[C], qii_[C];
                if (PSMEM) { ldp(P2_A, qa_); ldp(P2_B, qb_); ldp(P2_G, qg_); }
                if (PCOLD) { ldp(P2_EN, qen_); ldp(P2_MI, qmi_); ldp(P2_II, qii_); }
#pragma unroll
                for (int c = 0; c < C; c++) {
                    mnR[c] = (c < C - 1) ? mn[c + 1] : nb;
                    bp4[c & 3] = p2_fma(mn[c], p2b(PCOLD ? qen_[c] : ren[c]), bp4[c & 3]);
                }
#pragma unroll
                for (int c = 0; c < C; c++) {
                    Mp[c] = p2_fma(mnR[c], p2b(PSMEM ? qa_[c] : ra[c]), p2_mul(sI[c], p2b(PCOLD ? qmi_[c] : rmi[c])));
                    nI[c] = p2_fma(mnR[c], p2b(PSMEM ? qb_[c] : rb[c]), p2_mul(sI[c], p2b(PCOLD ? qii_[c] : rii[c])));
                    tm[c] = p2_mul(mnR[c], p2b(PSMEM ? qg_[c] : rg[c]));
                }
            }
            float2 bp = p2_add(p2_add(bp4[0], bp4[1]), p2_add(bp4[2], bp4[3]));
            float qdd_[C], qmd_[C];
            if (PSMEM) { ldp(P2_DD, qdd_); ldp(P2_MD, qmd_); }
            float2 y = tm[C - 1];
#pragma unroll
            for (int c = C - 2; c >= 0; c--) y = p2_fma(y, p2b(PSMEM ? qdd_[c] : rdd[c]), tm[c]);
#pragma unroll
            for (int s = 0; s < 5; s++) {
                const float2 dn = p2_down(y, 1 << s);
                bp = p2_add(bp, p2_xor(bp, 1 << s));
                y = p2_fma(p2b(coef[s]), dn, y);
            }
            float2 yexb = p2_down(y, 1);
            if (lane == 31) yexb = p2b(0.f);
            if (lane == 0) { R_ES(par, w) = bp; R_TOT(par, w) = y; }
            __syncthreads();
            float2 Bi = R_ES(par, lw), Z = p2_mul(R_TOT(par, lw), p2b(czl));
#pragma unroll
            for (int o = CW / 2; o > 0; o >>= 1) { Bi = p2_add(Bi, p2_xor(Bi, o)); Z = p2_add(Z, p2_xor(Z, o)); }
            // specials; each query starts at its own last row (B is exactly zero before)
            bJ = p2_fma(bJ, ploop, p2_mul(Bi, pmove)); bC = p2_mul(bC, ploop); bN = p2_fma(bN, ploop, p2_mul(Bi, pmove));
            if (i == L) { bC.x = pmove.x; bJ.x = 0.f; bN.x = 0.f; }
            if (i == LBe) { bC.y = pmove.y; bJ.y = 0.f; bN.y = 0.f; }
            bE = p2_mul(p2_add(bJ, bC), HALF);
            float2 Xp = p2_fma(p2b(Cexcl), Z, yexb);
            {
                const float bigA = fmaxf(fmaxf(bN.x, bJ.x), Bi.x), bigB = fmaxf(fmaxf(bN.y, bJ.y), Bi.y);
                if (bigA > 1.0e9f || bigB > 1.0e9f) {
                    float2 scl = p2b(1.f);
                    if (bigA > 1.0e9f) { const int e = fexp(bigA); scl.x = pow2i(-e); sBA += e; }
                    if (bigB > 1.0e9f) { const int e = fexp(bigB); scl.y = pow2i(-e); sBB += e; }
                    bN = p2_mul(bN, scl); bJ = p2_mul(bJ, scl); bC = p2_mul(bC, scl); bE = p2_mul(bE, scl); Bi = p2_mul(Bi, scl); Xp = p2_mul(Xp, scl);
#pragma unroll
                    for (int c = 0; c < C; c++) { Mp[c] = p2_mul(Mp[c], scl); nI[c] = p2_mul(nI[c], scl); tm[c] = p2_mul(tm[c], scl); }
                }
            }
            if (tid == 0) {
                bsA[0] = make_float4(Bi.x, bE.x, bN.x, bJ.x); bsA[1] = make_float4(bC.x, (float)sBA, 0.f, 0.f);
                if (i <= LBe) { bsB[0] = make_float4(Bi.y, bE.y, bN.y, bJ.y); bsB[1] = make_float4(bC.y, (float)sBB, 0.f, 0.f); }
            }
            bsA -= 2; bsB -= 2;
            if (i == 0) break;
            const float2 X = p2_fma(bE, p2b(GX), Xp);
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                const float2 dr = (c < C - 1) ? sD[c + 1] : X;
                sD[c] = p2_fma(p2b(PSMEM ? qdd_[c] : rdd[c]), dr, p2_add(tm[c], bE));
                sM[c] = p2_fma(p2b(PSMEM ? qmd_[c] : rmd[c]), dr, p2_add(Mp[c], bE));
                sI[c] = nI[c];
            }
            if (lane == 0) R_BM(i & 1, w) = sM[0];
            __syncthreads();
        }
        if (Wk.dbg_bwd != nullptr && tid == 0) {
            Wk.dbg_bwd[(size_t)qA * E.H + h] = (logf(bN.x) + (float)sBA * 0.69314718056f);
            if (hasB && LB > 0) Wk.dbg_bwd[(size_t)qB * E.H + h] = (logf(bN.y) + (float)sBB * 0.69314718056f);
        }
        // ---- decoding of the special states + regions, one query after the other ----
        parser_decode_regions(FsA, BsA, Lr, LA, sTA, invT.x, ploop.x, fwdA, resA, T, tid);
        if (hasB && LB > 0) parser_decode_regions(FsB, BsB, Lr, LB, sTB, invT.y, ploop.y, fwdB, resB, T, tid);
    }
#undef R_TOT
#undef R_ES
#undef R_BM
#undef R_BI
#undef R_BV
#undef R_PW
#undef R_KW
#undef R_AC
}

}  // namespace witch
