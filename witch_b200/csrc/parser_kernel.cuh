// Family S: multihit-local Forward + Backward "parser" pass with posterior domain decoding and region
// detection (hmmsearch's ForwardParser / BackwardParser / DomainDecoding / region heuristics, SURVEY.md 8(a)
// "Score semantics" items 3-6).
//
// Mapping: one CTA per (query, HMM) pair, row-synchronous. Thread j owns C consecutive model columns and keeps
// their transition parameters and the M/I/D state of the current row in REGISTERS; match emission odds for the
// symbols present in the query set are staged in shared memory (vector-load friendly layout). The multihit
// feedback B(i) <- E(i) makes every row depend on a CTA-wide sum, so a row is two (Forward) or three (Backward)
// barrier-separated stages; the in-row D->D chain is resolved exactly with a two-level (warp shuffle, then
// cross-warp) scan of affine maps whose multiplicative parts are per-thread constants.
// Arithmetic: FP32 probability space with power-of-two row rescaling (the exponent is an exact integer).
#pragma once
#include "device_types.cuh"

namespace witch {

struct ParserWork {
    const int *hmms;    // HMM indices of this launch class
    int nh;
    const int *qorder;  // query indices, longest first
    int nq;
    int Lcap;           // scratch rows per CTA (>= max query length + 1)
    float *scratch;     // per CTA: 6*(Lcap+1) Forward specials + 3*(Lcap+1) decode arrays + 2*(Lcap+1) prefix sums
    unsigned *counter;  // dynamic work counter
    PairParse *out;     // [n_queries * H]
    float *dbg_bwd;     // optional [n_queries * H] backward total (nats), debug only
};

__device__ __forceinline__ int fexp(float v) { return ((__float_as_int(v) >> 23) & 0xff) - 127; }
__device__ __forceinline__ float pow2i(int e) {
    e = e < -126 ? -126 : (e > 127 ? 127 : e);
    return __int_as_float((e + 127) << 23);
}

template <int C>
__device__ __forceinline__ void load_cols(const float *__restrict__ p, long long k, float (&r)[C]) {
#pragma unroll
    for (int c = 0; c < C; c++) r[c] = __ldg(p + k + c);
}

// shared-memory emission row for symbol x: element (thread j, column cc)
template <int C>
__device__ __forceinline__ int emis_index(int T, int j, int cc) {
    if (C % 4 == 0) return (cc >> 2) * (T * 4) + j * 4 + (cc & 3);
    return j * C + cc;
}

template <int C>
__device__ __forceinline__ void load_emis(const float *row, int T, int j, float (&e)[C]) {
    if (C % 4 == 0) {
#pragma unroll
        for (int v = 0; v < C / 4; v++) {
            float4 t = *reinterpret_cast<const float4 *>(row + v * (T * 4) + j * 4);
            e[4 * v] = t.x; e[4 * v + 1] = t.y; e[4 * v + 2] = t.z; e[4 * v + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; c++) e[c] = row[j * C + c];
    }
}

constexpr int S_RED = 32;  // words per row of the reduction area
constexpr int PARSER_RED_ROWS = 13;

#ifndef WITCH_HOST_SIM   // (tools/sim/simt.h provides host versions of these helpers)
// 32-bit shared-memory addressing helpers: keep hot-loop addresses as one pinned register + immediate offsets
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds_f4(unsigned a) {
    float4 v;
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f1(unsigned a) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f1v(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f1(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ int lds_u8(unsigned a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return (int)v; }
// Opaque identity: stops the compiler from rebuilding a cheap-looking address expression inside the row loop
#define PIN32(x) asm volatile("" : "+r"(x))
#define PIN64(x) asm volatile("" : "+l"(x))
#endif
constexpr int PARSER_SCRATCH_ROWS = 21;  // floats of scratch per sequence row and CTA

template <int C, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) mh_parser_kernel(DevEhmm E, DevQueries Q, ParserWork Wk) {
    WITCH_DYN_SMEM(float, smem);
    int T = blockDim.x, tid = threadIdx.x;
    PIN32(T); PIN32(tid);
    int lane = tid & 31, w = tid >> 5, NW = T >> 5;
    PIN32(lane); PIN32(w); PIN32(NW);
    const int TC = T * C;
    float *emis_s = smem;                       // [nsym][TC]
    float *s_tot = emis_s + (size_t)Q.nsym * TC;  // reduction area: PARSER_RED_ROWS rows of 32 words
    unsigned red_sa = smem_u32(s_tot);
    unsigned emis_ta = smem_u32(emis_s) + tid * 16;  // this thread's slot in an emission row (interleaved layout)
    unsigned erow_b = TC * 4, estep = T * 16;
    PIN32(red_sa); PIN32(emis_ta); PIN32(erow_b); PIN32(estep);
    // reduction area (32-word rows): TOT[2], ES[2], BM[2], BI[2], BV[2] double-buffered by row parity; PW, KW, AC constants
#define SA_TOT2(par, x) (red_sa + (par) + 4 * (x))
#define SA_ES2(par, x) (red_sa + 256 + (par) + 4 * (x))
#define SA_BM2(par, x) (red_sa + 512 + (par) + 4 * (x))
#define SA_BI2(par, x) (red_sa + 768 + (par) + 4 * (x))
#define SA_BV2(par, x) (red_sa + 1024 + (par) + 4 * (x))
#define SA_PW(x) (red_sa + 1280 + 4 * (x))
#define SA_KW(x) (red_sa + 1408 + 4 * (x))
#define SA_AC(x) (red_sa + 1536 + 4 * (x))
    // (Backward pass names for the same rows)
#define SA_TOT(x) SA_TOT2(0, x)
#define SA_ES(x) SA_ES2(0, x)
#define SA_BM(par, x) SA_BM2((par) * 128, x)
    __shared__ int s_item;

    const int Lr = (Wk.Lcap + 4) & ~3;                         // rows, rounded so that every array stays 16-byte aligned
    const int stride_scr = PARSER_SCRATCH_ROWS * Lr;
    float *Fs = Wk.scratch + (size_t)blockIdx.x * stride_scr;  // [(L+1)][8]: Forward N,B,E,J,C,exp
    float *Bs = Fs + 8 * Lr;                                   // [(L+1)][8]: Backward B,E,N,J,C,exp
    float *dPB = Bs + 8 * Lr;                                  // P(B at i)
    float *dPE = dPB + Lr;                                     // P(E at i)
    float *dMO = dPE + Lr;                                     // mocc[i]
    float *dBT = dMO + Lr;                                     // btot prefix
    float *dET = dBT + Lr;                                     // etot prefix

    for (int z = tid; z < PARSER_RED_ROWS * S_RED; z += T) s_tot[z] = 0.f;  // slots of warps >= NW stay zero
    const long long nitems = (long long)Wk.nh * Wk.nq;
    int loaded_h = -1;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (int)atomicAdd(Wk.counter, 1u);
        __syncthreads();
        const long long item = s_item;
        if (item >= nitems) break;
        const int h = Wk.hmms[item % Wk.nh];
        const int q = Wk.qorder[item / Wk.nh];
        const int L = Q.len[q];
        const uint8_t *__restrict__ qd = Q.dsq + Q.off[q];
        const long long po = E.poff[h];
        PairParse *res = Wk.out + (size_t)q * E.H + h;
        if (L <= 0) {
            if (tid == 0) { res->fwd_bits = 0.f; res->nenv = 0; res->flags = 0; }
            continue;
        }
        // ---- stage the emission rows of this HMM (only when the HMM changes) ----
        if (h != loaded_h) {
            const int st = E.stride[h];
            const float *eg = E.emis + E.eoff[h];
            for (int idx = tid; idx < Q.nsym * TC; idx += T) {
                int x = idx / TC, col = idx - x * TC;
                int j = col / C, cc = col - j * C;
                float v = (col < st - 1) ? __ldg(eg + (size_t)Q.symrow[x] * st + 1 + col) : 0.f;
                emis_s[(size_t)x * TC + emis_index<C>(T, j, cc)] = v;
            }
            loaded_h = h;
        }
        const int k0 = tid * C;  // owns model columns k0+1 .. k0+C
        const float nj = 1.0f;
        const float pmove = (2.0f + nj) / ((float)L + 2.0f + nj), ploop = 1.0f - pmove;
        const float EC = 0.5f, EJ = 0.5f;

        // =========================== Forward ===========================
        // One barrier per row. Everything a row needs from the other threads is linear in per-thread quantities
        // that are known before the barrier, so the three CTA-wide combinations run side by side after it:
        //   Z_w    = D entering this warp's first column          = sum_l tot[l] * cz[l]
        //   Z_{w-1} (for the D value the left warp's last column hands to lane 0 of this warp)
        //   E(i)   = sum_t [a_t + SP_t * X_t] = sum_w es[w] + sum_l tot[l] * KK[l]
        // with a_t = sum_c (M + local D), SP_t = sum_c pDD[c], X_t the D value entering the thread. All the weights
        // (cz, cz2, KK, R, K_w, A_w) are products/sums of tDD over columns: constants of the (HMM, thread).
        float pa[C], pb[C], pg[C], pmd[C], pdd[C], pmi[C], pii[C], pen[C], pDD[C];
        load_cols<C>(E.tMM + po, k0, pa);   // into column k0+1+cc from node k0+cc
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
        // cross-warp combinations: lane lw stands for warp lw; every group of CW lanes mirrors warps 0..CW-1 (NW <= CW), so a
        // log2(CW)-level butterfly finishes them: 3 levels for CTAs of at most 8 warps, 4 otherwise
        constexpr int CW = (MAXT <= 256) ? 8 : 16;
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
            // R_l = sum_{l' > l} SP_l' * prod_{l < l'' < l'} P_l''   (weight of this thread's local chain output in E)
            Rt = 0.f;
            for (int it = 0; it < 31; it++) {
                const float dn = __shfl_down_sync(0xffffffffu, fmaf(Pt, Rt, SP), 1);
                Rt = (lane < 31) ? dn : 0.f;
            }
            float kw = SP * Cexcl;   // K_w = sum_l SP_l * Cexcl_l
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) kw += __shfl_xor_sync(0xffffffffu, kw, o);
            if (lane == 31) { sts_f1(SA_PW(w), Pc); sts_f1(SA_AC(w), pDD[C - 1] * Cexcl); }
            if (lane == 0) sts_f1(SA_KW(w), kw);
        }
        __syncthreads();
        {
            // lane lw stands for warp lw: cz = prod_{lw < w'' < w} PW[w''] (0 for lw >= w), cz2 the same for w-1,
            // KK = sum_{w2 > lw} K[w2] * prod_{lw < w'' < w2} PW[w'']
            czl = 0.f; czl2 = 0.f; KKl = 0.f;
            if (lw < w) { czl = 1.f; for (int ww = lw + 1; ww < w; ww++) czl *= lds_f1v(SA_PW(ww)); }
            if (lw < w - 1) { czl2 = 1.f; for (int ww = lw + 1; ww < w - 1; ww++) czl2 *= lds_f1v(SA_PW(ww)); }
            if (lw < NW) {
                float pr = 1.f;
                for (int w2 = lw + 1; w2 < NW; w2++) { KKl = fmaf(lds_f1v(SA_KW(w2)), pr, KKl); pr *= lds_f1v(SA_PW(w2)); }
            }
            Aprev = (w > 0) ? lds_f1v(SA_AC(w - 1)) : 0.f;
        }
        float sM[C], sI[C], sD[C], dl[C];
#pragma unroll
        for (int c = 0; c < C; c++) { sM[c] = 0.f; sI[c] = 0.f; sD[c] = 0.f; dl[c] = 0.f; }
        float xN = 1.f, xB = pmove, xE = 0.f, xJ = 0.f, xC = 0.f, yex = 0.f, dL = 0.f;
        int sF = 0;
        if (tid == 0) { Fs[0] = xN; Fs[1] = xB; Fs[2] = 0.f; Fs[3] = 0.f; Fs[4] = 0.f; Fs[5] = 0.f; }
        __syncthreads();  // emis_s ready
        int xres = qd[0];
        float4 *fsrow = reinterpret_cast<float4 *>(Fs + 8);  // row i-1 is written while row i is computed
        PIN64(fsrow);
        // finish row r (after its barrier): D column values, E(r) and the special states; returns the rescale factor
        auto finish_row = [&](const int r) -> float {
            const unsigned par = (r & 1) * 128;
            const float tv = lds_f1v(SA_TOT2(par, lw)), ev = lds_f1v(SA_ES2(par, lw));
            float z = tv * czl, z2 = tv * czl2, et = fmaf(tv, KKl, ev);   // (slots >= NW hold zeros)
#pragma unroll
            for (int o = CW / 2; o > 0; o >>= 1) {
                z += __shfl_xor_sync(0xffffffffu, z, o);
                z2 += __shfl_xor_sync(0xffffffffu, z2, o);
                et += __shfl_xor_sync(0xffffffffu, et, o);
            }
            const float X = fmaf(Cexcl, z, yex);
#pragma unroll
            for (int c = 0; c < C; c++) sD[c] = fmaf(pDD[c], X, dl[c]);
            dL = __shfl_up_sync(0xffffffffu, sD[C - 1], 1);
            if (lane == 0) dL = (w > 0) ? fmaf(Aprev, z2, lds_f1v(SA_BV2(par, w - 1))) : 0.f;
            const float Et = et;
            xE = Et; xJ = xJ * ploop + Et * EJ; xC = xC * ploop + Et * EC; xN *= ploop; xB = (xN + xJ) * pmove;
            float scl = 1.f;
            if (Et > 1.0e12f) {
                int e = fexp(Et);
                scl = pow2i(-e); sF += e;
                xE *= scl; xJ *= scl; xC *= scl; xN *= scl; xB *= scl; dL *= scl;
#pragma unroll
                for (int c = 0; c < C; c++) { sM[c] *= scl; sI[c] *= scl; sD[c] *= scl; }
            }
            return scl;
        };
        for (int i = 1; i <= L; i++) {
            float scl = 1.f;
            if (i > 1) {
                scl = finish_row(i - 1);
                if (tid == 0) {
                    fsrow[0] = make_float4(xN, xB, xE, xJ); fsrow[1] = make_float4(xC, (float)sF, 0.f, 0.f);
                }
                fsrow += 2;
            }
            // row i
            float e[C];
            {
                const unsigned ea = emis_ta + xres * erow_b;
#pragma unroll
                for (int v = 0; v < C / 4; v++) {
                    const float4 t4 = lds_f4(ea + v * estep);
                    e[4 * v] = t4.x; e[4 * v + 1] = t4.y; e[4 * v + 2] = t4.z; e[4 * v + 3] = t4.w;
                }
            }
            if (i < L) xres = qd[i];
            float mL = __shfl_up_sync(0xffffffffu, sM[C - 1], 1);
            float iL = __shfl_up_sync(0xffffffffu, sI[C - 1], 1);
            if (lane == 0) {
                if (w > 0 && i > 1) {
                    const unsigned par = ((i - 1) & 1) * 128;
                    mL = lds_f1v(SA_BM2(par, w - 1)) * scl; iL = lds_f1v(SA_BI2(par, w - 1)) * scl;
                } else { mL = 0.f; iL = 0.f; }
            }
            float nM[C], nI[C];
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                float pm = c > 0 ? sM[c - 1] : mL, pi = c > 0 ? sI[c - 1] : iL, pd = c > 0 ? sD[c - 1] : dL;
                nI[c] = sM[c] * pmi[c] + sI[c] * pii[c];
                float acc = xB * pen[c];
                acc = fmaf(pm, pa[c], acc); acc = fmaf(pi, pb[c], acc); acc = fmaf(pd, pg[c], acc);
                nM[c] = acc * e[c];
            }
            dl[0] = 0.f;
#pragma unroll
            for (int c = 1; c < C; c++) dl[c] = fmaf(nM[c - 1], pmd[c], dl[c - 1] * pdd[c]);
            const float yl = fmaf(nM[C - 1], mdo, dl[C - 1] * ddo);   // local output of the thread's D chain
            float at = 0.f;
#pragma unroll
            for (int c = 0; c < C; c++) { sM[c] = nM[c]; sI[c] = nI[c]; at += nM[c] + dl[c]; }
            float y = yl, v = fmaf(yl, Rt, at);
#pragma unroll
            for (int s = 0; s < 5; s++) {
                const float up = __shfl_up_sync(0xffffffffu, y, 1 << s);
                v += __shfl_xor_sync(0xffffffffu, v, 1 << s);
                y = fmaf(coef[s], up, y);
            }
            yex = __shfl_up_sync(0xffffffffu, y, 1);
            if (lane == 0) yex = 0.f;
            {
                const unsigned par = (i & 1) * 128;
                if (lane == 31) {
                    sts_f1(SA_TOT2(par, w), y);
                    sts_f1(SA_BM2(par, w), nM[C - 1]);
                    sts_f1(SA_BI2(par, w), nI[C - 1]);
                    sts_f1(SA_BV2(par, w), fmaf(pDD[C - 1], yex, dl[C - 1]));
                }
                if (lane == 0) sts_f1(SA_ES2(par, w), v);
            }
            __syncthreads();
        }
        finish_row(L);
        if (tid == 0) {
            float4 *r = reinterpret_cast<float4 *>(Fs + 8 * L);
            r[0] = make_float4(xN, xB, xE, xJ); r[1] = make_float4(xC, (float)sF, 0.f, 0.f);
        }
        const float Tm = xC * pmove;  // total = Tm * 2^sF
        const int sT = sF;
        const float fwd_bits = log2f(Tm) + (float)sT;

        // =========================== Backward ===========================
        // Two barriers per row. B_b(i) = sum_k Mb(i+1,k) e_k entry_k and the D->D scan are both functions of row
        // i+1 only once the exit injection E_b(i) is split off by linearity: D_b(i,k) = D'(k) + E_b(i) * gD[k], with D'
        // the chain driven by the match terms alone and gD[k] = 1 + tDD[k] gD[k+1] a model constant. So the B sum and
        // the scan of D' run side by side before the barrier; after it E_b(i), the scan total and the finished row
        // follow, and a second barrier hands the first-column Mb to the left neighbour warp.
        // parameters leaving the owned columns
        load_cols<C>(E.tMM + po, k0 + 1, pa);
        load_cols<C>(E.tIM + po, k0 + 1, pb);
        load_cols<C>(E.tDM + po, k0 + 1, pg);
        load_cols<C>(E.tMD + po, k0 + 1, pmd);
        load_cols<C>(E.tDD + po, k0 + 1, pdd);
        // pmi, pii, pen already hold tMI, tII, entry of the owned columns
        const float GX = __ldg(E.gD + po + k0 + C + 1);   // D response of the right neighbour's first column to a unit exit
        __syncthreads();  // everyone done reading the forward pass's reduction rows
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
            if (lane == 0) sts_f1(SA_PW(w), Pc);
        }
        __syncthreads();
        czl = 0.f;  // Z_w = sum_{l > w} tot[l] * prod_{w < w'' < l} PW[w'']   (lane lw stands for warp lw)
        if (lw > w && lw < NW) {
            czl = 1.f;
            for (int ww = w + 1; ww < lw; ww++) czl *= lds_f1v(SA_PW(ww));
        }
#pragma unroll
        for (int c = 0; c < C; c++) { sM[c] = 0.f; sI[c] = 0.f; sD[c] = 0.f; }
        float bN = 0.f, bJ = 0.f, bC = 0.f, bE = 0.f;
        int sB = 0;
        const float invT = 1.0f / Tm;
        float4 *bsrow = reinterpret_cast<float4 *>(Bs + 8 * L);
        PIN64(bsrow);
        for (int i = L; i >= 0; i--) {
            // ---- before the barrier: everything that depends on row i+1 only ----
            const unsigned par = (i & 1) * 128;
            float mn[C], mnR[C];
            float eR = 0.f;
            if (i < L) {
                const int xr = qd[i];
                float e[C];
                const unsigned ea = emis_ta + xr * erow_b;
#pragma unroll
                for (int v = 0; v < C / 4; v++) {
                    const float4 t4 = lds_f4(ea + v * estep);
                    e[4 * v] = t4.x; e[4 * v + 1] = t4.y; e[4 * v + 2] = t4.z; e[4 * v + 3] = t4.w;
                }
                if (tid + 1 < T) eR = lds_f1(ea + 16);  // first column of the right neighbour
#pragma unroll
                for (int c = 0; c < C; c++) mn[c] = sM[c] * e[c];
            } else {
#pragma unroll
                for (int c = 0; c < C; c++) mn[c] = 0.f;
            }
            float nb = __shfl_down_sync(0xffffffffu, mn[0], 1);
            if (lane == 31) nb = (w + 1 < NW && i < L) ? lds_f1v(SA_BM((i + 1) & 1, w + 1)) * eR : 0.f;
            float bp = 0.f;
#pragma unroll
            for (int c = 0; c < C; c++) {
                mnR[c] = (c < C - 1) ? mn[c + 1] : nb;
                bp = fmaf(mn[c], pen[c], bp);
            }
            float Mp[C], nI[C], tm[C];
#pragma unroll
            for (int c = 0; c < C; c++) {
                Mp[c] = fmaf(mnR[c], pa[c], sI[c] * pmi[c]);
                nI[c] = fmaf(mnR[c], pb[c], sI[c] * pii[c]);
                tm[c] = mnR[c] * pg[c];
            }
            float y = tm[C - 1];   // D' chain (match terms only); its output at the thread's first column feeds the scan
#pragma unroll
            for (int c = C - 2; c >= 0; c--) y = fmaf(y, pdd[c], tm[c]);
#pragma unroll
            for (int s = 0; s < 5; s++) {   // warp reduction of the B partial and the D' scan side by side
                const float dn = __shfl_down_sync(0xffffffffu, y, 1 << s);
                bp += __shfl_xor_sync(0xffffffffu, bp, 1 << s);
                y = fmaf(coef[s], dn, y);
            }
            float yex = __shfl_down_sync(0xffffffffu, y, 1);
            if (lane == 31) yex = 0.f;
            if (lane == 0) { sts_f1(SA_ES2(par, w), bp); sts_f1(SA_TOT2(par, w), y); }
            __syncthreads();
            // ---- after it: E_b(i), the scan total, the finished row ----
            float Bi = lds_f1v(SA_ES2(par, lw)), Z = lds_f1v(SA_TOT2(par, lw)) * czl;   // (slots >= NW hold zeros)
#pragma unroll
            for (int o = CW / 2; o > 0; o >>= 1) {
                Bi += __shfl_xor_sync(0xffffffffu, Bi, o);
                Z += __shfl_xor_sync(0xffffffffu, Z, o);
            }
            if (i == L) { bC = pmove; bJ = 0.f; bN = 0.f; }
            else { bJ = bJ * ploop + Bi * pmove; bC = bC * ploop; bN = bN * ploop + Bi * pmove; }
            bE = bJ * EJ + bC * EC;
            float Xp = fmaf(Cexcl, Z, yex);   // D' entering from the right neighbour's first column
            {
                float big = fmaxf(fmaxf(bN, bJ), Bi);
                if (big > 1.0e9f) {
                    int e = fexp(big);
                    float scl = pow2i(-e);
                    sB += e;
                    bN *= scl; bJ *= scl; bC *= scl; bE *= scl; Bi *= scl; Xp *= scl;
#pragma unroll
                    for (int c = 0; c < C; c++) { Mp[c] *= scl; nI[c] *= scl; tm[c] *= scl; }
                }
            }
            if (tid == 0) {  // backward specials of row i, decoded after the sweep
                bsrow[0] = make_float4(Bi, bE, bN, bJ); bsrow[1] = make_float4(bC, (float)sB, 0.f, 0.f);
            }
            bsrow -= 2;
            if (i == 0) break;
            const float X = fmaf(bE, GX, Xp);  // D(i, first column of the right neighbour)
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                const float dr = (c < C - 1) ? sD[c + 1] : X;
                sD[c] = fmaf(pdd[c], dr, tm[c] + bE);
                sM[c] = fmaf(pmd[c], dr, Mp[c] + bE);
                sI[c] = nI[c];
            }
            if (lane == 0) sts_f1(SA_BM(i & 1, w), sM[0]);
            __syncthreads();
        }
        if (Wk.dbg_bwd != nullptr && tid == 0)
            Wk.dbg_bwd[(size_t)q * E.H + h] = (logf(bN) + (float)sB * 0.69314718056f);

        // ---- posterior decoding of the special states, parallel over rows ----
        __syncthreads();
        for (int i = tid; i <= L; i += T) {
            const float4 f1a = reinterpret_cast<const float4 *>(Fs + 8 * i)[0], f1b = reinterpret_cast<const float4 *>(Fs + 8 * i)[1];
            const float4 b1a = reinterpret_cast<const float4 *>(Bs + 8 * i)[0], b1b = reinterpret_cast<const float4 *>(Bs + 8 * i)[1];
            const float fii = exp2f(f1b.y + b1b.y - (float)sT) * invT;
            dPB[i] = f1a.y * b1a.x * fii;   // F_B(i) * B_B(i) / T
            dPE[i] = f1a.z * b1a.y * fii;   // F_E(i) * B_E(i) / T
            float mo = 0.f;
            if (i > 0) {
                const float4 f0a = reinterpret_cast<const float4 *>(Fs + 8 * (i - 1))[0], f0b = reinterpret_cast<const float4 *>(Fs + 8 * (i - 1))[1];
                const float fpi = exp2f(f0b.y + b1b.y - (float)sT) * invT * ploop;
                mo = 1.0f - (f0a.x * b1a.z + f0a.w * b1a.w + f0b.x * b1b.x) * fpi;   // N, J, C
            }
            dMO[i] = mo;
        }
        // =========================== regions (warp 0) ===========================
        __syncthreads();
        if (w == 0) {
            const float rt1 = 0.25f, rt2 = 0.10f, rt3 = 0.20f;
            // prefix sums btot[i] = sum_{i'<i} P(B at i'), etot[i] = sum_{i'<=i} P(E at i')
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
                    const float m = __shfl_sync(0xffffffffu, mo, z), b = __shfl_sync(0xffffffffu, db, z),
                                ee = __shfl_sync(0xffffffffu, de, z);
                    const int j = base + z;
                    if (!trig) {
                        if (m - b < rt2) i0 = j; else if (i0 == -1) i0 = j;
                        if (m >= rt1) trig = 1;
                    } else if (m - ee < rt2) {
                        // region i0..j : single- or multi-domain?
                        float mx = -1.f;
                        const float eb = dET[i0 - 1], bj = dBT[j];
                        for (int zz = i0 + lane; zz <= j; zz += 32)
                            mx = fmaxf(mx, fminf(dET[zz] - eb, bj - dBT[zz - 1]));
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                        if (mx >= rt3) flags |= 1 | (nenv < MAX_ENV ? (256 << nenv) : 0);   // bit 8+r: region r is multi-domain
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
}

}  // namespace witch
