// Family W, envelope mode, TWO ITEMS PER WARP (build-time option WITCH_WAVE_PAIR=1; not the default: DESIGN.md section 9).
//
// Same wavefront as wave_kernel<C, false, ...> (wave_kernels.cuh), but a warp owns two (query, envelope) items of one
// HMM and keeps item A in the .x half and item B in the .y half of every f32x2 register pair, so that the recurrence
// issues packed FFMA2 / FMUL2 / FADD2 (half the FP issue slots; tools/proto shows ptxas keeps such pairs packed without
// re-pairing moves). Transition parameters are shared by the two items (duplicated pairs), so a lane owns C = 4 columns
// and a strip is 128 columns wide. Stored Forward rows interleave the items ([step][quad][lane][2 columns x 2 items]: a 128-bit access
// moves two whole pairs), strip-boundary records hold both items (64 B per row), scaling exponents / rescaling /
// posterior scale are per item. Rows are aligned by index: row i of both items is processed in the same step. The longer
// item sets the number of steps (Ls = max); for the shorter one the Forward total is captured at its own last row, its
// residues are padded with its last residue (rows beyond its end are computed and never used), and its Backward sweep
// starts when its exit injection is switched on at its own last row (before that every Backward value of it is an exact 0).
// Envelope mode needs no per-row special-state arrays (N/C posteriors are an align-stage matter).
#pragma once
#include "wave_kernels.cuh"

namespace witch {

#ifdef WITCH_HOST_SIM
__device__ __forceinline__ float2 p_fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 p_mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 p_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
#else
__device__ __forceinline__ float2 p_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 p_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 p_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
#endif
__device__ __forceinline__ float2 p_mk(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 p_dup(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 p_up(float2 v) { return p_mk(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1)); }
__device__ __forceinline__ float2 p_down(float2 v) { return p_mk(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1)); }
__device__ __forceinline__ float2 p_sel(bool p, float2 a, float2 b) { return p ? a : b; }
__device__ __forceinline__ float2 p_max(float2 a, float2 b) { return p_mk(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
struct PairI { int x, y; };

constexpr int WP_C = 4;                 // columns per lane
constexpr int WP_SW = 32 * WP_C;        // strip width
constexpr int WP_STAGE = 32 * WP_C * 2 * 4;   // bytes of one step's stored rows (both items)
constexpr int WP_BND_REC = 16;          // words per boundary record: M.xy I.xy D.xy E.xy g.xy + 6 pad
constexpr int WP_BND_BLK = 8 * WP_BND_REC * 4;   // bytes of one 8-row block of records

struct WavePairLayout { long long tile, gF, bnd, total; int TT, TG; };
__host__ __device__ inline WavePairLayout wave_pair_layout(int Lcap, int max_strips) {
    WavePairLayout w;
    w.TT = Lcap + 32;
    w.TG = w.TT / 8 + 2;
    long long o = 0;
    w.tile = o; o += (long long)max_strips * w.TT * WP_STAGE;
    w.gF = o; o += ((long long)max_strips * w.TG * 8 + 15) / 16 * 16;
    w.bnd = o; o += (long long)(Lcap + 16) * WP_BND_REC * 4;
    w.total = (o + 255) / 256 * 256;
    return w;
}
__host__ __device__ constexpr int wave_pair_smem_per_warp(int ring, int res_cap) {
    return 2 * res_cap + ring * (WP_STAGE + 8) + W_BND_SLOTS * (WP_BND_BLK + 8) + WP_BND_REC * 4;   // + one all-zero record
}

template <int WAVE_WARPS, int MINB, int W_RING>
__global__ void __launch_bounds__(WAVE_WARPS * 32, MINB) wave_pair_kernel(DevEhmm E, DevQueries Q, WaveWork Wk) {
    WITCH_DYN_SMEM(float, smem);
    constexpr int C = WP_C, SW = WP_SW;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __shared__ int s_group;
    __shared__ float s_n2[WAVE_WARPS][2][MAX_SYM];
    float *emis_s = smem;  // [nsym][Mstr], plain column order (C = 4: a lane's quad is contiguous)
    const WavePairLayout lay = wave_pair_layout(Wk.Lcap, Wk.max_strips);
    __shared__ char *s_slot[WAVE_WARPS];
    if (lane == 0) s_slot[w] = Wk.scratch + ((long long)blockIdx.x * WAVE_WARPS + w) * Wk.slot_bytes;
    __syncwarp();
    char *slot = *((char *volatile *)&s_slot[w]);
    __builtin_assume(__isGlobal(slot));
    float *tile = (float *)(slot + lay.tile);
    int *gFarr = (int *)(slot + lay.gF);
    float *bnd = (float *)(slot + lay.bnd);
    const int TT = lay.TT, TG = lay.TG;
    const unsigned emis_sa = smem_u32(emis_s);
    const unsigned FULL = 0xffffffffu;

    int rd_stage = 0, wr_stage = 0;
    unsigned rd_phase = 0;
    int bq_w = 0, bq_r = 0;
    unsigned bq_ph = 0;
    // dynamic shared memory after the emission table: residues (2 per warp) | Forward-row ring | its mbarriers | boundary ring | its mbarriers
    const unsigned sm_dyn = emis_sa + Wk.emis_floats * 4;
    unsigned ring_w = sm_dyn + WAVE_WARPS * 2 * Wk.res_cap + w * (W_RING * WP_STAGE);
    unsigned ring_sa = ring_w + lane * 16;   // stage layout [quad v][lane][4 floats]: conflict-free LDS.128
    unsigned ring_bar = sm_dyn + WAVE_WARPS * 2 * Wk.res_cap + WAVE_WARPS * (W_RING * WP_STAGE) + w * (W_RING * 8);
    unsigned bnd_ring = sm_dyn + WAVE_WARPS * 2 * Wk.res_cap + WAVE_WARPS * (W_RING * (WP_STAGE + 8)) + w * (W_BND_SLOTS * WP_BND_BLK);
    unsigned bnd_bar = sm_dyn + WAVE_WARPS * 2 * Wk.res_cap + WAVE_WARPS * (W_RING * (WP_STAGE + 8) + W_BND_SLOTS * WP_BND_BLK) + w * (W_BND_SLOTS * 8);
    // an all-zero boundary record: what the first strip (Forward) / the last strip (Backward) reads instead of a neighbour's
    unsigned zrec = sm_dyn + WAVE_WARPS * 2 * Wk.res_cap + WAVE_WARPS * (W_RING * (WP_STAGE + 8) + W_BND_SLOTS * (WP_BND_BLK + 8)) + w * (WP_BND_REC * 4);
    PIN32(ring_w); PIN32(ring_sa); PIN32(ring_bar); PIN32(bnd_ring); PIN32(bnd_bar); PIN32(zrec);   // one register each, no re-derivation in the loops
    if (lane < WP_BND_REC) sts_f1(zrec + lane * 4, 0.f);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < W_RING; k++) mbar_init(ring_bar + k * 8, 1);
#pragma unroll
        for (int k = 0; k < W_BND_SLOTS; k++) mbar_init(bnd_bar + k * 8, 1);
        fence_mbar_init();
    }
    __syncwarp();
    auto bnd_issue = [&](const int b) {   // block b = records of rows 8b .. 8b+7
        if ((W_EXP & 1) ? wave_elect_one() : (lane == 0)) {
            mbar_expect_tx(bnd_bar + bq_w * 8, WP_BND_BLK);
            tma_load_1d(bnd_ring + bq_w * WP_BND_BLK, bnd + 8 * WP_BND_REC * b, WP_BND_BLK, bnd_bar + bq_w * 8);
        }
        bq_w = (bq_w == W_BND_SLOTS - 1) ? 0 : bq_w + 1;
    };
    auto bnd_wait = [&]() {
        mbar_wait(bnd_bar + bq_r * 8, bq_ph);
        if (bq_r == W_BND_SLOTS - 1) { bq_r = 0; bq_ph ^= 1u; } else bq_r++;
    };
    auto bnd_next = [](const int sl) { return sl == W_BND_SLOTS - 1 ? 0 : sl + 1; };
    int loaded_h = -1, Mstr = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_group = (int)atomicAdd(Wk.counter, 1u);
        __syncthreads();
        const int grp = s_group;
        if (grp >= Wk.ngroups) break;
        const int gfirst = Wk.group_first[grp], gcount = Wk.group_count[grp];
        const int h = Wk.items[gfirst].h;
        const int Mh = E.M[h];
        const int nstrips = (Mh + SW - 1) / SW;
        if (h != loaded_h) {
            Mstr = nstrips * SW;
            const int st = E.stride[h];
            const float *eg = E.emis + E.eoff[h];
            for (int idx = threadIdx.x; idx < Q.nsym * Mstr; idx += blockDim.x) {
                const int x = idx / Mstr, col = idx - x * Mstr;
                emis_s[idx] = (col < st - 1) ? __ldg(eg + (size_t)Q.symrow[x] * st + 1 + col) : 0.f;
            }
            loaded_h = h;
        }
        __syncthreads();
        if (2 * w >= gcount) continue;
        const bool hasB = (2 * w + 1 < gcount);
        const WaveItem itA = Wk.items[gfirst + 2 * w], itB = Wk.items[gfirst + 2 * w + (hasB ? 1 : 0)];
        const PairI Ls2 = {itA.Ls, itB.Ls};
        const int Ls = max(Ls2.x, Ls2.y);
        const int dL = Ls - min(Ls2.x, Ls2.y);
        // residues of both items in shared memory, the shorter one padded with its last residue up to Ls
        unsigned sresA = emis_sa + Wk.emis_floats * 4 + (2 * w) * Wk.res_cap, sresB = sresA + Wk.res_cap;
        PIN32(sresA); PIN32(sresB);
        {
            uint8_t *sr = reinterpret_cast<uint8_t *>(emis_s + Wk.emis_floats) + (2 * w) * Wk.res_cap;
            const uint8_t *da = Q.dsq + Q.off[itA.q] + (itA.i0 - 1), *db = Q.dsq + Q.off[itB.q] + (itB.i0 - 1);
            for (int z = lane; z < Ls; z += 32) {
                sr[z] = da[min(z, Ls2.x - 1)];
                sr[Wk.res_cap + z] = db[min(z, Ls2.y - 1)];
            }
            __syncwarp();
        }
        const long long po = E.poff[h];
        const float2 pmove = p_mk(2.0f / ((float)Q.len[itA.q] + 2.0f), 2.0f / ((float)Q.len[itB.q] + 2.0f));
        const float2 ploop = p_mk(1.0f - pmove.x, 1.0f - pmove.y);
        const int nsteps = Ls + 31;
        unsigned erow = Mstr * 4;
        PIN32(erow);

        // ======================================= Forward =======================================
        float2 xCv = p_dup(0.f), xCfin = p_dup(0.f);
        PairI xCg = {0, 0}, xCgfin = {0, 0};
        for (int s = 0; s < nstrips; s++) {
            const long long k0 = (long long)s * SW + lane * C;  // owns columns k0+1..k0+C
            float2 pa[C], pb[C], pg[C], pmd[C], pdd[C], pmi[C], pii[C], pen[C];
            {
                float t[C];
                load_cols<C>(E.tMM + po, k0, t);
#pragma unroll
                for (int c = 0; c < C; c++) pa[c] = p_dup(t[c]);
                load_cols<C>(E.tIM + po, k0, t);
#pragma unroll
                for (int c = 0; c < C; c++) pb[c] = p_dup(t[c]);
                load_cols<C>(E.tDM + po, k0, t);
#pragma unroll
                for (int c = 0; c < C; c++) pg[c] = p_dup(t[c]);
                load_cols<C>(E.tMD + po, k0, t);
#pragma unroll
                for (int c = 0; c < C; c++) pmd[c] = p_dup(t[c]);
                load_cols<C>(E.tDD + po, k0, t);
#pragma unroll
                for (int c = 0; c < C; c++) pdd[c] = p_dup(t[c]);
                load_cols<C>(E.tMI + po, k0 + 1, t);
#pragma unroll
                for (int c = 0; c < C; c++) pmi[c] = p_dup(t[c]);
                load_cols<C>(E.tII + po, k0 + 1, t);
#pragma unroll
                for (int c = 0; c < C; c++) pii[c] = p_dup(t[c]);
                load_cols<C>(E.entry + po, k0 + 1, t);
#pragma unroll
                for (int c = 0; c < C; c++) pen[c] = p_dup(t[c]);
            }
            float2 sM[C], sI[C], sD[C];
#pragma unroll
            for (int c = 0; c < C; c++) { sM[c] = p_dup(0.f); sI[c] = p_dup(0.f); sD[c] = p_dup(0.f); }
            float2 rM = p_dup(0.f), rI = p_dup(0.f), rD = p_dup(0.f), ep = p_dup(0.f);
            PairI g = {0, 0};
            if (s > 0) { const int *gi = (const int *)(bnd + WP_BND_REC * 1 + 8); g.x = gi[0]; g.y = gi[1]; }
            float2 xBs = p_mk(pmove.x * pow2i(-g.x), pmove.y * pow2i(-g.y));   // pmove * N(i-1) * 2^-g for the lane's next row
            const bool last = (s == nstrips - 1);
            unsigned ebase = emis_sa + (s * SW + lane * C) * 4;
            float *tMp = tile + (size_t)s * TT * (SW * 2) + (SW * 2) + lane * 4;   // (step 1, quad 0, lane); step = [v][lane][4]
            int *gFs = gFarr + s * TG * 2;
            PIN32(ebase); PIN64(tMp); PIN64(gFs);
            if (last) { xCv = p_dup(0.f); xCg = g; }
            int xa = lds_u8(sresA + min(max(-lane, 0), Ls - 1)), xb = lds_u8(sresB + min(max(-lane, 0), Ls - 1));
            int bcur = 0;
            if (s > 0) {
                fence_proxy_async();
                __syncwarp();
                bcur = bq_r;
                const int bmax = Ls >> 3;
                bnd_issue(0);
                if (bmax >= 1) bnd_issue(1);
                if (bmax >= 2) bnd_issue(2);
                bnd_wait();
                if (bmax >= 1) bnd_wait();
            }
            auto fstep = [&](const int t, auto allc, auto uc) {
                constexpr bool ALL = decltype(allc)::value;
                constexpr int U = decltype(uc)::value;   // steady state: t = 8j + 1 + U, so t & 7 is a compile-time constant
                const int i = t - lane;
                const bool act = ALL || (i >= 1 && i <= Ls);
                float2 cM = p_up(sM[C - 1]), cI = p_up(sI[C - 1]), cD = p_up(sD[C - 1]), cE = p_up(ep);
                {   // lane 0: strip boundary of the left strip (first strip: the all-zero record); scalar, predicated
                    unsigned ra = zrec;
                    if (s > 0) {
                        const int tr = ALL ? t : min(t, Ls);
                        const int t7 = ALL ? ((1 + U) & 7) : (t & 7), tr7 = ALL ? ((1 + U) & 7) : (tr & 7);
                        if (t7 == 0 && (ALL || t <= Ls)) {   // row t opens boundary block t >> 3
                            const int b = t >> 3;
                            __syncwarp();
                            if (8 * (b + 2) <= Ls) bnd_issue(b + 2);
                            if (8 * (b + 1) <= Ls) bnd_wait();
                            bcur = bnd_next(bcur);
                        }
                        ra = bnd_ring + bcur * WP_BND_BLK + tr7 * (WP_BND_REC * 4);
                    }
                    const float4 r0 = lds_f4v(ra), r1 = lds_f4v(ra + 16);
                    float fx = pow2i(lds_i1v(ra + 32) - g.x), fy = pow2i(lds_i1v(ra + 36) - g.y);
                    if (!ALL && !act) { fx = 0.f; fy = 0.f; }
                    if (lane == 0) {
                        cM.x = r0.x * fx; cM.y = r0.y * fy; cI.x = r0.z * fx; cI.y = r0.w * fy;
                        cD.x = r1.x * fx; cD.y = r1.y * fy; cE.x = r1.z * fx; cE.y = r1.w * fy;
                    }
                }
                const int xra = xa, xrb = xb;
                {
                    const int ri = ALL ? i : min(max(i, 0), Ls - 1);
                    xa = lds_u8(sresA + ri); xb = lds_u8(sresB + ri);
                }
                if (act) {
                    const float4 eA = lds_f4(ebase + xra * erow), eB = lds_f4(ebase + xrb * erow);
                    const float ea[4] = {eA.x, eA.y, eA.z, eA.w}, eb[4] = {eB.x, eB.y, eB.z, eB.w};
                    float2 nM[C], nI[C], nD[C];
#pragma unroll
                    for (int c = C - 1; c >= 0; c--) {
                        const float2 pm = c > 0 ? sM[c - 1] : rM, pi = c > 0 ? sI[c - 1] : rI, pd = c > 0 ? sD[c - 1] : rD;
                        nI[c] = p_fma(sM[c], pmi[c], p_mul(sI[c], pii[c]));
                        float2 acc = p_mul(xBs, pen[c]);
                        acc = p_fma(pm, pa[c], acc); acc = p_fma(pi, pb[c], acc); acc = p_fma(pd, pg[c], acc);
                        nM[c] = p_mk(acc.x * ea[c], acc.y * eb[c]);
                    }
                    nD[0] = p_fma(cD, pdd[0], p_mul(cM, pmd[0]));
#pragma unroll
                    for (int c = 1; c < C; c++) nD[c] = p_fma(nD[c - 1], pdd[c], p_mul(nM[c - 1], pmd[c]));
                    float2 es = cE;
#pragma unroll
                    for (int c = 0; c < C; c++) { sM[c] = nM[c]; sI[c] = nI[c]; sD[c] = nD[c]; es = p_add(es, p_add(nM[c], nD[c])); }
                    ep = es;
                    rM = cM; rI = cI; rD = cD;
                    xBs = p_mul(xBs, ploop);
#pragma unroll
                    for (int v = 0; v < C / 2; v++)
                        *reinterpret_cast<float4 *>(tMp + 128 * v) = make_float4(nM[2 * v].x, nM[2 * v].y, nM[2 * v + 1].x, nM[2 * v + 1].y);
                    if (lane == 31) {
                        if (!last) {
                            float *rec = bnd + WP_BND_REC * i;
                            *reinterpret_cast<float4 *>(rec) = make_float4(sM[C - 1].x, sM[C - 1].y, sI[C - 1].x, sI[C - 1].y);
                            *reinterpret_cast<float4 *>(rec + 4) = make_float4(sD[C - 1].x, sD[C - 1].y, ep.x, ep.y);
                            ((int *)rec)[8] = g.x; ((int *)rec)[9] = g.y;
                        } else {
                            // C(i) = C(i-1)*loop + E(i)   (unihit: E->C = 1), with exponent alignment; each item's total is
                            // the value at its own last row
                            const float2 f = p_mk(pow2i(xCg.x - g.x), pow2i(xCg.y - g.y));
                            xCv = p_add(p_mul(p_mul(xCv, f), ploop), ep); xCg = g;
                            if (i == Ls2.x) { xCfin.x = xCv.x; xCgfin.x = g.x; }
                            if (i == Ls2.y) { xCfin.y = xCv.y; xCgfin.y = g.y; }
                        }
                    }
                }
                tMp += SW * 2;
            };
            auto fmark = [&](const int t) { if (lane == 0) { gFs[2 * ((t - 1) >> 3)] = g.x; gFs[2 * ((t - 1) >> 3) + 1] = g.y; } };
            auto frescale = [&](const int t) {
                float2 mx = p_dup(0.f);
#pragma unroll
                for (int c = 0; c < C; c++) mx = p_max(mx, p_max(sM[c], p_max(sI[c], sD[c])));
                mx = p_max(mx, ep);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    mx = p_max(mx, p_mk(__shfl_xor_sync(FULL, mx.x, o), __shfl_xor_sync(FULL, mx.y, o)));
                PairI e_need = {(mx.x > 1048576.f) ? fexp(mx.x) : 0, (mx.y > 1048576.f) ? fexp(mx.y) : 0};
                if (s > 0) {   // exponent the left strip had 8 rows ahead (that block is already in the ring)
                    const int r = min(Ls, t + W_SCALE_EVERY), tc = min(t, Ls);
                    const int sl = ((r >> 3) == (tc >> 3)) ? bcur : bnd_next(bcur);
                    const unsigned ra = bnd_ring + sl * WP_BND_BLK + (r & 7) * (WP_BND_REC * 4);
                    e_need.x = max(e_need.x, lds_i1v(ra + 32) - 40 - g.x);
                    e_need.y = max(e_need.y, lds_i1v(ra + 36) - 40 - g.y);
                }
                if (e_need.x > 0 || e_need.y > 0) {
                    e_need.x = max(e_need.x, 0); e_need.y = max(e_need.y, 0);
                    const float2 f = p_mk(pow2i(-e_need.x), pow2i(-e_need.y));
                    g.x += e_need.x; g.y += e_need.y;
#pragma unroll
                    for (int c = 0; c < C; c++) { sM[c] = p_mul(sM[c], f); sI[c] = p_mul(sI[c], f); sD[c] = p_mul(sD[c], f); }
                    rM = p_mul(rM, f); rI = p_mul(rI, f); rD = p_mul(rD, f); ep = p_mul(ep, f); xBs = p_mul(xBs, f);
                }
            };
            {
                const std::integral_constant<bool, false> genc;
                const std::integral_constant<bool, true> allc;
                int t = 1;
                const std::integral_constant<int, -1> nou;
                for (; t <= 32; t++) {
                    if (((t - 1) & 7) == 0) fmark(t);
                    fstep(t, genc, nou);
                    if ((t & (W_SCALE_EVERY - 1)) == 0) frescale(t);
                }
                for (; t + 7 <= Ls; t += 8) {   // t = 33, 41, ...
                    fmark(t);
                    fstep(t, allc, std::integral_constant<int, 0>()); fstep(t + 1, allc, std::integral_constant<int, 1>());
                    fstep(t + 2, allc, std::integral_constant<int, 2>()); fstep(t + 3, allc, std::integral_constant<int, 3>());
                    fstep(t + 4, allc, std::integral_constant<int, 4>()); fstep(t + 5, allc, std::integral_constant<int, 5>());
                    fstep(t + 6, allc, std::integral_constant<int, 6>()); fstep(t + 7, allc, std::integral_constant<int, 7>());
                    frescale(t + 7);
                }
                for (; t <= nsteps; t++) {
                    if (((t - 1) & 7) == 0) fmark(t);
                    fstep(t, genc, nou);
                    if ((t & (W_SCALE_EVERY - 1)) == 0) frescale(t);
                }
            }
            if (last) {
                xCfin.x = __shfl_sync(FULL, xCfin.x, 31); xCfin.y = __shfl_sync(FULL, xCfin.y, 31);
                xCgfin.x = __shfl_sync(FULL, xCgfin.x, 31); xCgfin.y = __shfl_sync(FULL, xCgfin.y, 31);
            }
            __syncwarp();
        }
        const float2 Tm = p_mul(xCfin, pmove);
        const PairI gT = xCgfin;  // P = Tm * 2^gT
        const float2 fwd_nats = p_mk(logf(Tm.x) + (float)gT.x * 0.69314718056f, logf(Tm.y) + (float)gT.y * 0.69314718056f);
        const float2 invT = p_mk(1.0f / Tm.x, 1.0f / Tm.y);

        // ======================================= Backward =======================================
        // lane l processes row i = Ls - (t' - (31 - l)), rows Ls .. 0 of BOTH items (row 0 only feeds the B special)
        float2 accI = p_dup(0.f);
        if (lane < MAX_SYM) { s_n2[w][0][lane] = 0.f; s_n2[w][1][lane] = 0.f; }
        __syncwarp();
        const int nstepsB = Ls + 1 + 31;
        for (int s = nstrips - 1; s >= 0; s--) {
            const long long k0 = (long long)s * SW + lane * C;
            float2 oMM[C], oIM[C], oDM[C], oMD[C], oDD[C], oMI[C], oII[C], pen[C];
            {
                float t[C];
                load_cols<C>(E.tMM + po, k0 + 1, t);
#pragma unroll
                for (int c = 0; c < C; c++) oMM[c] = p_dup(t[c]);
                load_cols<C>(E.tIM + po, k0 + 1, t);
#pragma unroll
                for (int c = 0; c < C; c++) oIM[c] = p_dup(t[c]);
                load_cols<C>(E.tDM + po, k0 + 1, t);
#pragma unroll
                for (int c = 0; c < C; c++) oDM[c] = p_dup(t[c]);
                load_cols<C>(E.tMD + po, k0 + 1, t);
#pragma unroll
                for (int c = 0; c < C; c++) oMD[c] = p_dup(t[c]);
                load_cols<C>(E.tDD + po, k0 + 1, t);
#pragma unroll
                for (int c = 0; c < C; c++) oDD[c] = p_dup(t[c]);
                load_cols<C>(E.tMI + po, k0 + 1, t);
#pragma unroll
                for (int c = 0; c < C; c++) oMI[c] = p_dup(t[c]);
                load_cols<C>(E.tII + po, k0 + 1, t);
#pragma unroll
                for (int c = 0; c < C; c++) oII[c] = p_dup(t[c]);
                load_cols<C>(E.entry + po, k0 + 1, t);
#pragma unroll
                for (int c = 0; c < C; c++) pen[c] = p_dup(t[c]);
            }
            float2 sM[C], sI[C], sD[C], accM[C];
#pragma unroll
            for (int c = 0; c < C; c++) { sM[c] = p_dup(0.f); sI[c] = p_dup(0.f); sD[c] = p_dup(0.f); accM[c] = p_dup(0.f); }
            float2 rMb = p_dup(0.f), bp = p_dup(0.f);
            const bool lastS = (s == nstrips - 1), firstS = (s == 0);
            PairI g = {0, 0};
            if (!lastS) { const int *gi = (const int *)(bnd + WP_BND_REC * Ls + 8); g.x = gi[0]; g.y = gi[1]; }
            float2 ebs = p_dup(0.f);   // E(i) = C_b(i) = pmove * loop^(Ls_item - i), scaled; switched on at the item's own last row
            unsigned ebase = emis_sa + (s * SW + lane * C) * 4;
            const bool hasR = !(lastS && lane == 31);
            float *tM = tile + (size_t)s * TT * (SW * 2);
            const int *gFs = gFarr + s * TG * 2;
            PIN32(ebase); PIN64(gFs);
            int bcur = 0;
            int gblk = (Ls + 30) >> 3;
            PairI gFc = {gFs[2 * gblk], gFs[2 * gblk + 1]};
            PairI gFnext = {gFs[2 * max(gblk - 1, 0)], gFs[2 * max(gblk - 1, 0) + 1]};
            auto post_scale = [&](const PairI gf, const PairI gb) {
                return p_mk(exp2f((float)(gf.x + gb.x - gT.x)) * invT.x, exp2f((float)(gf.y + gb.y - gT.y)) * invT.y);
            };
            float2 fac = post_scale(gFc, g);
            int xa = lds_u8(sresA + Ls - 1), xb = lds_u8(sresB + Ls - 1);   // residue i+1 of the lane's row at the next step
            const float *tMrd = tM + (size_t)(Ls + 31) * (SW * 2);   // rows of step 0; step tq is SW*2 floats earlier
            PIN64(tMrd);
            int tq_next = 0;
            auto ring_issue = [&]() {
                if ((W_EXP & 1) ? wave_elect_one() : (lane == 0)) {
                    const unsigned dst = ring_w + wr_stage * WP_STAGE, bar = ring_bar + wr_stage * 8;
                    mbar_expect_tx(bar, WP_STAGE);
                    tma_load_1d(dst, tMrd, WP_STAGE, bar);
                }
                tMrd -= SW * 2;
                wr_stage = (wr_stage == W_RING - 1) ? 0 : wr_stage + 1;
                tq_next++;
            };
            fence_proxy_async();
            __syncwarp();
#pragma unroll
            for (int k = 0; k < W_RING - 1; k++) ring_issue();
            if (!lastS) {
                bcur = bq_r;
                const int b0 = Ls >> 3;
                bnd_issue(b0);
                if (b0 >= 1) bnd_issue(b0 - 1);
                if (b0 >= 2) bnd_issue(b0 - 2);
                bnd_wait();
                if (b0 >= 1) bnd_wait();
            }
            auto bstep = [&](const int tp, auto allc) {
                constexpr bool ALL = decltype(allc)::value;   // every lane has 1 <= i < min(Ls_A, Ls_B)
                const int i = Ls - (tp - (31 - lane));
                const bool act = ALL || (i >= 0 && i <= Ls);
                float2 cMb = p_down(sM[0]), cDb = p_down(sD[0]), cB = p_down(bp);
                {   // lane 31: boundary of the strip to the right (last strip: the all-zero record); scalar, predicated
                    unsigned ra = zrec;
                    if (!lastS) {
                        const int i31 = ALL ? Ls - tp : max(Ls - tp, 0);
                        if ((i31 & 7) == 7 && tp > 0 && (ALL || tp <= Ls)) {
                            const int b = i31 >> 3;
                            __syncwarp();
                            if (b >= 2) bnd_issue(b - 2);
                            if (b >= 1) bnd_wait();
                            bcur = bnd_next(bcur);
                        }
                        ra = bnd_ring + bcur * WP_BND_BLK + (i31 & 7) * (WP_BND_REC * 4);
                    }
                    const float4 r0 = lds_f4v(ra), r1 = lds_f4v(ra + 16);
                    float fx = pow2i(lds_i1v(ra + 32) - g.x), fy = pow2i(lds_i1v(ra + 36) - g.y);
                    if (!ALL && !act) { fx = 0.f; fy = 0.f; }
                    if (lane == 31) {
                        cMb.x = r0.x * fx; cMb.y = r0.y * fy; cDb.x = r1.x * fx; cDb.y = r1.y * fy; cB.x = r1.z * fx; cB.y = r1.w * fy;
                    }
                }
                __syncwarp();
                if (ALL || tq_next < nstepsB) ring_issue();
                mbar_wait(ring_bar + rd_stage * 8, rd_phase);
                const unsigned rs = ring_sa + rd_stage * WP_STAGE;
                if (rd_stage == W_RING - 1) { rd_stage = 0; rd_phase ^= 1u; } else rd_stage++;
                const int xra = xa, xrb = xb;
                {
                    const int ri = ALL ? i - 1 : min(max(i - 1, 0), Ls - 1);
                    xa = lds_u8(sresA + ri); xb = lds_u8(sresB + ri);
                }
                if (act) {
                    if (!ALL) {   // the exit injection of an item starts at its own last row (its state is exactly 0 before)
                        if (i == Ls2.x) ebs.x = pmove.x * pow2i(-g.x);
                        if (i == Ls2.y) ebs.y = pmove.y * pow2i(-g.y);
                    }
                    float2 mn[C], mnR;
                    if (ALL || i < Ls) {
                        const float4 eA = lds_f4(ebase + xra * erow), eB = lds_f4(ebase + xrb * erow);
                        const float ea[4] = {eA.x, eA.y, eA.z, eA.w}, eb[4] = {eB.x, eB.y, eB.z, eB.w};
#pragma unroll
                        for (int c = 0; c < C; c++) mn[c] = p_mk(sM[c].x * ea[c], sM[c].y * eb[c]);
                        mnR = hasR ? p_mk(rMb.x * lds_f1(ebase + 16 + xra * erow), rMb.y * lds_f1(ebase + 16 + xrb * erow)) : p_dup(0.f);
                    } else {
#pragma unroll
                        for (int c = 0; c < C; c++) mn[c] = p_dup(0.f);
                        mnR = p_dup(0.f);
                    }
                    float2 bs = cB;
#pragma unroll
                    for (int c = 0; c < C; c++) bs = p_fma(mn[c], pen[c], bs);
                    bp = bs;
                    if (ALL || i >= 1) {
                        float2 nM[C], nI[C], nD[C];
#pragma unroll
                        for (int c = C - 1; c >= 0; c--) {
                            const float2 m1 = (c < C - 1) ? mn[c + 1] : mnR;
                            const float2 dr = (c < C - 1) ? nD[c + 1] : cDb;
                            nD[c] = p_fma(dr, oDD[c], p_fma(m1, oDM[c], ebs));
                            nM[c] = p_fma(m1, oMM[c], p_fma(sI[c], oMI[c], p_fma(dr, oMD[c], ebs)));
                            nI[c] = p_fma(m1, oIM[c], p_mul(sI[c], oII[c]));
                        }
#pragma unroll
                        for (int v = 0; v < C / 2; v++) {
                            const float4 a = lds_f4v(rs + v * 512);
                            accM[2 * v] = p_fma(p_mul(p_mk(a.x, a.y), nM[2 * v]), fac, accM[2 * v]);
                            accM[2 * v + 1] = p_fma(p_mul(p_mk(a.z, a.w), nM[2 * v + 1]), fac, accM[2 * v + 1]);
                        }
#pragma unroll
                        for (int c = 0; c < C; c++) { sM[c] = nM[c]; sI[c] = nI[c]; sD[c] = nD[c]; }
                    }
                    rMb = cMb;
                    ebs = p_mul(ebs, ploop);
                    if (lane == 0 && !firstS) {
                        float *rec = bnd + WP_BND_REC * i;
                        *reinterpret_cast<float4 *>(rec) = make_float4(sM[0].x, sM[0].y, 0.f, 0.f);
                        *reinterpret_cast<float4 *>(rec + 4) = make_float4(sD[0].x, sD[0].y, bp.x, bp.y);
                        ((int *)rec)[8] = g.x; ((int *)rec)[9] = g.y;
                    }
                }
                const int tF = Ls + 31 - tp;
                if (((tF - 1) & 7) == 0 && tF > 1) {  // next step enters the previous exponent block
                    __syncwarp();
                    gFc = gFnext;
                    gblk--;
                    gFnext.x = gFs[2 * max(gblk - 1, 0)]; gFnext.y = gFs[2 * max(gblk - 1, 0) + 1];
                    fac = post_scale(gFc, g);
                }
            };
            auto brescale = [&](const int tp) {
                float2 mx = p_dup(0.f);
#pragma unroll
                for (int c = 0; c < C; c++) mx = p_max(mx, p_max(sM[c], p_max(sI[c], sD[c])));
                mx = p_max(mx, bp);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    mx = p_max(mx, p_mk(__shfl_xor_sync(FULL, mx.x, o), __shfl_xor_sync(FULL, mx.y, o)));
                PairI e_need = {(mx.x > 1048576.f) ? fexp(mx.x) : 0, (mx.y > 1048576.f) ? fexp(mx.y) : 0};
                if (!lastS) {
                    const int ic = max(Ls - tp, 0), r = max(0, Ls - (tp + W_SCALE_EVERY));
                    const int sl = ((r >> 3) == (ic >> 3)) ? bcur : bnd_next(bcur);
                    const unsigned ra = bnd_ring + sl * WP_BND_BLK + (r & 7) * (WP_BND_REC * 4);
                    e_need.x = max(e_need.x, lds_i1v(ra + 32) - 40 - g.x);
                    e_need.y = max(e_need.y, lds_i1v(ra + 36) - 40 - g.y);
                }
                if (e_need.x > 0 || e_need.y > 0) {
                    e_need.x = max(e_need.x, 0); e_need.y = max(e_need.y, 0);
                    const float2 f = p_mk(pow2i(-e_need.x), pow2i(-e_need.y));
                    g.x += e_need.x; g.y += e_need.y;
#pragma unroll
                    for (int c = 0; c < C; c++) { sM[c] = p_mul(sM[c], f); sI[c] = p_mul(sI[c], f); sD[c] = p_mul(sD[c], f); }
                    rMb = p_mul(rMb, f); bp = p_mul(bp, f); ebs = p_mul(ebs, f);
                    fac = post_scale(gFc, g);
                }
            };
            {
                const std::integral_constant<bool, false> genc;
                const std::integral_constant<bool, true> allc;
                int tp = 0;
                const int ramp = (32 + dL + 7) & ~7;   // every lane is past the last row of BOTH items after 32 + dL steps
                for (; tp < ramp && tp < nstepsB; tp++) {
                    bstep(tp, genc);
                    if ((tp & (W_SCALE_EVERY - 1)) == (W_SCALE_EVERY - 1)) brescale(tp);
                }
                for (; tp + 7 <= Ls - 1; tp += 8) {   // steady state: every lane has 1 <= i < min Ls; blocks of 8 steps
#pragma unroll 2
                    for (int u = 0; u < 8; u++) bstep(tp + u, allc);
                    brescale(tp + 7);
                }
                for (; tp < nstepsB; tp++) {
                    bstep(tp, genc);
                    if ((tp & (W_SCALE_EVERY - 1)) == (W_SCALE_EVERY - 1)) brescale(tp);
                }
            }
            {   // null2 numerators of both items: sum_k fM(k) * e_k(x) for the canonical symbols
                const int K = (E.Kp == 29) ? 20 : 4;
                for (int x = 0; x < K; x++) {
                    const float4 e4 = lds_f4(ebase + x * erow);
                    const float e[4] = {e4.x, e4.y, e4.z, e4.w};
                    float2 v = p_dup(0.f);
#pragma unroll
                    for (int c = 0; c < C; c++) v = p_fma(accM[c], p_dup(e[c]), v);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { v.x += __shfl_xor_sync(FULL, v.x, o); v.y += __shfl_xor_sync(FULL, v.y, o); }
                    if (lane == 0) { s_n2[w][0][x] += v.x; s_n2[w][1][x] += v.y; }
                }
                float2 v = p_dup(0.f);
#pragma unroll
                for (int c = 0; c < C; c++) v = p_add(v, accM[c]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { v.x += __shfl_xor_sync(FULL, v.x, o); v.y += __shfl_xor_sync(FULL, v.y, o); }
                accI = p_add(accI, v);
            }
            __syncwarp();
        }

        // ---- null2 by expectation, per item ----
        __syncwarp();
        const int K = (E.Kp == 29) ? 20 : 4;
#pragma unroll
        for (int item = 0; item < 2; item++) {
            const int Li = item == 0 ? Ls2.x : Ls2.y;
            const float aI = item == 0 ? accI.x : accI.y;
            const float norm = 1.0f / (float)Li;
            float ln2 = 0.f;
            if (lane < K) ln2 = logf((s_n2[w][item][lane] - aI) * norm + 1.0f);
            __syncwarp();
            if (lane < K) s_n2[w][item][lane] = (s_n2[w][item][lane] - aI) * norm + 1.0f;
            __syncwarp();
            if (lane >= K && lane < Q.nsym) {
                const int code = Q.symrow[lane];
                unsigned mask = 0;
                if (E.Kp == 29) {
                    const unsigned m[6] = {(1u << 11) | (1u << 2), (1u << 7) | (1u << 9), (1u << 13) | (1u << 3), 1u << 8, 1u << 1, 0xFFFFFu};
                    mask = (code >= 21 && code <= 26) ? m[code - 21] : 0u;
                } else {
                    const unsigned m[11] = {5, 10, 3, 12, 6, 9, 11, 14, 7, 13, 15};
                    mask = (code >= 5 && code <= 15) ? m[code - 5] : 0u;
                }
                float sum = 0.f; int cnt = 0;
                for (int x = 0; x < K; x++) if (mask >> x & 1u) { sum += s_n2[w][item][x]; cnt++; }
                ln2 = cnt ? logf(sum / (float)cnt) : 0.f;
            }
            const unsigned sr = item == 0 ? sresA : sresB;
            float dc = 0.f;
            for (int base = 0; base < Li; base += 32) {
                const int p = base + lane;
                const int xr = (p < Li) ? lds_u8(sr + p) : 0;
                for (int z = 0; z < 32 && base + z < Li; z++) {
                    const int xx = __shfl_sync(FULL, xr, z);
                    dc += __shfl_sync(FULL, ln2, xx);
                }
            }
            if (lane == 0 && (item == 0 || hasB)) {
                const WaveItem &it = item == 0 ? itA : itB;
                Wk.envsc[it.pair] = item == 0 ? fwd_nats.x : fwd_nats.y;
                Wk.domcorr[it.pair] = dc;
            }
            __syncwarp();
        }
    }
}

}  // namespace witch
