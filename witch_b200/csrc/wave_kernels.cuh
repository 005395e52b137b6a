// Family W: unihit-local Forward / Backward / posterior decoding / optimal accuracy as a warp-per-pair,
// strip-mined WAVEFRONT (SURVEY.md 8(a) "Score semantics" item 7 and "Align semantics").
//
// In unihit mode B(i) = N(i)*move is known in closed form, so no row needs a model-wide sum before the next row
// can start. One warp owns one (query, HMM, envelope) item and sweeps the model in strips of 32*C columns:
// lane l keeps the transition parameters of its C columns and the previous row's M/I/D in registers and, at step
// t, works on row i = t - l. Every dependency (diagonal, vertical, and the in-row D->D chain) then comes from
// the lane's own registers or from the left neighbour's values of the previous step: 4 shuffles per step, no
// barrier, no scan, and the D chain is exact. Strip boundaries (last column of every row) go through a small
// per-item array in L2. Forward rows are kept in a "wave layout" [strip][t][lane][C] so that the Backward sweep,
// which visits tile row t = L+31-t', reads them back fully coalesced for posterior decoding.
// Scaling: one power-of-two exponent per warp and step (exact integer bookkeeping), renormalised every 8 steps.
#pragma once
#include <type_traits>
#include "device_types.cuh"
#include "parser_kernel.cuh"

namespace witch {

struct WaveItem {
    int q, h;      // query, HMM
    int i0, Ls;    // envelope start (1-based) and length; align: i0 = 1, Ls = L
    int pair;      // output slot (envelope list index or align pair index)
};

struct WaveWork {
    const WaveItem *items;   // sorted: items of one HMM contiguous
    const int *group_first;  // first item of every group (a group = up to WAVE_WARPS consecutive items of one HMM; one CTA each)
    const int *grange;       // DEVICE pair {first group, end group} of this launch (the lists are built on the device)
    int item_end;            // items of this launch end here (a group never crosses it)
    unsigned *counter;
    // per-warp-slot scratch
    char *scratch;
    long long slot_bytes;
    int Lcap;       // max Ls in this launch
    int max_strips;
    int emis_floats;  // floats reserved for the emission table in dynamic shared memory (residue staging follows)
    int res_cap;      // bytes of residue staging per warp (multiple of 16, >= Lcap + 1)
    // outputs (envelope mode)
    float *envsc;   // [nitems] ln P(envelope | unihit model)
    float *domcorr; // [nitems] sum of ln null2 over the envelope
    // outputs (align mode)
    int *cols;               // concatenated column lists
    const long long *col_off;  // [npairs]
    float *dbg_fwd, *dbg_bwd;  // optional per item totals (nats)
};

constexpr int WAVE_WARPS_MAX = 8;  // warps per CTA (they share one HMM's emission table in shared memory)

constexpr int W_SCALE_EVERY = 8;
// Steps of the steady-state 8-step blocks unrolled together. Forward: all 8 (the boundary-block test on t & 7 and the
// ring-slot arithmetic become compile-time, and the loop-carried row state needs no copies): +6.8 % on the envelope pass.
// Backward: 2 (ping-pong register sets, +2.5 %); 4 or 8 lose to instruction-cache misses (its step is 280 instructions).
#ifndef WITCH_WAVE_UNROLL_F
#define WITCH_WAVE_UNROLL_F 8
#endif
#ifndef WITCH_WAVE_UNROLL_B
#define WITCH_WAVE_UNROLL_B 2
#endif
constexpr int W_UNROLL_F = WITCH_WAVE_UNROLL_F, W_UNROLL_B = WITCH_WAVE_UNROLL_B;
// Issue-slot trims of the Backward step (bit mask; 7 = all three is the default since round 2: +5 % on the envelope pass,
// bit-identical scores; 0 = the round-1 kernel):
//   1: the lane that issues a TMA copy is chosen with elect.sync instead of `lane == 0` (all operands are warp-uniform);
//      ptxas then drops the ELECT/R2UR "waterfall" loop it builds around UBLKCP for a possibly divergent issuer
//   2: the once-per-8-steps exponent-block switch of the Backward step is a real (warp-uniform) branch instead of 16
//      predicated-off instructions in every step
//   4: steady-state Backward steps skip the "rows left to request?" test of the Forward-row ring (always true there)
#ifndef WITCH_WAVE_EXP
#define WITCH_WAVE_EXP 7
#endif
constexpr int W_EXP = WITCH_WAVE_EXP;
// Stored Forward match rows of the ENVELOPE pass (DNA/RNA, warp-uniform exponent) in 16 bits: halves the pass's HBM traffic
// (8 -> 4 B/cell). 0 = FP32 rows; 1 = bf16 (round to nearest, 2^-9 relative); 2 = unsigned E8M8 (the values are >= 0, so
// the sign bit's place is given to the mantissa: bits [30:15] of the FP32, round to nearest, 2^-10 relative).
#ifndef WITCH_WAVE_ROW16
#define WITCH_WAVE_ROW16 0
#endif
constexpr int W_ROW16 = WITCH_WAVE_ROW16;
__device__ __forceinline__ unsigned pack_row16(float lo, float hi) {
    if (W_ROW16 == 1) {
#ifndef WITCH_HOST_SIM
        unsigned r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
#else
        const unsigned a = __float_as_uint(lo), b = __float_as_uint(hi);   // round to nearest even
        return ((a + 0x7fffu + ((a >> 16) & 1u)) >> 16) | ((b + 0x7fffu + ((b >> 16) & 1u)) & 0xffff0000u);
#endif
    }
    return (((__float_as_uint(lo) + 0x4000u) >> 15) & 0xffffu) | (((__float_as_uint(hi) + 0x4000u) << 1) & 0xffff0000u);
}
__device__ __forceinline__ float unpack_row16_lo(unsigned w) { return __uint_as_float(W_ROW16 == 1 ? (w << 16) : ((w << 15) & 0x7fff8000u)); }
__device__ __forceinline__ float unpack_row16_hi(unsigned w) { return __uint_as_float(W_ROW16 == 1 ? (w & 0xffff0000u) : ((w >> 1) & 0x7fff8000u)); }
#ifndef WITCH_HOST_SIM
__device__ __forceinline__ bool wave_elect_one() {
    unsigned p;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ uint4 lds_u4v(unsigned a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
#endif
__host__ __device__ constexpr int wave_ring_stage_bytes(int C, bool align, bool row16 = false) { return 32 * C * (row16 ? 2 : 4) * (align ? 2 : 1); }
#ifndef WITCH_HOST_SIM
// ---- TMA (1-D bulk async copy) + mbarrier: completion is tracked in shared memory, not on a register scoreboard
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ int lds_i1v(unsigned a) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
#endif
constexpr int W_BND_SLOTS = 3;      // ring of 8-row blocks of strip-boundary records (256 B each) per warp
__host__ __device__ constexpr int wave_bnd_ring_bytes() { return W_BND_SLOTS * (256 + 8); }
#ifndef WITCH_HOST_SIM
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float4 lds_f4v(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
#endif

// scratch layout helper (all offsets in bytes, per warp slot)
struct WaveLayout {
    long long tileM, tileI, gF, bnd, rows, bits, total;
    int TT;
};
__host__ __device__ inline WaveLayout wave_layout(int Lcap, int max_strips, int C, bool align, bool lane_exp) {
    WaveLayout w;
    w.TT = Lcap + 32;
    const bool row16 = W_ROW16 != 0 && !align && !lane_exp;
    long long tile = (long long)max_strips * w.TT * 32 * C * (row16 ? 2 : 4);
    long long o = 0;
    w.tileM = o; o += tile;
    w.tileI = o; if (align) o += tile;  // insert rows are kept for the align stage only
    w.gF = o; o += ((long long)max_strips * (w.TT / 8 + 2) * (lane_exp ? 32 : 1) * 4 + 15) / 16 * 16;  // exponents per block of 8 steps
    w.bnd = o; o += (long long)8 * (Lcap + 16) * 4;   // boundary records {M,I,D,E,G,-,-,-} per row (read in blocks of 8 rows)
    w.rows = o; o += (long long)10 * (Lcap + 2) * 4;  // FC,FCg,NB,NBg,NOA,PPC,EOA,KE,...
    w.bits = o; if (align) o += (long long)max_strips * w.TT * 32 * 4;
    w.total = (o + 255) / 256 * 256;
    return w;
}

// emission odds of C consecutive columns for symbol row at shared address `a` (strip-interleaved layout)
template <int C>
__device__ __forceinline__ void lds_emis(unsigned a, float (&e)[C]) {
    static_assert(C % 4 == 0, "wave kernels need C % 4 == 0");
#pragma unroll
    for (int v = 0; v < C / 4; v++) {
        const float4 t = lds_f4(a + v * 512);
        e[4 * v] = t.x; e[4 * v + 1] = t.y; e[4 * v + 2] = t.z; e[4 * v + 3] = t.w;
    }
}

// comparator for select_e (HMMER visits cells in striped order; M uses >=, D uses >): returns true if
// candidate (v2,isD2,ord2) replaces current (v1,isD1,ord1)
__device__ __forceinline__ bool oa_e_better(float v2, int d2, int o2, float v1, int d1, int o1) {
    if (v2 > v1) return true;
    if (v2 < v1) return false;
    if (d1 != d2) return d2 == 0;           // M beats D on ties
    return d2 == 0 ? (o2 > o1) : (o2 < o1);  // last M in visiting order, first D in visiting order
}

template <int C, bool ALIGN, int WAVE_WARPS, int MINB, int W_RING, bool LANE_EXP>
__global__ void __launch_bounds__(WAVE_WARPS * 32, MINB) wave_kernel(DevEhmm E, DevQueries Q, WaveWork Wk) {
    WITCH_DYN_SMEM(float, smem);
    static_assert(W_RING >= 2, "the Backward sweep reads stored Forward rows through the TMA ring");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __shared__ int s_group;
    __shared__ float s_n2[WAVE_WARPS][MAX_SYM];
    float *emis_s = smem;  // [nsym][Mstr] in strip-interleaved layout
    const WaveLayout lay = wave_layout(Wk.Lcap, Wk.max_strips, C, ALIGN, LANE_EXP);
    __shared__ char *s_slot[WAVE_WARPS];
    if (lane == 0) s_slot[w] = Wk.scratch + ((long long)blockIdx.x * WAVE_WARPS + w) * Wk.slot_bytes;
    __syncwarp();
    char *slot = *((char *volatile *)&s_slot[w]);  // a loaded value: keeps the compiler from rebuilding it from blockIdx
    __builtin_assume(__isGlobal(slot));
    float *tileM = (float *)(slot + lay.tileM), *tileI = (float *)(slot + lay.tileI);
    int *gFarr = (int *)(slot + lay.gF);
    const int LB = Wk.Lcap + 2;
    float *bnd = (float *)(slot + lay.bnd);  // records of 8 words per row
    int *bndi = (int *)bnd;
#define BND_M(i) bnd[8 * (i)]
#define BND_I(i) bnd[8 * (i) + 1]
#define BND_D(i) bnd[8 * (i) + 2]
#define BND_E(i) bnd[8 * (i) + 3]
#define BND_G(i) bndi[8 * (i) + 4]
#define BND_V(i) (*reinterpret_cast<float4 *>(bnd + 8 * (i)))  // {M, I, D, E} of one record as one 128-bit access
    float *rFC = (float *)(slot + lay.rows);
    int *rFCg = (int *)(rFC + LB);
    float *rNB = (float *)(rFCg + LB);
    int *rNBg = (int *)(rNB + LB);
    float *rNOA = (float *)(rNBg + LB), *rPPC = rNOA + LB, *rEOA = rPPC + LB;
    int *rKE = (int *)(rEOA + LB);
    unsigned *bits = (unsigned *)(slot + lay.bits);
    const int TT = lay.TT, TG = (lay.TT / 8 + 2) * (LANE_EXP ? 32 : 1);
    // exponent of a block of 8 Forward steps: one word per warp (uniform exponent) or per lane (LANE_EXP)
#define GF_AT(blk) (LANE_EXP ? (blk) * 32 + lane : (blk))
    const unsigned emis_sa = smem_u32(emis_s);
    const int SW = 32 * C;  // strip width
    constexpr bool ROW16 = W_ROW16 != 0 && !ALIGN && !LANE_EXP;   // stored Forward match rows in 16 bits (envelope pass, DNA/RNA)
    static_assert(!ROW16 || C == 8, "16-bit rows: one 128-bit word per lane and step");
    constexpr int TSTEP = ROW16 ? 32 * C / 2 : 32 * C;   // floats (4-byte words) of one stored step of a strip

    // ring state persists across strips and items (the mbarriers are initialised once; phases keep alternating)
    int rd_stage = 0, wr_stage = 0;
    unsigned rd_phase = 0;
    int bq_w = 0, bq_r = 0;   // boundary-record ring: next slot to fill / to wait for
    unsigned bq_ph = 0;
    // dynamic shared memory after the emission table: residues | Forward-row ring | its mbarriers | boundary ring | its mbarriers
    constexpr int RSB = wave_ring_stage_bytes(C, ALIGN, ROW16);
    const unsigned sm_dyn = emis_sa + Wk.emis_floats * 4;
    const unsigned ring_w = sm_dyn + WAVE_WARPS * Wk.res_cap + w * (W_RING * RSB);
    const unsigned ring_sa = ring_w + lane * 16;
    const unsigned ring_bar = sm_dyn + WAVE_WARPS * Wk.res_cap + WAVE_WARPS * (W_RING * RSB) + w * (W_RING * 8);
    const unsigned bnd_ring = sm_dyn + WAVE_WARPS * Wk.res_cap + WAVE_WARPS * (W_RING * (RSB + 8)) + w * (W_BND_SLOTS * 256);
    const unsigned bnd_bar = sm_dyn + WAVE_WARPS * Wk.res_cap + WAVE_WARPS * (W_RING * (RSB + 8) + W_BND_SLOTS * 256) + w * (W_BND_SLOTS * 8);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < W_RING; k++) mbar_init(ring_bar + k * 8, 1);
#pragma unroll
        for (int k = 0; k < W_BND_SLOTS; k++) mbar_init(bnd_bar + k * 8, 1);
        fence_mbar_init();
    }
    __syncwarp();
    // boundary records of the neighbouring strip arrive in blocks of 8 rows through a small TMA-fed ring
    auto bnd_issue = [&](const int b) {
        if ((W_EXP & 1) ? wave_elect_one() : (lane == 0)) {
            mbar_expect_tx(bnd_bar + bq_w * 8, 256);
            tma_load_1d(bnd_ring + bq_w * 256, bnd + 64 * b, 256, bnd_bar + bq_w * 8);
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
        if (threadIdx.x == 0) s_group = Wk.grange[0] + (int)atomicAdd(Wk.counter, 1u);
        __syncthreads();
        const int grp = s_group;
        if (grp >= Wk.grange[1]) break;
        const int gfirst = Wk.group_first[grp];
        const int h = Wk.items[gfirst].h;
        int gcount = 1;   // consecutive items of the same HMM, at most one per warp
        while (gcount < WAVE_WARPS && gfirst + gcount < Wk.item_end && Wk.items[gfirst + gcount].h == h) gcount++;
        const int Mh = E.M[h];
        const int nstrips = (Mh + SW - 1) / SW;
        if (h != loaded_h) {
            Mstr = nstrips * SW;
            const int st = E.stride[h];
            const float *eg = E.emis + E.eoff[h];
            for (int idx = threadIdx.x; idx < Q.nsym * Mstr; idx += blockDim.x) {
                int x = idx / Mstr, col = idx - x * Mstr;
                int s = col / SW, r = col - s * SW, l = r / C, cc = r - l * C;
                float v = (col < st - 1) ? __ldg(eg + (size_t)Q.symrow[x] * st + 1 + col) : 0.f;
                emis_s[(size_t)x * Mstr + s * SW + emis_index<C>(32, l, cc)] = v;
            }
            loaded_h = h;
        }
        __syncthreads();
        if (w >= gcount) continue;
        const WaveItem it = Wk.items[gfirst + w];
        const int Ls = it.Ls, Lfull = Q.len[it.q];
        const uint8_t *dsq = Q.dsq + Q.off[it.q] + (it.i0 - 1);  // dsq[i-1] = residue i of the envelope
        // the item's residues are staged in shared memory (one byte each; the launch sizes res_cap >= Lcap + 1)
        const unsigned sres = emis_sa + Wk.emis_floats * 4 + w * Wk.res_cap;
        {
            uint8_t *sr = reinterpret_cast<uint8_t *>(emis_s + Wk.emis_floats) + w * Wk.res_cap;
            for (int z = lane; z < it.Ls; z += 32) sr[z] = dsq[z];
            __syncwarp();
        }
#define RES_AT(idx) lds_u8(sresp + (idx))
        // (this warp's ring of stored Forward rows: [stage][M|I][v][lane][4 floats], conflict-free LDS.128)
        const long long po = E.poff[h];
        const float pmove = 2.0f / ((float)Lfull + 2.0f), ploop = 1.0f - pmove;
        const unsigned FULL = 0xffffffffu;
        const int nsteps = Ls + 31;

        // ======================================= Forward =======================================
        float xCv = 0.f;  // C special (lane 31 of the last strip), at exponent xCg
        int xCg = 0;
        float Tm = 1.f; int gT = 0;
        for (int s = 0; s < nstrips; s++) {
            const long long k0 = (long long)s * SW + lane * C;  // owns columns k0+1..k0+C
            float pa[C], pb[C], pg[C], pmd[C], pdd[C], pmi[C], pii[C], pen[C];
            load_cols<C>(E.tMM + po, k0, pa); load_cols<C>(E.tIM + po, k0, pb); load_cols<C>(E.tDM + po, k0, pg);
            load_cols<C>(E.tMD + po, k0, pmd); load_cols<C>(E.tDD + po, k0, pdd);
            load_cols<C>(E.tMI + po, k0 + 1, pmi); load_cols<C>(E.tII + po, k0 + 1, pii);
            load_cols<C>(E.entry + po, k0 + 1, pen);
            float sM[C], sI[C], sD[C];
#pragma unroll
            for (int c = 0; c < C; c++) { sM[c] = 0.f; sI[c] = 0.f; sD[c] = 0.f; }
            float rM = 0.f, rI = 0.f, rD = 0.f;      // row i-1 at the column left of the owned block
            float ep = 0.f;                          // running E(i) partial of the lane's last row
            // Scaling exponent. DNA: one per warp (32 rows of <= 2 bits each fit FP32's range). LANE_EXP (amino: a row of a
            // conserved rare residue is worth > 6 bits): one per lane, changed only at the common 8-step checks and
            // coupled so that a neighbour's exponent is at most 16 above (values entering a lane convert by <= 2^16).
            int g = (s > 0) ? BND_G(1) : 0;
            float xBs = pmove * pow2i(-g);           // pmove * N(i-1) * 2^-g for the lane's next row
            const bool last = (s == nstrips - 1);
            unsigned ebase = emis_sa + (s * SW + lane * 4) * 4;
            unsigned erow = Mstr * 4;
            float *tM = tileM + (size_t)s * TT * TSTEP, *tI = tileI + (size_t)s * TT * SW;
            // wave-layout tile of a strip: [step t][v][lane][4 floats] -- every 128-bit access of a warp is one
            // contiguous 512-byte run (v-th quad of the lane's C columns); 16-bit rows: [step t][lane][8 x 16 bit]
            int toff = TSTEP + lane * 4;  // element offset of (step t, lane) quad 0
            int *gFs = gFarr + s * TG;
            unsigned sresp = sres;
            PIN32(ebase); PIN32(erow); PIN64(tM); PIN64(gFs); PIN32(sresp);
            if (ALIGN) PIN64(tI);
            if (last) { xCv = 0.f; xCg = g; }
            int xcur = RES_AT(min(max(-lane, 0), Ls - 1));  // residue of the lane's row at the next step
            int bcur = 0;  // ring slot of the boundary block that holds lane 0's current row
            if (s > 0) {
                fence_proxy_async();  // the left strip's records (generic-proxy stores of this warp) before the TMA reads
                __syncwarp();
                bcur = bq_r;
                const int bmax = Ls >> 3;
                bnd_issue(0);
                if (bmax >= 1) bnd_issue(1);
                if (bmax >= 2) bnd_issue(2);
                bnd_wait();
                if (bmax >= 1) bnd_wait();
            }
            float *tMp = tM + toff, *tIp = tI + toff;  // running tile pointers of the current step
            // One wavefront step. ALL = every lane is inside the sequence (steady state: no activity predicate, no clamps).
            auto fstep = [&](const int t, auto allc) {
                constexpr bool ALL = decltype(allc)::value;
                const int i = t - lane;
                const bool act = ALL || (i >= 1 && i <= Ls);
                // values of row i at the column left of my block: left lane's last step, or the strip boundary
                float cM = __shfl_up_sync(FULL, sM[C - 1], 1);
                float cI = __shfl_up_sync(FULL, sI[C - 1], 1);
                float cD = __shfl_up_sync(FULL, sD[C - 1], 1);
                float cE = __shfl_up_sync(FULL, ep, 1);
                if (LANE_EXP) {   // the left lane's values are in its own scale
                    const float fl = pow2i(__shfl_up_sync(FULL, g, 1) - g);
                    cM *= fl; cI *= fl; cD *= fl; cE *= fl;
                }
                {   // lane 0: strip boundary of the left strip, or zeros (first strip / outside the sequence); branch-free
                    float f = 0.f;
                    float4 pbv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (s > 0) {
                        const int tr = ALL ? t : min(t, Ls);
                        if ((t & 7) == 0 && (ALL || t <= Ls)) {   // row t opens boundary block t >> 3
                            const int b = t >> 3;
                            __syncwarp();  // the slot refilled now was last read a block ago
                            if (8 * (b + 2) <= Ls) bnd_issue(b + 2);
                            if (8 * (b + 1) <= Ls) bnd_wait();   // keep the next block readable too (exponent look-ahead)
                            bcur = bnd_next(bcur);
                        }
                        const unsigned ra = bnd_ring + bcur * 256 + (tr & 7) * 32;
                        pbv = lds_f4v(ra);
                        f = act ? pow2i(lds_i1v(ra + 16) - g) : 0.f;
                    }
                    const float bM = pbv.x * f, bI = pbv.y * f, bD = pbv.z * f, bE = pbv.w * f;
                    cM = lane == 0 ? bM : cM; cI = lane == 0 ? bI : cI; cD = lane == 0 ? bD : cD; cE = lane == 0 ? bE : cE;
                }
                const int xres = xcur;
                xcur = RES_AT(ALL ? i : min(max(i, 0), Ls - 1));
                if (act) {
                    float e[C];
                    lds_emis<C>(ebase + xres * erow, e);
                    float nM[C], nI[C], nD[C];
#pragma unroll
                    for (int c = C - 1; c >= 0; c--) {
                        float pm = c > 0 ? sM[c - 1] : rM, pi = c > 0 ? sI[c - 1] : rI, pd = c > 0 ? sD[c - 1] : rD;
                        nI[c] = sM[c] * pmi[c] + sI[c] * pii[c];
                        float acc = xBs * pen[c];
                        acc = fmaf(pm, pa[c], acc); acc = fmaf(pi, pb[c], acc); acc = fmaf(pd, pg[c], acc);
                        nM[c] = acc * e[c];
                    }
                    // in-row D chain: the match terms are products off the chain, each link is ONE dependent FFMA
                    nD[0] = fmaf(cD, pdd[0], cM * pmd[0]);
#pragma unroll
                    for (int c = 1; c < C; c++) nD[c] = fmaf(nD[c - 1], pdd[c], nM[c - 1] * pmd[c]);
                    float es = cE;
#pragma unroll
                    for (int c = 0; c < C; c++) { sM[c] = nM[c]; sI[c] = nI[c]; sD[c] = nD[c]; es += nM[c] + nD[c]; }
                    ep = es;
                    rM = cM; rI = cI; rD = cD;
                    xBs *= ploop;
                    // keep the row for posterior decoding (wave layout)
                    float *dm = tMp, *di = tIp;
                    // (envelope mode needs match posteriors only: sum over emitting states of a row's posteriors is 1,
                    //  so fI + fNCJ = 1 - sum_k fM(k); insert rows are stored for the align stage only)
                    if (ROW16) {
                        *reinterpret_cast<uint4 *>(dm) = make_uint4(pack_row16(nM[0], nM[1]), pack_row16(nM[2], nM[3]),
                                                                    pack_row16(nM[4], nM[5]), pack_row16(nM[C - 2], nM[C - 1]));
                    } else {
#pragma unroll
                        for (int v = 0; v < C / 4; v++) {
                            *reinterpret_cast<float4 *>(dm + 128 * v) = make_float4(nM[4 * v], nM[4 * v + 1], nM[4 * v + 2], nM[4 * v + 3]);
                            if (ALIGN) *reinterpret_cast<float4 *>(di + 128 * v) = make_float4(nI[4 * v], nI[4 * v + 1], nI[4 * v + 2], nI[4 * v + 3]);
                        }
                    }
                    if (lane == 31) {
                        if (!last) { BND_V(i) = make_float4(sM[C - 1], sI[C - 1], sD[C - 1], ep); BND_G(i) = g; }
                        else {
                            // C(i) = C(i-1)*loop + E(i)   (unihit: E->C = 1), with exponent alignment
                            const float f = pow2i(xCg - g);
                            xCv = xCv * f * ploop + ep; xCg = g;
                            rFC[i] = xCv; rFCg[i] = g;
                        }
                    }
                }
                tMp += TSTEP;
                if (ALIGN) tIp += 32 * C;
            };
            // every 8 steps: record the exponent of the block (for the Backward pass) / renormalise
            auto fmark = [&](const int t) { if (LANE_EXP || lane == 0) gFs[GF_AT((t - 1) >> 3)] = g; };  // exponent of steps t .. t+7
            auto frescale = [&](const int t) {
                {
                    float mx = 0.f;
#pragma unroll
                    for (int c = 0; c < C; c++) mx = fmaxf(mx, fmaxf(sM[c], fmaxf(sI[c], sD[c])));
                    mx = fmaxf(mx, ep);
                    if (!LANE_EXP) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
                    }
                    int e_need = (mx > 1048576.f) ? fexp(mx) : 0;
                    if (s > 0) {   // exponent the left strip had 8 rows ahead (that block is already in the ring)
                        const int r = min(Ls, t + W_SCALE_EVERY), tc = min(t, Ls);
                        const int sl = ((r >> 3) == (tc >> 3)) ? bcur : bnd_next(bcur);
                        const int ahead = lds_i1v(bnd_ring + sl * 256 + (r & 7) * 32 + 16) - 40 - g;
                        if (!LANE_EXP || lane == 0) e_need = max(e_need, ahead);
                    }
                    if (LANE_EXP) {   // couple the lanes: g_l >= g_{l-1} - 16 (max-plus scan over the new exponents)
                        int need = g + e_need;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int up = __shfl_up_sync(FULL, need, o);
                            if (lane >= o) need = max(need, up - 16 * o);
                        }
                        e_need = need - g;
                    }
                    if (e_need > 0) {
                        const float f = pow2i(-e_need);
                        g += e_need;
#pragma unroll
                        for (int c = 0; c < C; c++) { sM[c] *= f; sI[c] *= f; sD[c] *= f; }
                        rM *= f; rI *= f; rD *= f; ep *= f; xBs *= f;
                    }
                }
            };
            {
                const std::integral_constant<bool, false> genc;
                const std::integral_constant<bool, true> allc;
                int t = 1;
                for (; t <= 32; t++) {   // ramp-up (nsteps = Ls + 31 >= 32)
                    if (((t - 1) & 7) == 0) fmark(t);
                    fstep(t, genc);
                    if ((t & (W_SCALE_EVERY - 1)) == 0) frescale(t);
                }
                for (; t + 7 <= Ls; t += 8) {   // steady state: all 32 lanes inside the sequence, blocks of 8 steps
                    fmark(t);
#pragma unroll W_UNROLL_F
                    for (int u = 0; u < 8; u++) fstep(t + u, allc);
                    frescale(t + 7);
                }
                for (; t <= nsteps; t++) {   // drain
                    if (((t - 1) & 7) == 0) fmark(t);
                    fstep(t, genc);
                    if ((t & (W_SCALE_EVERY - 1)) == 0) frescale(t);
                }
            }
            if (last) {
                xCv = __shfl_sync(FULL, xCv, 31); xCg = __shfl_sync(FULL, xCg, 31);
            }
            __syncwarp();
        }
        Tm = xCv * pmove; gT = xCg;  // P = Tm * 2^gT
        const float fwd_nats = logf(Tm) + (float)gT * 0.69314718056f;
        if (lane == 0 && Wk.dbg_fwd) Wk.dbg_fwd[it.pair] = fwd_nats;
        const float invT = 1.0f / Tm;

        // ======================================= Backward =======================================
        // lane l processes row i = Ls - (t' - (31 - l)), rows Ls .. 0 (row 0 only feeds the B special)
        float xNv = 0.f; int xNg = 0;  // N special (lane 0 of strip 0)
        float accI = 0.f;             // sum of all match posteriors (envelope mode)
        if (!ALIGN && lane < MAX_SYM) s_n2[w][lane] = 0.f;
        __syncwarp();
        const int nstepsB = Ls + 1 + 31;
        for (int s = nstrips - 1; s >= 0; s--) {
            const long long k0 = (long long)s * SW + lane * C;
            float oMM[C], oIM[C], oDM[C], oMD[C], oDD[C], oMI[C], oII[C], pen[C];
            load_cols<C>(E.tMM + po, k0 + 1, oMM); load_cols<C>(E.tIM + po, k0 + 1, oIM);
            load_cols<C>(E.tDM + po, k0 + 1, oDM); load_cols<C>(E.tMD + po, k0 + 1, oMD);
            load_cols<C>(E.tDD + po, k0 + 1, oDD); load_cols<C>(E.tMI + po, k0 + 1, oMI);
            load_cols<C>(E.tII + po, k0 + 1, oII); load_cols<C>(E.entry + po, k0 + 1, pen);
            float sM[C], sI[C], sD[C], accM[C];
#pragma unroll
            for (int c = 0; c < C; c++) { sM[c] = 0.f; sI[c] = 0.f; sD[c] = 0.f; accM[c] = 0.f; }
            float rMb = 0.f;   // Mb(i+1, first column right of my block)
            float bp = 0.f;    // running B(i) partial
            const bool lastS = (s == nstrips - 1), firstS = (s == 0);
            int g = lastS ? 0 : BND_G(Ls);
            float ebs = pmove * pow2i(-g);  // E(i) = C_b(i) = pmove * loop^(Ls-i), scaled
            unsigned ebase = emis_sa + (s * SW + lane * 4) * 4;
            unsigned erow = Mstr * 4;
            // emission of the first column of the right neighbour's block (k0 + C + 1)
            const int rs = (lane == 31) ? s + 1 : s, rl = (lane == 31) ? 0 : lane + 1;
            const bool hasR = !(lastS && lane == 31);
            unsigned eright = emis_sa + (rs * SW + emis_index<C>(32, rl, 0)) * 4;
            float *tM = tileM + (size_t)s * TT * TSTEP, *tI = tileI + (size_t)s * TT * SW;
            const int *gFs = gFarr + s * TG;
            unsigned sresp = sres;
            PIN32(ebase); PIN32(erow); PIN32(eright); PIN64(tM); PIN64(gFs); PIN32(sresp);
            if (ALIGN) PIN64(tI);
            if (firstS) { xNv = 0.f; xNg = g; }
            int xcur = RES_AT(Ls - 1);  // residue i+1 of the lane's row at the next step
            int bcur = 0;  // ring slot of the boundary block that holds lane 31's current row
            int gblk = (Ls + 30) >> 3;
            int gFc = gFs[GF_AT(gblk)];
            int gFnext = gFs[GF_AT(max(gblk - 1, 0))];  // exponent of the next (earlier) block of forward steps, prefetched
            auto post_scale = [&](const int gf, const int gb) {
                const float x = (float)(gf + gb - gT);
                return exp2f(LANE_EXP ? fminf(x, 120.f) : x) * invT;
            };
            float fac = post_scale(gFc, g);  // posterior scale; refreshed when an exponent changes
            // Stored Forward rows come back through a shared-memory ring filled by TMA bulk copies, W_RING-1 steps ahead
            // of their use (one lane issues one copy per step: the step's rows of all 32 lanes are one contiguous run;
            // completion is an mbarrier transaction count, so no register scoreboard is tied up by the prefetch).
            const float *tMrd = tM + (size_t)(Ls + 31) * TSTEP;   // rows of step 0; step tq is TSTEP words earlier
            const float *tIrd = tI + (size_t)(Ls + 31) * 32 * C;
            float *tMw = tM + (size_t)(Ls + 31) * 32 * C + lane * 4, *tIw = tI + (size_t)(Ls + 31) * 32 * C + lane * 4;
            int tq_next = 0;   // next step whose rows have not been requested yet
            auto ring_issue = [&]() {
                if ((W_EXP & 1) ? wave_elect_one() : (lane == 0)) {
                    const unsigned dst = ring_w + wr_stage * RSB, bar = ring_bar + wr_stage * 8;
                    mbar_expect_tx(bar, RSB);
                    tma_load_1d(dst, tMrd, TSTEP * 4, bar);
                    if (ALIGN) tma_load_1d(dst + 32 * C * 4, tIrd, 32 * C * 4, bar);
                }
                tMrd -= TSTEP;
                if (ALIGN) tIrd -= 32 * C;
                wr_stage = (wr_stage == W_RING - 1) ? 0 : wr_stage + 1;
                tq_next++;
            };
            fence_proxy_async();  // this warp's generic-proxy writes (tile, boundary records) are ordered before the TMA reads
            __syncwarp();
#pragma unroll
            for (int k = 0; k < W_RING - 1; k++) ring_issue();   // nstepsB >= 32 > W_RING
            if (!lastS) {
                bcur = bq_r;
                const int b0 = Ls >> 3;
                bnd_issue(b0);
                if (b0 >= 1) bnd_issue(b0 - 1);
                if (b0 >= 2) bnd_issue(b0 - 2);
                bnd_wait();
                if (b0 >= 1) bnd_wait();
            }

            // One wavefront step. ALL = every lane has 1 <= i < Ls (steady state: no predicates, no clamps).
            auto bstep = [&](const int tp, auto allc) {
                constexpr bool ALL = decltype(allc)::value;
                const int i = Ls - (tp - (31 - lane));
                const bool act = ALL || (i >= 0 && i <= Ls);
                float cMb = __shfl_down_sync(FULL, sM[0], 1);   // Mb(i, right column)   (row i of the right lane)
                float cDb = __shfl_down_sync(FULL, sD[0], 1);   // Db(i, right column)
                float cB = __shfl_down_sync(FULL, bp, 1);
                if (LANE_EXP) {   // the right lane's values are in its own scale
                    const float fl = pow2i(__shfl_down_sync(FULL, g, 1) - g);
                    cMb *= fl; cDb *= fl; cB *= fl;
                }
                {   // lane 31: boundary of the strip to the right, or zeros (last strip / outside the sequence); branch-free
                    float f = 0.f;
                    float4 pbv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!lastS) {
                        const int i31 = ALL ? Ls - tp : max(Ls - tp, 0);   // lane 31's row
                        if ((i31 & 7) == 7 && tp > 0 && (ALL || tp <= Ls)) {   // row i31 opens the next lower boundary block
                            const int b = i31 >> 3;
                            __syncwarp();
                            if (b >= 2) bnd_issue(b - 2);
                            if (b >= 1) bnd_wait();
                            bcur = bnd_next(bcur);
                        }
                        const unsigned ra = bnd_ring + bcur * 256 + (i31 & 7) * 32;
                        pbv = lds_f4v(ra);
                        f = act ? pow2i(lds_i1v(ra + 16) - g) : 0.f;
                    }
                    const float bM = pbv.x * f, bD = pbv.z * f, bB = pbv.w * f;
                    cMb = lane == 31 ? bM : cMb; cDb = lane == 31 ? bD : cDb; cB = lane == 31 ? bB : cB;
                }
                const int tF = Ls + 31 - tp;  // forward tile row holding row i of this lane (valid for i >= 1)
                // keep the ring W_RING-1 steps ahead (tile row 0 exists and is never used), then wait for this step's rows
                __syncwarp();  // every lane has consumed the stage that is refilled now (it was read one step ago)
                if (((W_EXP & 4) && ALL) || tq_next < nstepsB) ring_issue();   // steady state: tq_next <= Ls + 1 < nstepsB always
                mbar_wait(ring_bar + rd_stage * 8, rd_phase);
                const unsigned rs = ring_sa + rd_stage * RSB;
                if (rd_stage == W_RING - 1) { rd_stage = 0; rd_phase ^= 1u; } else rd_stage++;
                const int xres = xcur;
                xcur = RES_AT(ALL ? i - 1 : min(max(i - 1, 0), Ls - 1));
                if (act) {
                    float mn[C], mnR;
                    if (ALL || i < Ls) {
                        const int xr = xres;
                        float e[C];
                        lds_emis<C>(ebase + xr * erow, e);
#pragma unroll
                        for (int c = 0; c < C; c++) mn[c] = sM[c] * e[c];
                        mnR = hasR ? rMb * lds_f1(eright + xr * erow) : 0.f;
                    } else {
#pragma unroll
                        for (int c = 0; c < C; c++) mn[c] = 0.f;
                        mnR = 0.f;
                    }
                    float bs = cB;
#pragma unroll
                    for (int c = 0; c < C; c++) bs = fmaf(mn[c], pen[c], bs);
                    bp = bs;
                    if (ALL || i >= 1) {
                        float nM[C], nI[C], nD[C];
#pragma unroll
                        for (int c = C - 1; c >= 0; c--) {
                            const float m1 = (c < C - 1) ? mn[c + 1] : mnR;
                            const float dr = (c < C - 1) ? nD[c + 1] : cDb;
                            nD[c] = fmaf(dr, oDD[c], fmaf(m1, oDM[c], ebs));   // one dependent FFMA per link of the D chain
                            nM[c] = fmaf(m1, oMM[c], fmaf(sI[c], oMI[c], fmaf(dr, oMD[c], ebs)));
                            nI[c] = fmaf(m1, oIM[c], sI[c] * oII[c]);
                        }
                        // posterior decoding against the stored forward row (this step's ring stage is complete)
                        float FMv[C], FIv[C];
                        if (ROW16) {
                            const uint4 a = lds_u4v(rs);
                            FMv[0] = unpack_row16_lo(a.x); FMv[1] = unpack_row16_hi(a.x); FMv[2] = unpack_row16_lo(a.y); FMv[3] = unpack_row16_hi(a.y);
                            FMv[4] = unpack_row16_lo(a.z); FMv[5] = unpack_row16_hi(a.z); FMv[C - 2] = unpack_row16_lo(a.w); FMv[C - 1] = unpack_row16_hi(a.w);
                        }
#pragma unroll
                        for (int v = 0; v < C / 4; v++) {
                            if (!ROW16) {
                                const float4 a = lds_f4v(rs + v * 512);
                                FMv[4 * v] = a.x; FMv[4 * v + 1] = a.y; FMv[4 * v + 2] = a.z; FMv[4 * v + 3] = a.w;
                            }
                            if (ALIGN) {
                                const float4 b = lds_f4v(rs + 32 * C * 4 + v * 512);
                                FIv[4 * v] = b.x; FIv[4 * v + 1] = b.y; FIv[4 * v + 2] = b.z; FIv[4 * v + 3] = b.w;
                            } else { FIv[4 * v] = 0.f; FIv[4 * v + 1] = 0.f; FIv[4 * v + 2] = 0.f; FIv[4 * v + 3] = 0.f; }
                        }
                        if (ALIGN) {
                            float pM[C], pI[C];
#pragma unroll
                            for (int c = 0; c < C; c++) { pM[c] = (FMv[c] * fac) * nM[c]; pI[c] = (FIv[c] * fac) * nI[c]; }
#pragma unroll
                            for (int v = 0; v < C / 4; v++) {
                                *reinterpret_cast<float4 *>(tMw + 128 * v) = make_float4(pM[4 * v], pM[4 * v + 1], pM[4 * v + 2], pM[4 * v + 3]);
                                *reinterpret_cast<float4 *>(tIw + 128 * v) = make_float4(pI[4 * v], pI[4 * v + 1], pI[4 * v + 2], pI[4 * v + 3]);
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < C; c++)   // (amino: scale first -- F*B of two unscaled values can leave FP32's range)
                                accM[c] = LANE_EXP ? fmaf(FMv[c] * fac, nM[c], accM[c]) : fmaf(FMv[c] * nM[c], fac, accM[c]);
                        }
#pragma unroll
                        for (int c = 0; c < C; c++) { sM[c] = nM[c]; sI[c] = nI[c]; sD[c] = nD[c]; }
                    }
                    rMb = cMb;
                    ebs *= ploop;
                    if (lane == 0) {
                        if (!firstS) { BND_V(i) = make_float4(sM[0], 0.f, sD[0], bp); BND_G(i) = g; }
                        else {
                            // N_b(i) = N_b(i+1)*loop + B_b(i)*move   (N_b(Ls) = 0)
                            const float f = pow2i(xNg - g);
                            xNv = (i == Ls) ? 0.f : xNv * f * ploop + bp * pmove;
                            xNg = g;
                            rNB[i] = xNv; rNBg[i] = g;
                        }
                    }
                }
                if (ALIGN) { tMw -= 32 * C; tIw -= 32 * C; }
                if (((tF - 1) & 7) == 0 && tF > 1) {  // next step enters the previous exponent block
                    if (W_EXP & 2) __syncwarp();   // (not if-convertible: the block becomes a branch taken once per 8 steps)
                    gFc = gFnext;
                    gblk--;
                    gFnext = gFs[GF_AT(max(gblk - 1, 0))];
                    fac = post_scale(gFc, g);
                }
            };
            auto brescale = [&](const int tp) {
                {
                    float mx = 0.f;
#pragma unroll
                    for (int c = 0; c < C; c++) mx = fmaxf(mx, fmaxf(sM[c], fmaxf(sI[c], sD[c])));
                    // the running B(i) partial carries the mass of every strip to the right, which can be far above
                    // this strip's own cells: it must drive the exponent too (cells that flush are negligible)
                    mx = fmaxf(mx, bp);
                    if (!LANE_EXP) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
                    }
                    int e_need = (mx > 1048576.f) ? fexp(mx) : 0;
                    if (!lastS) {   // exponent the right strip had 8 rows ahead (that block is already in the ring)
                        const int ic = max(Ls - tp, 0), r = max(0, Ls - (tp + W_SCALE_EVERY));
                        const int sl = ((r >> 3) == (ic >> 3)) ? bcur : bnd_next(bcur);
                        const int ahead = lds_i1v(bnd_ring + sl * 256 + (r & 7) * 32 + 16) - 40 - g;
                        if (!LANE_EXP || lane == 31) e_need = max(e_need, ahead);
                    }
                    if (LANE_EXP) {   // couple the lanes: g_l >= g_{l+1} - 16
                        int need = g + e_need;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int dn = __shfl_down_sync(FULL, need, o);
                            if (lane + o < 32) need = max(need, dn - 16 * o);
                        }
                        e_need = need - g;
                    }
                    if (e_need > 0) {
                        const float f = pow2i(-e_need);
                        g += e_need;
#pragma unroll
                        for (int c = 0; c < C; c++) { sM[c] *= f; sI[c] *= f; sD[c] *= f; }
                        rMb *= f; bp *= f; ebs *= f;
                        fac = post_scale(gFc, g);
                    }
                }
            };
            {
                const std::integral_constant<bool, false> genc;
                const std::integral_constant<bool, true> allc;
                int tp = 0;
                for (; tp < 32; tp++) {   // ramp-up (nstepsB = Ls + 32 >= 33)
                    bstep(tp, genc);
                    if ((tp & (W_SCALE_EVERY - 1)) == (W_SCALE_EVERY - 1)) brescale(tp);
                }
                for (; tp + 7 <= Ls - 1; tp += 8) {   // steady state: every lane has 1 <= i < Ls; blocks of 8 steps
#pragma unroll W_UNROLL_B
                    for (int u = 0; u < 8; u++) bstep(tp + u, allc);
                    brescale(tp + 7);
                }
                for (; tp < nstepsB; tp++) {   // drain
                    bstep(tp, genc);
                    if ((tp & (W_SCALE_EVERY - 1)) == (W_SCALE_EVERY - 1)) brescale(tp);
                }
            }
            if (!ALIGN) {
                // null2 numerators: sum_k fM(k) * e_k(x) for the canonical symbols (rows 0..K-1 of the table)
                const int K = (E.Kp == 29) ? 20 : 4;
                for (int x = 0; x < K; x++) {
                    float e[C];
                    lds_emis<C>(ebase + x * erow, e);
                    float v = 0.f;
#pragma unroll
                    for (int c = 0; c < C; c++) v = fmaf(accM[c], e[c], v);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                    if (lane == 0) s_n2[w][x] += v;
                }
                float v = 0.f;
#pragma unroll
                for (int c = 0; c < C; c++) v += accM[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                accI += v;  // running sum_k sum_i ppM(i,k)  (warp-uniform)
            }
            if (firstS) { xNv = __shfl_sync(FULL, xNv, 0); xNg = __shfl_sync(FULL, xNg, 0); }
            __syncwarp();
        }
        if (lane == 0 && Wk.dbg_bwd) Wk.dbg_bwd[it.pair] = logf(xNv) + (float)xNg * 0.69314718056f;

        // N / C flank posteriors: ppN(i) = F_N(i-1) B_N(i) loop / T ; ppC(i) = F_C(i-1) B_C(i) loop / T
        // F_N(i) = loop^i, B_C(i) = move * loop^(Ls-i)
        if (ALIGN) {
            const float l2loop = log2f(ploop), l2move = log2f(pmove), l2T = log2f(Tm) + (float)gT;
            for (int base = 1; base <= Ls; base += 32) {
                const int i = base + lane;
                float pn = 0.f, pc = 0.f;
                if (i <= Ls) {
                    // ppN(i) = loop^(i-1) * N_b(i) * loop / T
                    const float nb = rNB[i];
                    pn = (nb > 0.f) ? exp2f((float)i * l2loop + log2f(nb) + (float)rNBg[i] - l2T) : 0.f;
                    if (i >= 2) {
                        const float fc = rFC[i - 1];
                        pc = (fc > 0.f) ? exp2f(log2f(fc) + (float)rFCg[i - 1] + l2move + (float)(Ls - i + 1) * l2loop - l2T) : 0.f;
                    }
                    rNOA[i] = pn; rPPC[i] = pc;
                }
            }
        }

        if (!ALIGN) {
            // ---- null2 by expectation ----
            __syncwarp();
            const int K = (E.Kp == 29) ? 20 : 4;
            const float norm = 1.0f / (float)Ls;
            // log null2 per dense symbol, computed by lanes x < nsym
            float ln2 = 0.f;
            if (lane < Q.nsym) {
                if (lane < K) ln2 = logf((s_n2[w][lane] - accI) * norm + 1.0f);
            }
            __syncwarp();
            if (lane < K) s_n2[w][lane] = (s_n2[w][lane] - accI) * norm + 1.0f;
            __syncwarp();
            if (lane >= K && lane < Q.nsym) {
                // degenerate symbol: unweighted mean of the canonical null2 odds over its members
                const int code = Q.symrow[lane];
                unsigned mask = 0;
                if (E.Kp == 29) {  // amino: B=ND J=IL Z=QE O=K U=C X=all   (codes 21..26)
                    const unsigned m[6] = {(1u << 11) | (1u << 2), (1u << 7) | (1u << 9), (1u << 13) | (1u << 3), 1u << 8, 1u << 1, 0xFFFFFu};
                    mask = (code >= 21 && code <= 26) ? m[code - 21] : 0u;
                } else {  // nucleic: R Y M K S W H B V D N (codes 5..15), bits A=1 C=2 G=4 T=8
                    const unsigned m[11] = {5, 10, 3, 12, 6, 9, 11, 14, 7, 13, 15};
                    mask = (code >= 5 && code <= 15) ? m[code - 5] : 0u;
                }
                float sum = 0.f; int cnt = 0;
                for (int x = 0; x < K; x++) if (mask >> x & 1u) { sum += s_n2[w][x]; cnt++; }
                ln2 = cnt ? logf(sum / (float)cnt) : 0.f;
            }
            // domain correction = sum over envelope residues of ln null2[x]
            const unsigned sresp = sres;
            float dc = 0.f;
            for (int base = 0; base < Ls; base += 32) {
                const int p = base + lane;
                const int xr = (p < Ls) ? RES_AT(p) : 0;
                for (int z = 0; z < 32 && base + z < Ls; z++) {
                    const int xx = __shfl_sync(FULL, xr, z);
                    dc += __shfl_sync(FULL, ln2, xx);
                }
            }
            if (lane == 0) { Wk.envsc[it.pair] = fwd_nats; Wk.domcorr[it.pair] = dc; }
            __syncwarp();
            continue;
        }

        // ======================================= Optimal accuracy (align mode) =======================================
        if (ALIGN) {
            __syncwarp();
            // N_oa(i) = sum_{i'<=i} ppN(i') : in-place inclusive prefix sum over rNOA[1..Ls]; rNOA[0] = 0
            {
                float carry = 0.f;
                if (lane == 0) rNOA[0] = 0.f;
                for (int base = 1; base <= Ls; base += 32) {
                    const int i = base + lane;
                    float v = (i <= Ls) ? rNOA[i] : 0.f;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { float u = __shfl_up_sync(FULL, v, o); if (lane >= o) v += u; }
                    v += carry;
                    if (i <= Ls) rNOA[i] = v;
                    carry = __shfl_sync(FULL, v, 31);
                }
            }
            __syncwarp();
            const float NEG = -INFINITY;
            const int Qs = max(2, (Mh - 1) / 4 + 1);  // HMMER's striped segment count, for select_e tie order
            for (int s = 0; s < nstrips; s++) {
                const long long k0 = (long long)s * SW + lane * C;
                float pa[C], pb[C], pg[C], pmd[C], pdd[C], pmi[C], pii[C], pen[C];
                load_cols<C>(E.tMM + po, k0, pa); load_cols<C>(E.tIM + po, k0, pb); load_cols<C>(E.tDM + po, k0, pg);
                load_cols<C>(E.tMD + po, k0, pmd); load_cols<C>(E.tDD + po, k0, pdd);
                load_cols<C>(E.tMI + po, k0 + 1, pmi); load_cols<C>(E.tII + po, k0 + 1, pii);
                load_cols<C>(E.entry + po, k0 + 1, pen);
                float oM[C], oI[C], oD[C];
#pragma unroll
                for (int c = 0; c < C; c++) { oM[c] = NEG; oI[c] = NEG; oD[c] = NEG; }
                float rM = NEG, rI = NEG, rD = NEG;
                float eV = NEG; int eK = 0;  // running select_e candidate of the lane's last row: eK = k | isD<<30
                const bool first = (s == 0), last = (s == nstrips - 1);
                float *tM = tileM + (size_t)s * TT * SW, *tI = tileI + (size_t)s * TT * SW;
                unsigned *tb = bits + (size_t)s * TT * 32;
                for (int t = 1; t <= nsteps; t++) {
                    const int i = t - lane;
                    const bool act = (i >= 1 && i <= Ls);
                    float cM = __shfl_up_sync(FULL, oM[C - 1], 1);
                    float cI = __shfl_up_sync(FULL, oI[C - 1], 1);
                    float cD = __shfl_up_sync(FULL, oD[C - 1], 1);
                    float cEV = __shfl_up_sync(FULL, eV, 1);
                    int cEK = __shfl_up_sync(FULL, eK, 1);
                    if (lane == 0) {
                        if (!first && act) { cM = BND_M(i); cI = BND_I(i); cD = BND_D(i); cEV = BND_E(i); cEK = BND_G(i); }
                        else { cM = NEG; cI = NEG; cD = NEG; cEV = NEG; cEK = 0; }
                    }
                    if (act) {
                        const float xB = rNOA[i - 1];
                        const float *fm = tM + (size_t)t * 32 * C + lane * 4, *fi = tI + (size_t)t * 32 * C + lane * 4;
                        float pM[C], pI[C];
#pragma unroll
                        for (int v = 0; v < C / 4; v++) {
                            const float4 a = *reinterpret_cast<const float4 *>(fm + 128 * v), b = *reinterpret_cast<const float4 *>(fi + 128 * v);
                            pM[4 * v] = a.x; pM[4 * v + 1] = a.y; pM[4 * v + 2] = a.z; pM[4 * v + 3] = a.w;
                            pI[4 * v] = b.x; pI[4 * v + 1] = b.y; pI[4 * v + 2] = b.z; pI[4 * v + 3] = b.w;
                        }
                        float nM[C], nI[C], nD[C];
                        unsigned word = 0;
#pragma unroll
                        for (int c = C - 1; c >= 0; c--) {
                            const float pm = c > 0 ? oM[c - 1] : rM, pi = c > 0 ? oI[c - 1] : rI, pd = c > 0 ? oD[c - 1] : rD;
                            const long long kcol = k0 + 1 + c;
                            const bool valid = kcol <= Mh;
                            // DP value: zero-probability transitions contribute the constant 0.0 (HMMER's masking)
                            float sv = pen[c] > 0.f ? xB : 0.f;
                            sv = fmaxf(sv, pa[c] > 0.f ? pm : 0.f);
                            sv = fmaxf(sv, pb[c] > 0.f ? pi : 0.f);
                            sv = fmaxf(sv, pg[c] > 0.f ? pd : 0.f);
                            nM[c] = valid ? sv + pM[c] : NEG;
                            // traceback choice (select_m): order M, I, D, B; zero-probability transitions are -inf
                            const float q0 = pa[c] != 0.f ? pm : NEG, q1 = pb[c] != 0.f ? pi : NEG,
                                        q2 = pg[c] != 0.f ? pd : NEG, q3 = pen[c] != 0.f ? xB : NEG;
                            int best = 0; float bv = q0;
                            if (q1 > bv) { bv = q1; best = 1; }
                            if (q2 > bv) { bv = q2; best = 2; }
                            if (q3 > bv) { bv = q3; best = 3; }
                            // I(i,k)
                            float iv = pmi[c] > 0.f ? oM[c] : 0.f;
                            iv = fmaxf(iv, pii[c] > 0.f ? oI[c] : 0.f);
                            nI[c] = (valid && kcol < Mh) ? iv + pI[c] : NEG;
                            const float i0v = pmi[c] != 0.f ? oM[c] : NEG, i1v = pii[c] != 0.f ? oI[c] : NEG;
                            const int ibit = (i1v > i0v) ? 1 : 0;
                            word |= ((unsigned)best | ((unsigned)ibit << 2)) << (4 * c);
                        }
#pragma unroll
                        for (int c = 0; c < C; c++) {
                            const float ml = c > 0 ? nM[c - 1] : cM, dlv = c > 0 ? nD[c - 1] : cD;
                            const long long kcol = k0 + 1 + c;
                            float dv = pmd[c] > 0.f ? ml : 0.f;
                            dv = fmaxf(dv, pdd[c] > 0.f ? dlv : 0.f);
                            nD[c] = (kcol >= 2 && kcol <= Mh) ? dv : NEG;
                            const float d0 = pmd[c] != 0.f ? ml : NEG, d1 = pdd[c] != 0.f ? dlv : NEG;
                            word |= ((d1 > d0) ? 8u : 0u) << (4 * c);
                        }
                        tb[(size_t)t * 32 + lane] = word;
                        // select_e candidates
                        float bV = cEV; int bK = cEK;
#pragma unroll
                        for (int c = 0; c < C; c++) {
                            const int kcol = (int)(k0 + 1 + c);
                            if (kcol <= Mh) {
                                const int ord = ((kcol - 1) % Qs) * 4 + (kcol - 1) / Qs;
                                const int bord = (((bK & 0x3fffffff) - 1) % Qs) * 4 + ((bK & 0x3fffffff) - 1) / Qs;
                                if ((bK & 0x3fffffff) == 0 || oa_e_better(nM[c], 0, ord, bV, bK >> 30, bord)) { bV = nM[c]; bK = kcol; }
                                const int bord2 = (((bK & 0x3fffffff) - 1) % Qs) * 4 + ((bK & 0x3fffffff) - 1) / Qs;
                                if (oa_e_better(nD[c], 1, ord, bV, bK >> 30, bord2)) { bV = nD[c]; bK = kcol | (1 << 30); }
                            }
                        }
                        eV = bV; eK = bK;
#pragma unroll
                        for (int c = 0; c < C; c++) { oM[c] = nM[c]; oI[c] = nI[c]; oD[c] = nD[c]; }
                        rM = cM; rI = cI; rD = cD;
                        if (lane == 31) {
                            if (!last) { BND_M(i) = oM[C - 1]; BND_I(i) = oI[C - 1]; BND_D(i) = oD[C - 1]; BND_E(i) = eV; BND_G(i) = eK; }
                            else { rEOA[i] = eV; rKE[i] = eK; }
                        }
                    }
                }
                __syncwarp();
            }
            __syncwarp();
            // ---- traceback (lane 0; SURVEY 8a "Align semantics" item 4) ----
            {
                int *colsw = Wk.cols + Wk.col_off[it.pair];
                for (int z = lane; z < Ls; z += 32) colsw[z] = -1;
            }
            __syncwarp();
            if (lane == 0) {
                int *cols = Wk.cols + Wk.col_off[it.pair];
                // C_oa(i) = max(C_oa(i-1) + ppC(i), E_oa(i)); C_oa(0) = -inf : computed forward into rFC (reuse)
                float c = NEG;
                for (int i = 1; i <= Ls; i++) {
                    float a = c + rPPC[i];
                    float e = rEOA[i];
                    c = fmaxf(a, e);
                    rFC[i] = c;
                }
                rFC[0] = NEG;
                int i = Ls, k = 0, st = 4;  // 0=M 1=I 2=D 3=B 4=C 5=E
                int guard = 2 * (Ls + Mh) + 8;  // a valid trace never needs more steps
                while (st != 3 && guard-- > 0) {
                    if (st != 4 && st != 5 && (k < 1 || i < 1 || k > Mh)) break;
                    if (st == 4) {
                        if (i == 0) break;
                        const float a = rFC[i - 1] + rPPC[i], e = rEOA[i];
                        if (a >= e) i--; else st = 5;
                    } else if (st == 5) {
                        const int ke = rKE[i];
                        k = ke & 0x3fffffff; st = (ke >> 30) ? 2 : 0;
                    } else {
                        const int s = (k - 1) / SW, r = (k - 1) - s * SW, l = r / C, cc = r - l * C;
                        const unsigned wd = bits[((size_t)s * TT + (i + l)) * 32 + l] >> (4 * cc);
                        if (st == 0) { cols[i - 1] = k - 1; st = (int)(wd & 3u); k--; i--; }
                        else if (st == 2) { st = (wd & 8u) ? 2 : 0; k--; }
                        else { st = (wd & 4u) ? 1 : 0; i--; }
                    }
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace witch
