// Static feasibility study (compile-only, never linked into the library): the Forward wavefront step of wave_kernel with
// TWO items per warp -- the same C = 4 columns of two queries of one HMM in the two halves of every f32x2 register pair.
// Question answered by `nvcc -Xptxas -v` + `cuobjdump -sass` (tools/proto/README.md): does ptxas keep the pairs packed
// (FFMA2/FMUL2/FADD2 without re-pairing MOVs) inside the kernel's register budget?
#include <cuda_runtime.h>
#include <cstdint>

#ifndef PC
#define PC 4
#endif
constexpr int C = PC;

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 shfl_up2(float2 v) {
    return f2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ float4 lds4(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int ldsu8(unsigned a) { int v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

struct Params {  // per-lane transition parameters of the C owned columns, duplicated into both halves
    const float *tMM, *tIM, *tDM, *tMD, *tDD, *tMI, *tII, *ent;
    float *tile;      // [step][lane][c][item] floats
    float *out;
    int Ls;           // common number of rows (steady state only)
    unsigned emis_sa; // shared-memory emission table [sym][32 lanes][C]
    unsigned erow;
    unsigned resA, resB;  // shared-memory residue strings of the two items
    float pmoveA, pmoveB;
};

__global__ void __launch_bounds__(128, 3) pair_forward(Params P) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31;
    float2 pa[C], pb[C], pg[C], pmd[C], pdd[C], pmi[C], pii[C], pen[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int k = lane * C + c;
        pa[c] = f2(P.tMM[k], P.tMM[k]); pb[c] = f2(P.tIM[k], P.tIM[k]); pg[c] = f2(P.tDM[k], P.tDM[k]);
        pmd[c] = f2(P.tMD[k], P.tMD[k]); pdd[c] = f2(P.tDD[k], P.tDD[k]); pmi[c] = f2(P.tMI[k], P.tMI[k]);
        pii[c] = f2(P.tII[k], P.tII[k]); pen[c] = f2(P.ent[k], P.ent[k]);
    }
    float2 sM[C], sI[C], sD[C];
#pragma unroll
    for (int c = 0; c < C; c++) { sM[c] = f2(0.f, 0.f); sI[c] = f2(0.f, 0.f); sD[c] = f2(0.f, 0.f); }
    float2 rM = f2(0.f, 0.f), rI = rM, rD = rM, ep = rM;
    float2 xBs = f2(P.pmoveA, P.pmoveB);
    const float2 ploop = f2(1.f - P.pmoveA, 1.f - P.pmoveB);
    const unsigned ebase = P.emis_sa + lane * C * 4;
    float *tp = P.tile + lane * C * 2;
    int xa = ldsu8(P.resA), xb = ldsu8(P.resB);
    for (int t = 33; t + 7 <= P.Ls; t += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int i = t + u - lane;
            float2 cM = shfl_up2(sM[C - 1]), cI = shfl_up2(sI[C - 1]), cD = shfl_up2(sD[C - 1]), cE = shfl_up2(ep);
            if (lane == 0) { cM = f2(0.f, 0.f); cI = cM; cD = cM; cE = cM; }
            const float4 eA = lds4(ebase + xa * P.erow), eB = lds4(ebase + xb * P.erow);
            xa = ldsu8(P.resA + i); xb = ldsu8(P.resB + i);
            const float ea[4] = {eA.x, eA.y, eA.z, eA.w}, eb[4] = {eB.x, eB.y, eB.z, eB.w};
            float2 nM[C], nI[C], nD[C];
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                const float2 pm = c > 0 ? sM[c - 1] : rM, pi = c > 0 ? sI[c - 1] : rI, pd = c > 0 ? sD[c - 1] : rD;
                nI[c] = __ffma2_rn(sM[c], pmi[c], __fmul2_rn(sI[c], pii[c]));
                float2 acc = __fmul2_rn(xBs, pen[c]);
                acc = __ffma2_rn(pm, pa[c], acc); acc = __ffma2_rn(pi, pb[c], acc); acc = __ffma2_rn(pd, pg[c], acc);
                nM[c] = f2(acc.x * ea[c], acc.y * eb[c]);
            }
            nD[0] = __ffma2_rn(cD, pdd[0], __fmul2_rn(cM, pmd[0]));
#pragma unroll
            for (int c = 1; c < C; c++) nD[c] = __ffma2_rn(nD[c - 1], pdd[c], __fmul2_rn(nM[c - 1], pmd[c]));
            float2 es = cE;
#pragma unroll
            for (int c = 0; c < C; c++) { sM[c] = nM[c]; sI[c] = nI[c]; sD[c] = nD[c]; es = __fadd2_rn(es, __fadd2_rn(nM[c], nD[c])); }
            ep = es;
            rM = cM; rI = cI; rD = cD;
            xBs = __fmul2_rn(xBs, ploop);
#pragma unroll
            for (int v = 0; v < C / 2; v++)
                *reinterpret_cast<float4 *>(tp + 4 * v) = make_float4(nM[2 * v].x, nM[2 * v].y, nM[2 * v + 1].x, nM[2 * v + 1].y);
            tp += 32 * C * 2;
        }
    }
    float2 acc = ep;
#pragma unroll
    for (int c = 0; c < C; c++) acc = __fadd2_rn(acc, __fadd2_rn(sM[c], __fadd2_rn(sI[c], sD[c])));
    P.out[threadIdx.x + blockIdx.x * blockDim.x] = acc.x + acc.y + rM.x + rI.y + rD.x;
}

// ---- Backward + posterior accumulation (envelope mode) for two items per warp, steady state ----
__device__ __forceinline__ float2 shfl_dn2(float2 v) {
    return f2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ float lds1(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }

struct ParamsB {
    const float *tMM, *tIM, *tDM, *tMD, *tDD, *tMI, *tII, *ent;
    float *out;
    int Ls;
    unsigned emis_sa, erow, eright, resA, resB, ring_sa;   // ring: stored Forward rows [stage][lane][c][item]
    float pmoveA, pmoveB, facA, facB;
};

__global__ void __launch_bounds__(128, 3) pair_backward(ParamsB P) {
    const int lane = threadIdx.x & 31;
    float2 oMM[C], oIM[C], oDM[C], oMD[C], oDD[C], oMI[C], oII[C], pen[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int k = lane * C + c;
        oMM[c] = f2(P.tMM[k], P.tMM[k]); oIM[c] = f2(P.tIM[k], P.tIM[k]); oDM[c] = f2(P.tDM[k], P.tDM[k]);
        oMD[c] = f2(P.tMD[k], P.tMD[k]); oDD[c] = f2(P.tDD[k], P.tDD[k]); oMI[c] = f2(P.tMI[k], P.tMI[k]);
        oII[c] = f2(P.tII[k], P.tII[k]); pen[c] = f2(P.ent[k], P.ent[k]);
    }
    float2 sM[C], sI[C], sD[C], accM[C];
#pragma unroll
    for (int c = 0; c < C; c++) { sM[c] = f2(0.f, 0.f); sI[c] = sM[c]; sD[c] = sM[c]; accM[c] = sM[c]; }
    float2 rMb = f2(0.f, 0.f), bp = rMb;
    float2 ebs = f2(P.pmoveA, P.pmoveB);
    const float2 ploop = f2(1.f - P.pmoveA, 1.f - P.pmoveB), fac = f2(P.facA, P.facB);
    const unsigned ebase = P.emis_sa + lane * C * 4;
    int xa = ldsu8(P.resA), xb = ldsu8(P.resB);
    int stage = 0;
    for (int tp = 32; tp + 7 <= P.Ls - 1; tp += 8) {
#pragma unroll 2
        for (int u = 0; u < 8; u++) {
            const int i = P.Ls - (tp + u) + lane;
            float2 cMb = shfl_dn2(sM[0]), cDb = shfl_dn2(sD[0]), cB = shfl_dn2(bp);
            if (lane == 31) { cMb = f2(0.f, 0.f); cDb = cMb; cB = cMb; }
            const float4 eA = lds4(ebase + xa * P.erow), eB = lds4(ebase + xb * P.erow);
            const float erA = lds1(P.eright + xa * P.erow), erB = lds1(P.eright + xb * P.erow);
            xa = ldsu8(P.resA + i); xb = ldsu8(P.resB + i);
            const float ea[4] = {eA.x, eA.y, eA.z, eA.w}, eb[4] = {eB.x, eB.y, eB.z, eB.w};
            float2 mn[C];
#pragma unroll
            for (int c = 0; c < C; c++) mn[c] = f2(sM[c].x * ea[c], sM[c].y * eb[c]);
            const float2 mnR = f2(rMb.x * erA, rMb.y * erB);
            float2 bs = cB;
#pragma unroll
            for (int c = 0; c < C; c++) bs = __ffma2_rn(mn[c], pen[c], bs);
            bp = bs;
            float2 nM[C], nI[C], nD[C];
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                const float2 m1 = (c < C - 1) ? mn[c + 1] : mnR, dr = (c < C - 1) ? nD[c + 1] : cDb;
                nD[c] = __ffma2_rn(dr, oDD[c], __ffma2_rn(m1, oDM[c], ebs));
                nM[c] = __ffma2_rn(m1, oMM[c], __ffma2_rn(sI[c], oMI[c], __ffma2_rn(dr, oMD[c], ebs)));
                nI[c] = __ffma2_rn(m1, oIM[c], __fmul2_rn(sI[c], oII[c]));
            }
            const unsigned rs = P.ring_sa + stage * (32 * C * 8) + lane * 16;
            stage = stage == 2 ? 0 : stage + 1;
#pragma unroll
            for (int v = 0; v < C / 2; v++) {
                const float4 a = lds4(rs + v * 512);
                accM[2 * v] = __ffma2_rn(__fmul2_rn(f2(a.x, a.y), nM[2 * v]), fac, accM[2 * v]);
                accM[2 * v + 1] = __ffma2_rn(__fmul2_rn(f2(a.z, a.w), nM[2 * v + 1]), fac, accM[2 * v + 1]);
            }
#pragma unroll
            for (int c = 0; c < C; c++) { sM[c] = nM[c]; sI[c] = nI[c]; sD[c] = nD[c]; }
            rMb = cMb;
            ebs = __fmul2_rn(ebs, ploop);
        }
    }
    float2 acc = bp;
#pragma unroll
    for (int c = 0; c < C; c++) acc = __fadd2_rn(acc, __fadd2_rn(accM[c], __fadd2_rn(sM[c], __fadd2_rn(sI[c], sD[c]))));
    P.out[threadIdx.x + blockIdx.x * blockDim.x] = acc.x + acc.y + rMb.x;
}
