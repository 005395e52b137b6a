// Device-side views shared by all kernels (plain structs passed by value at launch).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// Kernel launch and dynamic shared memory go through two macros so that the same sources also build as a host-side
// SIMT simulation (tools/sim: TEST TOOLING, g++ -DWITCH_HOST_SIM; the product library is always the nvcc build).
#ifndef WITCH_HOST_SIM
#define WITCH_LAUNCH(kernel, ...) kernel<<<__VA_ARGS__>>>
#define WITCH_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

namespace witch {

constexpr int MAX_SYM = 32;   // emission rows a query set may use (Kp of amino is 29)
constexpr int MAX_ENV = 6;    // envelopes kept per (query, HMM) pair
constexpr float LOG2E_F = 1.4426950408889634f;

// Ensemble of profiles in HBM: nine per-node float arrays (probability space), one emission-odds table.
// Layout: every HMM h owns a zero-padded slice [poff[h], poff[h] + stride[h]) of each transition array
// (index = node k, node 0 and nodes > M are all-zero) and a [Kp][stride[h]] block of `emis` at eoff[h].
struct DevEhmm {
    const float *tMM, *tMI, *tMD, *tIM, *tII, *tDM, *tDD, *entry;
    const float *gD;  // gD[k] = 1 + tDD[k]*gD[k+1] (Backward D-state response to a unit E exit; used by the parser)
    const float *emis;
    const int *M;
    const int *stride;
    const long long *poff;
    const long long *eoff;
    // hmmsearch's own striped float tables (hmm_profile.cpp:build_striped), read by the multi-domain branch only
    const float *otfv;       // per HMM at otoff[h]: [(7*Q + Q)][4]
    const float *orfv;       // per HMM at oroff[h]: [Kp][Q][4]
    const long long *otoff;
    const long long *oroff;
    const int *oQ;
    // the same values indexed by NODE for the trace walk of that branch (one 128-bit load per step instead of scattered
    // scalars): ont8[onoff[h] + k] = {BM,MM,IM,DM into node k | MD,DD,MI,II of node k}, onem[(onoff[h] + k) * KE + x] =
    // match odds of canonical residue x at node k (KE = 4 or 20)
    const float *ont8;
    const float *onem;
    const long long *onoff;
    int H;
    int Kp;
};

// Query set in HBM: residues as dense symbol codes (one byte each, 0..nsym-1), concatenated.
struct DevQueries {
    const uint8_t *dsq;
    const long long *off;  // [n+1]
    const int *len;        // [n]
    int n;
    int nsym;              // distinct symbols present in the set
    int symrow[MAX_SYM];   // dense code -> alphabet symbol code (row of the emission table)
};

// Result of the multihit parser pass for one pair.
struct PairParse {
    float fwd_bits;        // log2 of the multihit Forward probability (odds space)
    int nenv;              // number of envelopes (regions); 0 => pair is not reported
    int flags;             // bit 0: some region failed the single-domain test; bit 8+r: region r did; bit 2: > MAX_ENV regions
    int env_i[MAX_ENV];    // 1-based inclusive envelope coordinates
    int env_j[MAX_ENV];
};

}  // namespace witch
