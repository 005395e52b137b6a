// Host-side HMMER3/f reader and local-mode Plan-7 profile configuration (product code, C++).
// Replaces what hmmsearch/hmmalign do after opening the model file the reference hands them
// (witch_msa/gcmm/algorithm.py:526-532, gcmm/aligner.py:98-100); semantics per SURVEY.md 8(a) items 1-2.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace witch {

enum Alphabet { ALPH_DNA = 0, ALPH_RNA = 1, ALPH_AMINO = 2 };

struct AlphabetInfo {
    int type;
    int K;                 // canonical residues
    int Kp;                // all symbols incl. gap, degenerate, '*', '~'
    std::string syms;      // symbol order (Easel's)
    int8_t code[256];      // ASCII -> symbol code, -1 invalid
    std::vector<double> bg;             // [K] background
    std::vector<std::vector<int>> degen;  // [Kp] canonical members of each symbol (single member for canonical)
};

const AlphabetInfo &alphabet_info(int type);

// One profile in probability space, float, padded: arrays are indexed by node k = 0..Mpad (Mpad+1 entries + slack).
struct HostProfile {
    int M = 0, nseq = 0, alph = 0;
    std::string name;
    // transitions out of node k (k = 1..M-1 non-zero; node 0 and node M are zero as in p7_ProfileConfig)
    std::vector<float> tMM, tMI, tMD, tIM, tII, tDM, tDD;
    std::vector<float> entry;            // B->M_k (occupancy-weighted local entry), k = 1..M
    std::vector<float> gD;               // gD[k] = 1 + tDD[k]*gD[k+1]: Backward D_k response to a unit E exit (model constant)
    std::vector<float> emis;             // [Kp][stride] match odds ratios; insert odds == 1
    int stride = 0;                      // row stride of emis (= padded length)
    // The same model as hmmsearch's own float "optimized profile" (4-way striped vectors, Q = max(2, (M-1)/4+1) per
    // row), value for value: used only by the multi-domain branch of the domain definition (md_kernel.cuh), whose
    // sampled traces depend on FP32 comparisons and therefore on HMMER's exact parameter rounding.
    int Q = 0;
    std::vector<float> otfv;             // [(7*Q + Q)][4]: BM,MM,IM,DM,MD,MI,II per q, then the Q DD vectors
    std::vector<float> orfv;             // [Kp][Q][4] match odds
};

// Parse + configure. Throws std::runtime_error with a message on malformed input.
HostProfile load_profile(const std::string &path, int pad_to);

// Serialised configured profiles next to the HMM text ("profile cache", SURVEY.md 8f-3: a `-p` re-run of the reference
// re-reads the same hmmbuild.model.* files, gcmm/loader.py:17-58): one binary file holding every array of every
// HostProfile plus the size and modification time of each source file. load_profile_cache returns false (and leaves
// `out` empty) when the file is missing, malformed, from another version, or any source file changed.
bool load_profile_cache(const std::string &cache_path, const std::vector<std::string> &hmm_paths, std::vector<HostProfile> &out);
// Atomic (temporary file + rename); returns false when the file cannot be written (the caller carries on without it).
bool save_profile_cache(const std::string &cache_path, const std::vector<std::string> &hmm_paths, const std::vector<HostProfile> &ps);

}  // namespace witch
