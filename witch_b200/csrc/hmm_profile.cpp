// HMMER3/f ASCII profile reader + local-mode configuration (host side of the product path).
// Follows SURVEY.md 8(a) "Score semantics" items 1-2 and Appendix A.1 (format), which describe what the
// reference's HMMER 3.1b2 binaries do with the file written at witch_msa/gcmm/algorithm.py:463-470.
#include "hmm_profile.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>
#include <stdexcept>

namespace witch {

static AlphabetInfo make_alphabet(int type) {
    AlphabetInfo a;
    a.type = type;
    std::memset(a.code, -1, sizeof(a.code));
    std::vector<std::pair<char, std::string>> deg;
    std::string canon;
    if (type == ALPH_AMINO) {
        a.syms = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~";
        a.K = 20;
        canon = a.syms.substr(0, 20);
        static const double f[20] = {0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198,
                                     0.0590092, 0.0594422, 0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639,
                                     0.0540978, 0.0683364, 0.0540687, 0.0673417, 0.0114135, 0.0304133};
        a.bg.assign(f, f + 20);
        deg = {{'B', "ND"}, {'J', "IL"}, {'Z', "QE"}, {'O', "K"}, {'U', "C"}, {'X', canon}};
    } else {
        const char t = (type == ALPH_RNA) ? 'U' : 'T';
        a.syms = std::string("ACG") + t + "-RYMKSWHBVDN*~";
        a.K = 4;
        canon = a.syms.substr(0, 4);
        a.bg.assign(4, 0.25);
        std::string T(1, t);
        deg = {{'R', "AG"},      {'Y', "C" + T},    {'M', "AC"},       {'K', "G" + T},
               {'S', "CG"},      {'W', "A" + T},    {'H', "AC" + T},   {'B', "CG" + T},
               {'V', "ACG"},     {'D', "AG" + T},   {'N', "ACG" + T}};
    }
    a.Kp = (int)a.syms.size();
    for (int i = 0; i < a.Kp; i++) {
        unsigned char c = (unsigned char)a.syms[i];
        a.code[c] = (int8_t)i;
        if (c >= 'A' && c <= 'Z') a.code[c - 'A' + 'a'] = (int8_t)i;
    }
    auto equiv = [&](char from, char to) {
        a.code[(unsigned char)from] = a.code[(unsigned char)to];
        if (from >= 'A' && from <= 'Z') a.code[(unsigned char)(from - 'A' + 'a')] = a.code[(unsigned char)to];
    };
    if (type == ALPH_DNA) { equiv('U', 'T'); equiv('X', 'N'); equiv('I', 'A'); }
    if (type == ALPH_RNA) { equiv('T', 'U'); equiv('X', 'N'); equiv('I', 'A'); }
    equiv('_', '-');
    equiv('.', '-');
    a.degen.assign(a.Kp, {});
    for (int i = 0; i < a.K; i++) a.degen[i] = {i};
    for (auto &d : deg) {
        int x = a.code[(unsigned char)d.first];
        for (char c : d.second) a.degen[x].push_back((int)canon.find(c));
    }
    return a;
}

const AlphabetInfo &alphabet_info(int type) {
    static AlphabetInfo infos[3];
    static std::once_flag once;
    std::call_once(once, [] { for (int t = 0; t < 3; t++) infos[t] = make_alphabet(t); });
    if (type < 0 || type > 2) throw std::runtime_error("bad alphabet type");
    return infos[type];
}

static double parse_prob(const std::string &tok) {
    if (tok == "*") return 0.0;
    return std::exp(-std::strtod(tok.c_str(), nullptr));
}

static std::vector<std::string> split(const std::string &s) {
    std::vector<std::string> out;
    std::istringstream is(s);
    std::string t;
    while (is >> t) out.push_back(t);
    return out;
}

HostProfile load_profile(const std::string &path, int /*pad_to*/) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open HMM file: " + path);
    HostProfile p;
    std::string line;
    if (!std::getline(in, line) || line.compare(0, 7, "HMMER3/") != 0)
        throw std::runtime_error("not a HMMER3 ASCII profile: " + path);
    int M = -1, alph = -1;
    bool have_hmm = false;
    while (std::getline(in, line)) {
        auto tok = split(line);
        if (tok.empty()) continue;
        if (tok[0] == "NAME" && tok.size() > 1) p.name = tok[1];
        else if (tok[0] == "LENG" && tok.size() > 1) M = std::atoi(tok[1].c_str());
        else if (tok[0] == "NSEQ" && tok.size() > 1) p.nseq = std::atoi(tok[1].c_str());
        else if (tok[0] == "ALPH" && tok.size() > 1) {
            std::string a = tok[1];
            for (auto &c : a) c = (char)std::tolower(c);
            alph = (a == "dna") ? ALPH_DNA : (a == "rna") ? ALPH_RNA : (a == "amino") ? ALPH_AMINO : -1;
        } else if (tok[0] == "HMM") { have_hmm = true; break; }
    }
    if (!have_hmm || M <= 0 || alph < 0) throw std::runtime_error("bad HMM header in " + path);
    const AlphabetInfo &A = alphabet_info(alph);
    const int K = A.K, Kp = A.Kp;
    p.M = M;
    p.alph = alph;
    std::getline(in, line);  // transition header
    std::getline(in, line);
    auto tok = split(line);
    if (!tok.empty() && tok[0] == "COMPO") std::getline(in, line);  // now node-0 insert emissions (ignored)
    std::getline(in, line);                                         // node-0 transitions
    tok = split(line);
    if (tok.size() < 7) throw std::runtime_error("bad node-0 transition line in " + path);
    std::vector<double> t((size_t)(M + 1) * 7, 0.0), mat((size_t)(M + 1) * K, 0.0);
    for (int x = 0; x < 7; x++) t[x] = parse_prob(tok[x]);
    for (int k = 1; k <= M; k++) {
        if (!std::getline(in, line)) throw std::runtime_error("truncated HMM file " + path);
        tok = split(line);
        if ((int)tok.size() < K + 1 || std::atoi(tok[0].c_str()) != k)
            throw std::runtime_error("bad match line at node " + std::to_string(k) + " in " + path);
        for (int x = 0; x < K; x++) mat[(size_t)k * K + x] = parse_prob(tok[1 + x]);
        std::getline(in, line);  // insert emissions: ignored, insert odds are hard-wired to 1
        if (!std::getline(in, line)) throw std::runtime_error("truncated HMM file " + path);
        tok = split(line);
        if (tok.size() < 7) throw std::runtime_error("bad transition line at node " + std::to_string(k));
        for (int x = 0; x < 7; x++) t[(size_t)k * 7 + x] = parse_prob(tok[x]);
    }
    // occupancy -> local entry distribution
    std::vector<double> occ(M + 1, 0.0);
    occ[1] = t[1] + t[0];
    for (int k = 2; k <= M; k++) {
        const double *tp = &t[(size_t)(k - 1) * 7];
        occ[k] = occ[k - 1] * (tp[0] + tp[1]) + (1.0 - occ[k - 1]) * tp[5];
    }
    double Z = 0;
    for (int k = 1; k <= M; k++) Z += occ[k] * (double)(M - k + 1);
    const int stride = ((M + 1 + 511) / 512) * 512 + 512;
    p.stride = stride;
    auto zeros = [&] { return std::vector<float>((size_t)stride, 0.0f); };
    p.tMM = zeros(); p.tMI = zeros(); p.tMD = zeros(); p.tIM = zeros(); p.tII = zeros(); p.tDM = zeros();
    p.tDD = zeros(); p.entry = zeros();
    for (int k = 1; k < M; k++) {  // node 0 and node M carry no core transitions in a local profile
        const double *tp = &t[(size_t)k * 7];
        p.tMM[k] = (float)tp[0]; p.tMI[k] = (float)tp[1]; p.tMD[k] = (float)tp[2]; p.tIM[k] = (float)tp[3];
        p.tII[k] = (float)tp[4]; p.tDM[k] = (float)tp[5]; p.tDD[k] = (float)tp[6];
    }
    for (int k = 1; k <= M; k++) p.entry[k] = (float)(occ[k] / Z);
    p.gD.assign(stride, 1.0f);  // beyond the model tDD = 0, so the response stays 1 (padded columns never feed back)
    for (int k = M; k >= 0; k--) p.gD[k] = 1.0f + p.tDD[k] * (k + 1 < stride ? p.gD[k + 1] : 0.0f);
    p.emis.assign((size_t)Kp * stride, 0.0f);
    std::vector<double> sc(Kp);
    for (int k = 1; k <= M; k++) {
        for (int x = 0; x < K; x++) sc[x] = std::log(mat[(size_t)k * K + x] / A.bg[x]);
        for (int x = K; x < Kp; x++) {
            const auto &mem = A.degen[x];
            if (mem.empty()) { sc[x] = -INFINITY; continue; }
            double num = 0, den = 0;
            for (int m : mem) { num += sc[m] * A.bg[m]; den += A.bg[m]; }
            sc[x] = num / den;
        }
        for (int x = 0; x < Kp; x++) p.emis[(size_t)x * stride + k] = (float)std::exp(sc[x]);
    }
    return p;
}

}  // namespace witch
