// HMMER3/f ASCII profile reader + local-mode configuration (host side of the product path).
// Follows SURVEY.md 8(a) "Score semantics" items 1-2 and Appendix A.1 (format), which describe what the
// reference's HMMER 3.1b2 binaries do with the file written at witch_msa/gcmm/algorithm.py:463-470.
#include "hmm_profile.h"

#include <sys/stat.h>
#include <unistd.h>
#include <cstdio>

#include <algorithm>
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
static double parse_raw(const std::string &tok) { return tok == "*" ? INFINITY : std::strtod(tok.c_str(), nullptr); }

// exp() the way Easel's vector code evaluates it for the striped tables (Cephes-style range reduction + degree-5
// polynomial, every operation a separate FP32 rounding): the striped parameters must round exactly like hmmsearch's.
static float striped_expf(float x) {
    const float x0 = x;
    float fx = x * 1.44269502f;
    fx = fx + 0.5f;
    float fl = (float)(int)fx;
    if (fx < fl) fl = fl - 1.0f;
    const int n = (int)fl;
    const float hi = fl * 0.693359375f, lo = fl * -2.12194440e-4f;
    x = x - hi;
    x = x - lo;
    const float z = x * x;
    static const uint32_t cbits[5] = {961571175u, 985088974u, 1007192328u, 1026206145u, 1042983594u};
    float c[5];
    std::memcpy(c, cbits, sizeof(c));
    float y = c[0] * x;
    for (int j = 1; j < 5; j++) { y = y + c[j]; y = y * x; }
    y = y + 0.5f;
    y = y * z;
    y = y + x;
    y = y + 1.0f;
    const uint32_t pw = (uint32_t)(n + 127) << 23;
    float p2;
    std::memcpy(&p2, &pw, 4);
    y = y * p2;
    if (x0 > 88.3762588501f) return INFINITY;
    if (x0 <= -88.3762588501f) return 0.0f;
    return y;
}

// hmmsearch's float pipeline from the text numbers to the striped probability tables: p = expf(-x); occupancy and
// local entry in float (the 1 - occ term in double); scores = (float)log(double); tables = striped_expf(score).
static void build_striped(HostProfile &p, const AlphabetInfo &A, const std::vector<double> &raw_t, const std::vector<double> &raw_mat) {
    const int M = p.M, K = A.K, Kp = A.Kp;
    const int Q = std::max(2, (M - 1) / 4 + 1);
    p.Q = Q;
    std::vector<float> t((size_t)(M + 1) * 7), mat((size_t)(M + 1) * K, 0.f);
    for (size_t z = 0; z < t.size(); z++) t[z] = std::isinf(raw_t[z]) ? 0.0f : expf((float)(-1.0 * raw_t[z]));
    for (size_t z = (size_t)K; z < mat.size(); z++) mat[z] = std::isinf(raw_mat[z]) ? 0.0f : expf((float)(-1.0 * raw_mat[z]));
    std::vector<float> occ(M + 2, 0.f);
    occ[1] = t[1] + t[0];
    for (int k = 2; k <= M; k++) {
        const float a = occ[k - 1] * (t[(size_t)(k - 1) * 7 + 0] + t[(size_t)(k - 1) * 7 + 1]);
        occ[k] = (float)((double)a + (1.0 - (double)occ[k - 1]) * (double)t[(size_t)(k - 1) * 7 + 5]);
    }
    float Z = 0.f;
    for (int k = 1; k <= M; k++) Z += occ[k] * (float)(M - k + 1);
    const float NEG = -INFINITY;
    std::vector<float> tsc((size_t)(M + 1) * 8, NEG);   // [k][MM,MI,MD,IM,II,DM,DD, entry into k+1]
    for (int k = 1; k <= M; k++) tsc[(size_t)(k - 1) * 8 + 7] = (float)std::log((double)(occ[k] / Z));
    for (int k = 1; k < M; k++)
        for (int x = 0; x < 7; x++) tsc[(size_t)k * 8 + x] = (float)std::log((double)t[(size_t)k * 7 + x]);
    std::vector<float> bg(K);
    for (int x = 0; x < K; x++) bg[x] = (float)A.bg[x];
    p.orfv.assign((size_t)Kp * Q * 4, 0.f);
    std::vector<float> sc(Kp);
    for (int k = 1; k <= M; k++) {
        for (int x = 0; x < Kp; x++) sc[x] = NEG;
        for (int x = 0; x < K; x++) sc[x] = (float)std::log((double)mat[(size_t)k * K + x] / bg[x]);
        for (int x = K + 1; x <= Kp - 3; x++) {
            float num = 0.f, den = 0.f;
            for (int m : A.degen[x]) { num += sc[m] * bg[m]; den += bg[m]; }
            if (!A.degen[x].empty()) sc[x] = num / den;
        }
        const int q = (k - 1) % Q, z = (k - 1) / Q;
        for (int x = 0; x < Kp; x++) p.orfv[((size_t)x * Q + q) * 4 + z] = striped_expf(sc[x]);
    }
    p.otfv.assign((size_t)8 * Q * 4, 0.f);
    static const int src[7] = {7, 0, 3, 5, 2, 1, 4};   // BM, MM, IM, DM (from node k-1), MD, MI, II (of node k)
    for (int q = 0; q < Q; q++)
        for (int z = 0; z < 4; z++) {
            const int k = q + 1 + z * Q;
            for (int tt = 0; tt < 7; tt++) {
                const int kb = (tt <= 3) ? k - 1 : k;
                p.otfv[((size_t)7 * q + tt) * 4 + z] = striped_expf(kb < M ? tsc[(size_t)kb * 8 + src[tt]] : NEG);
            }
            p.otfv[((size_t)7 * Q + q) * 4 + z] = striped_expf(k < M ? tsc[(size_t)k * 8 + 6] : NEG);
        }
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
    std::vector<double> raw_t((size_t)(M + 1) * 7, INFINITY), raw_mat((size_t)(M + 1) * K, INFINITY);
    for (int x = 0; x < 7; x++) { t[x] = parse_prob(tok[x]); raw_t[x] = parse_raw(tok[x]); }
    for (int k = 1; k <= M; k++) {
        if (!std::getline(in, line)) throw std::runtime_error("truncated HMM file " + path);
        tok = split(line);
        if ((int)tok.size() < K + 1 || std::atoi(tok[0].c_str()) != k)
            throw std::runtime_error("bad match line at node " + std::to_string(k) + " in " + path);
        for (int x = 0; x < K; x++) { mat[(size_t)k * K + x] = parse_prob(tok[1 + x]); raw_mat[(size_t)k * K + x] = parse_raw(tok[1 + x]); }
        std::getline(in, line);  // insert emissions: ignored, insert odds are hard-wired to 1
        if (!std::getline(in, line)) throw std::runtime_error("truncated HMM file " + path);
        tok = split(line);
        if (tok.size() < 7) throw std::runtime_error("bad transition line at node " + std::to_string(k));
        for (int x = 0; x < 7; x++) { t[(size_t)k * 7 + x] = parse_prob(tok[x]); raw_t[(size_t)k * 7 + x] = parse_raw(tok[x]); }
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
    build_striped(p, A, raw_t, raw_mat);
    return p;
}

// ---------------------------------------------------------------------------------------------------------------
// profile cache
namespace {
const char CACHE_MAGIC[8] = {'W', 'B', '2', 'E', 'H', 'M', 'M', '2'};
struct SrcStamp { long long size, mtime_ns; };
bool stamp_of(const std::string &path, SrcStamp &st) {
    struct stat sb;
    if (stat(path.c_str(), &sb) != 0) return false;
    st.size = (long long)sb.st_size;
    st.mtime_ns = (long long)sb.st_mtim.tv_sec * 1000000000LL + (long long)sb.st_mtim.tv_nsec;
    return true;
}
template <typename T> bool wr(FILE *f, const T &v) { return fwrite(&v, sizeof(T), 1, f) == 1; }
template <typename T> bool rd(FILE *f, T &v) { return fread(&v, sizeof(T), 1, f) == 1; }
bool wr_vec(FILE *f, const std::vector<float> &v) {
    const unsigned long long n = v.size();
    return wr(f, n) && (n == 0 || fwrite(v.data(), sizeof(float), n, f) == n);
}
bool rd_vec(FILE *f, std::vector<float> &v, unsigned long long max_n) {
    unsigned long long n = 0;
    if (!rd(f, n) || n > max_n) return false;
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(float), n, f) == n;
}
std::vector<float> HostProfile::*const CACHE_FIELDS[] = {&HostProfile::tMM, &HostProfile::tMI, &HostProfile::tMD, &HostProfile::tIM,
                                                        &HostProfile::tII, &HostProfile::tDM, &HostProfile::tDD, &HostProfile::entry,
                                                        &HostProfile::gD, &HostProfile::emis, &HostProfile::otfv, &HostProfile::orfv};
}  // namespace

bool save_profile_cache(const std::string &cache_path, const std::vector<std::string> &hmm_paths, const std::vector<HostProfile> &ps) {
    if (hmm_paths.size() != ps.size() || ps.empty()) return false;
    const std::string tmp = cache_path + ".tmp." + std::to_string((long long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    bool ok = fwrite(CACHE_MAGIC, 1, 8, f) == 8;
    const int n = (int)ps.size();
    ok = ok && wr(f, n);
    for (int h = 0; ok && h < n; h++) {
        SrcStamp st;
        ok = stamp_of(hmm_paths[h], st) && wr(f, st.size) && wr(f, st.mtime_ns);
        const HostProfile &p = ps[h];
        ok = ok && wr(f, p.M) && wr(f, p.nseq) && wr(f, p.alph) && wr(f, p.stride) && wr(f, p.Q);
        for (auto fld : CACHE_FIELDS) ok = ok && wr_vec(f, p.*fld);
    }
    ok = (fclose(f) == 0) && ok;
    if (ok) ok = rename(tmp.c_str(), cache_path.c_str()) == 0;
    if (!ok) remove(tmp.c_str());
    return ok;
}

bool load_profile_cache(const std::string &cache_path, const std::vector<std::string> &hmm_paths, std::vector<HostProfile> &out) {
    out.clear();
    FILE *f = fopen(cache_path.c_str(), "rb");
    if (!f) return false;
    char magic[8];
    int n = 0;
    bool ok = fread(magic, 1, 8, f) == 8 && std::memcmp(magic, CACHE_MAGIC, 8) == 0 && rd(f, n) && n == (int)hmm_paths.size();
    std::vector<HostProfile> ps;
    for (int h = 0; ok && h < n; h++) {
        SrcStamp now, was;
        ok = stamp_of(hmm_paths[h], now) && rd(f, was.size) && rd(f, was.mtime_ns) && now.size == was.size && now.mtime_ns == was.mtime_ns;
        HostProfile p;
        ok = ok && rd(f, p.M) && rd(f, p.nseq) && rd(f, p.alph) && rd(f, p.stride) && rd(f, p.Q);
        ok = ok && p.M > 0 && p.M <= (1 << 20) && p.stride > p.M && p.Q >= 2 && p.alph >= 0 && p.alph <= 2;
        for (auto fld : CACHE_FIELDS) ok = ok && rd_vec(f, p.*fld, 64ull * (unsigned long long)(p.stride > 0 ? p.stride : 1));
        if (ok) {
            const int Kp = alphabet_info(p.alph).Kp;
            ok = p.tMM.size() == (size_t)p.stride && p.tDD.size() == (size_t)p.stride && p.entry.size() == (size_t)p.stride &&
                 p.gD.size() == (size_t)p.stride && p.emis.size() == (size_t)Kp * p.stride && p.otfv.size() == (size_t)32 * p.Q &&
                 p.orfv.size() == (size_t)Kp * p.Q * 4;
        }
        if (ok) ps.push_back(std::move(p));
    }
    fclose(f);
    if (!ok) return false;
    out = std::move(ps);
    return true;
}

}  // namespace witch
