// C ABI of libwitch_b200.so (see include/witch_b200.h). Host orchestration + kernel launches.
#include "../../include/witch_b200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>
#ifndef WITCH_HOST_SIM
#include <cub/cub.cuh>
#endif

#include "device_types.cuh"
#include "hmm_profile.h"
#include "parser_kernel.cuh"
#include "parser2_kernel.cuh"
#include "wave_kernels.cuh"
#include "md_kernel.cuh"
#include "worklist_kernels.cuh"
#include "post_kernels.cuh"
#include "graph_kernel.cuh"
#include "merge_kernel.cuh"

using namespace witch;

// ----------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};
static std::atomic<bool> g_prof{false};
struct ProfAcc { double ms = 0, cells = 0; uint64_t launches = 0; };
static ProfAcc g_acc[4];   // 0 parser, 1 envelope, 2 align, 3 multi-domain branch
// Timed launches leave a pair of CUDA events behind; they are resolved (cudaEventElapsedTime) when the numbers are read,
// so profiling adds no synchronisation to the stage calls.
struct ProfPending { cudaEvent_t a, b; int which; double cells; uint64_t launches; };
static std::vector<ProfPending> g_pending;
static std::mutex g_prof_mu;

static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CUDA_TRY(x)                                                                                          \
    do {                                                                                                     \
        cudaError_t e_ = (x);                                                                                \
        if (e_ != cudaSuccess)                                                                               \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #x);      \
    } while (0)

struct ScopedTimer {  // CUDA-event timing of a group of launches on `st` (only when profiling is enabled); never synchronises
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    int which;
    double cells;
    uint64_t n0;
    ScopedTimer(int which_, cudaStream_t st_, double cells_) : st(st_), which(which_), cells(cells_) {
        n0 = g_launches.load();
        if (g_prof.load()) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~ScopedTimer() {
        if (!a) return;
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_pending.push_back({a, b, which, cells, g_launches.load() - n0});
    }
};
static void prof_resolve() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &p : g_pending) {
        float ms = 0;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            g_acc[p.which].ms += ms; g_acc[p.which].cells += p.cells; g_acc[p.which].launches += p.launches;
        }
        cudaEventDestroy(p.a); cudaEventDestroy(p.b);
    }
    g_pending.clear();
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    void alloc(size_t n_) {
        if (n_ <= n && p) return;
        release();
        n = n_;
        CUDA_TRY(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    }
    void upload(const std::vector<T> &v, cudaStream_t st = nullptr) {
        alloc(v.size());
        if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

struct witch_ehmm {
    int device = 0, H = 0, alph = 0, Kp = 0, num_sms = 148;
    std::vector<int> M, nseq, stride;
    std::vector<long long> poff, eoff;
    std::vector<int> hrank;   // launch order of the HMMs: longer models first (stable)
    int maxQ = 0;
    DevBuf<float> tMM, tMI, tMD, tIM, tII, tDM, tDD, entry, gD, emis, otfv, orfv, ont8, onem;
    DevBuf<int> dM, dstride, dnseq, doQ, dhrank;
    DevBuf<long long> dpoff, deoff, dotoff, doroff, donoff;
    // reusable workspaces (owned by the handle: stage calls neither allocate nor free once they have grown)
    DevBuf<float> scratch, f1, f2;
    DevBuf<unsigned> counter;
    DevBuf<int> i1, i2, cntA, cntB, baseA, baseB, runhead, runstart, gflag, gid, group_first, grange;
    DevBuf<PairParse> parse;
    DevBuf<uint8_t> bytes, mdbytes, cubtmp;
    DevBuf<unsigned long long> keys, keys2;
    DevBuf<WaveItem> items, items2;
    DevBuf<MdRegion> mdregs;
    DevBuf<int> mdorder;
    DevBuf<long long> mdslot;
    DevBuf<unsigned> mdcounter;
    DevBuf<MdOut> mdout;
    DevBuf<WlDesc> desc;
    DevBuf<long long> coloff;
    WlDesc *hdesc = nullptr;   // pinned host copy of the work-list descriptor
    // side stream for the multi-domain branch (runs next to the envelope pass of the single-domain regions)
    // (two of them: the batches of a large region list alternate between two halves of the scratch, so that the tail of
    //  one batch's trace kernel -- its slowest walker -- overlaps with the next batch's Forward kernel)
    cudaStream_t aux = nullptr, aux2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_md = nullptr, ev_md2 = nullptr;
    void ensure_aux() {
        if (aux) return;
        CUDA_TRY(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&aux2, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ev_md, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&ev_md2, cudaEventDisableTiming));
    }
    ~witch_ehmm() {
        if (hdesc) cudaFreeHost(hdesc);
        if (aux) {
            cudaStreamDestroy(aux); cudaStreamDestroy(aux2);
            cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); cudaEventDestroy(ev_md); cudaEventDestroy(ev_md2);
        }
    }
    DevEhmm view() const {
        DevEhmm v;
        v.tMM = tMM.p; v.tMI = tMI.p; v.tMD = tMD.p; v.tIM = tIM.p; v.tII = tII.p; v.tDM = tDM.p; v.tDD = tDD.p;
        v.entry = entry.p; v.gD = gD.p; v.emis = emis.p; v.M = dM.p; v.stride = dstride.p; v.poff = dpoff.p; v.eoff = deoff.p;
        v.otfv = otfv.p; v.orfv = orfv.p; v.otoff = dotoff.p; v.oroff = doroff.p; v.oQ = doQ.p;
        v.ont8 = ont8.p; v.onem = onem.p; v.onoff = donoff.p;
        v.H = H; v.Kp = Kp;
        return v;
    }
};

// Device buffers of a query set come from a small per-device pool: a stage loop that creates and destroys one query set
// per batch then issues no cudaMalloc/cudaFree at all (cudaFree synchronises the device and was measured at up to 160 ms
// per call next to a polling nvidia-smi). Blocks are reused when they are at most 2x the request; at most 12 blocks / 1 GB
// are kept, the rest is freed as before.
struct QueryPool {
    struct Block { void *p; size_t bytes; int device; };
    std::mutex mu;
    std::vector<Block> free_blocks;
    size_t kept = 0;
    void *get(size_t bytes, int device, size_t *got) {
        {
            std::lock_guard<std::mutex> lk(mu);
            int best = -1;
            for (int i = 0; i < (int)free_blocks.size(); i++) {
                const Block &b = free_blocks[i];
                if (b.device == device && b.bytes >= bytes && b.bytes <= 2 * bytes + 4096 && (best < 0 || b.bytes < free_blocks[best].bytes)) best = i;
            }
            if (best >= 0) {
                Block b = free_blocks[best];
                free_blocks.erase(free_blocks.begin() + best);
                kept -= b.bytes;
                *got = b.bytes;
                return b.p;
            }
        }
        void *p = nullptr;
        CUDA_TRY(cudaMalloc(&p, std::max<size_t>(bytes, 256)));
        *got = std::max<size_t>(bytes, 256);
        return p;
    }
    void put(void *p, size_t bytes, int device) {
        if (!p) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            if (free_blocks.size() < 12 && kept + bytes <= ((size_t)1 << 30)) { free_blocks.push_back({p, bytes, device}); kept += bytes; return; }
        }
        cudaFree(p);
    }
};
static QueryPool g_qpool;
template <typename T>
struct PoolBuf {
    T *p = nullptr;
    size_t bytes = 0;
    int device = 0;
    void upload(const std::vector<T> &v, int dev) {
        device = dev;
        p = (T *)g_qpool.get(v.size() * sizeof(T), dev, &bytes);
        if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, nullptr));
    }
    ~PoolBuf() { g_qpool.put(p, bytes, device); }
};

struct witch_queries {
    int device = 0, n = 0, nsym = 0, alph = 0;
    std::vector<int> len;
    std::vector<long long> off;
    int symrow[MAX_SYM];
    int maxlen = 0;
    long long total = 0;
    PoolBuf<uint8_t> dsq;
    PoolBuf<long long> doff;
    PoolBuf<int> dlen;
    DevQueries view() const {
        DevQueries v;
        v.dsq = dsq.p; v.off = doff.p; v.len = dlen.p; v.n = n; v.nsym = nsym;
        for (int i = 0; i < MAX_SYM; i++) v.symrow[i] = symrow[i];
        return v;
    }
};

// ----------------------------------------------------------------------------------------------------------
extern "C" const char *witch_last_error(void) { return g_err.c_str(); }
extern "C" const char *witch_version(void) { return "witch_b200 0.1.0 sm_100a"; }
extern "C" int witch_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" uint64_t witch_kernel_launches(void) { return g_launches.load(); }
extern "C" void witch_prof_enable(int on) { g_prof.store(on != 0); }
extern "C" void witch_prof_reset(void) { prof_resolve(); for (auto &a : g_acc) a = ProfAcc(); }
extern "C" double witch_prof_get(int which, double *cells, uint64_t *launches) {
    if (which < 0 || which > 3) return 0;
    prof_resolve();
    if (cells) *cells = g_acc[which].cells;
    if (launches) *launches = g_acc[which].launches;
    return g_acc[which].ms;
}

static void require_device() {
    if (witch_device_count() <= 0) throw std::runtime_error("no CUDA device available (witch_b200 has no CPU fallback)");
}

// Upload configured profiles to the current device (shared by the text path and the profile-cache path).
static int ehmm_from_profiles(std::vector<HostProfile> &ps, witch_ehmm **out) {
    const int n_hmm = (int)ps.size();
    witch_ehmm *e = nullptr;
    try {
        for (int h = 0; h < n_hmm; h++)
            if (ps[h].alph != ps[0].alph) return fail(WITCH_ERR_ARG, "profiles use different alphabets");
        require_device();
        e = new witch_ehmm();
        CUDA_TRY(cudaGetDevice(&e->device));
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, e->device));
        e->num_sms = prop.multiProcessorCount;
        e->H = n_hmm; e->alph = ps[0].alph; e->Kp = alphabet_info(e->alph).Kp;
        long long po = 0, eo = 0;
        for (auto &p : ps) {
            e->M.push_back(p.M); e->nseq.push_back(p.nseq); e->stride.push_back(p.stride);
            e->poff.push_back(po); e->eoff.push_back(eo);
            po += p.stride; eo += (long long)e->Kp * p.stride;
        }
        auto cat = [&](std::vector<float> HostProfile::*f) {
            std::vector<float> v; v.reserve(po);
            for (auto &p : ps) v.insert(v.end(), (p.*f).begin(), (p.*f).end());
            return v;
        };
        e->tMM.upload(cat(&HostProfile::tMM)); e->tMI.upload(cat(&HostProfile::tMI));
        e->tMD.upload(cat(&HostProfile::tMD)); e->tIM.upload(cat(&HostProfile::tIM));
        e->tII.upload(cat(&HostProfile::tII)); e->tDM.upload(cat(&HostProfile::tDM));
        e->tDD.upload(cat(&HostProfile::tDD)); e->entry.upload(cat(&HostProfile::entry));
        e->gD.upload(cat(&HostProfile::gD));
        e->emis.upload(cat(&HostProfile::emis));
        e->dM.upload(e->M); e->dstride.upload(e->stride); e->dnseq.upload(e->nseq);
        e->dpoff.upload(e->poff); e->deoff.upload(e->eoff);
        {   // hmmsearch's striped float tables (multi-domain branch) and the launch rank of every HMM
            std::vector<long long> otoff, oroff, onoff;
            std::vector<int> oQ;
            std::vector<float> tf, rf, nt8, nem;
            const int KE = alphabet_info(e->alph).K;
            for (auto &p : ps) {
                otoff.push_back((long long)tf.size()); oroff.push_back((long long)rf.size()); oQ.push_back(p.Q);
                tf.insert(tf.end(), p.otfv.begin(), p.otfv.end());
                rf.insert(rf.end(), p.orfv.begin(), p.orfv.end());
                e->maxQ = std::max(e->maxQ, p.Q);
                // node-indexed copies for the trace walk: node k sits at striped position q = (k-1) % Q, z = (k-1) / Q
                const long long n0 = (long long)nt8.size() / 8;
                onoff.push_back(n0);
                const int Q = p.Q, NN = 4 * Q + 1;
                nt8.resize((size_t)(n0 + NN) * 8, 0.f);
                nem.resize((size_t)(n0 + NN) * KE, 0.f);
                for (int k = 1; k < NN; k++) {
                    const int q = (k - 1) % Q, z = (k - 1) / Q;
                    float *t8 = &nt8[(size_t)(n0 + k) * 8];
                    for (int tt = 0; tt < 4; tt++) t8[tt] = p.otfv[((size_t)7 * q + tt) * 4 + z];          // BM MM IM DM
                    t8[4] = p.otfv[((size_t)7 * q + 4) * 4 + z];                                           // MD
                    t8[5] = p.otfv[((size_t)7 * Q + q) * 4 + z];                                           // DD
                    t8[6] = p.otfv[((size_t)7 * q + 5) * 4 + z]; t8[7] = p.otfv[((size_t)7 * q + 6) * 4 + z];   // MI II
                    for (int x = 0; x < KE; x++) nem[(size_t)(n0 + k) * KE + x] = p.orfv[((size_t)x * Q + q) * 4 + z];
                }
            }
            e->otfv.upload(tf); e->orfv.upload(rf); e->dotoff.upload(otoff); e->doroff.upload(oroff); e->doQ.upload(oQ);
            e->ont8.upload(nt8); e->onem.upload(nem); e->donoff.upload(onoff);
            std::vector<int> horder(e->H);
            e->hrank.assign(e->H, 0);
            std::iota(horder.begin(), horder.end(), 0);
            std::stable_sort(horder.begin(), horder.end(), [&](int a, int b) { return e->M[a] > e->M[b]; });
            for (int r = 0; r < e->H; r++) e->hrank[horder[r]] = r;
            e->dhrank.upload(e->hrank);
            CUDA_TRY(cudaMallocHost(&e->hdesc, sizeof(WlDesc)));
        }
        CUDA_TRY(cudaDeviceSynchronize());
        *out = e;
        return WITCH_OK;
    } catch (const std::exception &ex) {
        delete e;
        return fail(WITCH_ERR_CUDA, ex.what());
    }
}
extern "C" int witch_ehmm_create(int n_hmm, const char *const *paths, witch_ehmm **out) {
    if (n_hmm <= 0 || !paths || !out) return fail(WITCH_ERR_ARG, "witch_ehmm_create: bad arguments");
    std::vector<HostProfile> ps;
    ps.reserve(n_hmm);
    for (int h = 0; h < n_hmm; h++) {
        if (!paths[h]) return fail(WITCH_ERR_ARG, "witch_ehmm_create: null path");
        try { ps.push_back(load_profile(paths[h], 0)); }
        catch (const std::exception &ex) { return fail(WITCH_ERR_IO, ex.what()); }
    }
    return ehmm_from_profiles(ps, out);
}

extern "C" int witch_ehmm_create_cached(int n_hmm, const char *const *paths, const char *cache_path, int *cache_hit, witch_ehmm **out) {
    if (n_hmm <= 0 || !paths || !out || !cache_path) return fail(WITCH_ERR_ARG, "witch_ehmm_create_cached: bad arguments");
    if (cache_hit) *cache_hit = 0;
    std::vector<std::string> pv;
    for (int h = 0; h < n_hmm; h++) {
        if (!paths[h]) return fail(WITCH_ERR_ARG, "witch_ehmm_create_cached: null path");
        pv.push_back(paths[h]);
    }
    std::vector<HostProfile> ps;
    if (load_profile_cache(cache_path, pv, ps)) {
        if (cache_hit) *cache_hit = 1;
    } else {
        ps.reserve(n_hmm);
        for (int h = 0; h < n_hmm; h++) {
            try { ps.push_back(load_profile(paths[h], 0)); }
            catch (const std::exception &ex) { return fail(WITCH_ERR_IO, ex.what()); }
        }
        save_profile_cache(cache_path, pv, ps);   // (best effort: an unwritable directory only costs the next run its parse)
    }
    return ehmm_from_profiles(ps, out);
}
extern "C" void witch_ehmm_destroy(witch_ehmm *e) {
    if (!e) return;
    cudaSetDevice(e->device);   // the buffers are freed on the device that owns them
    delete e;
}
extern "C" int witch_ehmm_count(const witch_ehmm *e) { return e ? e->H : 0; }
extern "C" int witch_ehmm_alphabet(const witch_ehmm *e) { return e ? e->alph : -1; }
extern "C" int witch_ehmm_info(const witch_ehmm *e, int32_t *M, int32_t *nseq) {
    if (!e) return fail(WITCH_ERR_ARG, "null handle");
    for (int h = 0; h < e->H; h++) { if (M) M[h] = e->M[h]; if (nseq) nseq[h] = e->nseq[h]; }
    return WITCH_OK;
}

extern "C" int witch_queries_create(const witch_ehmm *e, int n, const char *residues, const int64_t *offsets,
                                    witch_queries **out) {
    if (!e || n < 0 || !offsets || !out || (n > 0 && !residues)) return fail(WITCH_ERR_ARG, "witch_queries_create: bad arguments");
    witch_queries *q = nullptr;
    try {
        const AlphabetInfo &A = alphabet_info(e->alph);
        if (offsets[0] != 0) return fail(WITCH_ERR_ARG, "witch_queries_create: offsets[0] must be 0");
        for (int i = 0; i < n; i++)
            if (offsets[i + 1] < offsets[i]) return fail(WITCH_ERR_ARG, "witch_queries_create: offsets not monotone");
        require_device();
        CUDA_TRY(cudaSetDevice(e->device));   // the query set lives on the eHMM's device
        q = new witch_queries();
        q->n = n; q->alph = e->alph; q->device = e->device;
        const long long total = offsets[n];
        std::vector<uint8_t> codes((size_t)total);
        int dense[64];
        for (int i = 0; i < 64; i++) dense[i] = -1;
        q->nsym = 0;
        for (int x = 0; x < A.K; x++) { dense[x] = q->nsym; q->symrow[q->nsym++] = x; }  // canonical rows first
        for (int i = 0; i < n; i++) {
            if (offsets[i + 1] < offsets[i]) { delete q; return fail(WITCH_ERR_ARG, "offsets not monotone"); }
            q->len.push_back((int)(offsets[i + 1] - offsets[i]));
            q->off.push_back(offsets[i]);
            q->maxlen = std::max(q->maxlen, q->len.back());
        }
        q->off.push_back(total);
        q->total = total;
        for (long long p = 0; p < total; p++) {
            int c = A.code[(unsigned char)residues[p]];
            if (c < 0 || c == A.K || c >= A.Kp - 2) {  // invalid char, gap, '*' or '~' inside an unaligned query
                delete q;
                return fail(WITCH_ERR_ARG, std::string("invalid residue '") + residues[p] + "' in query");
            }
            if (dense[c] < 0) {
                if (q->nsym >= MAX_SYM) { delete q; return fail(WITCH_ERR_LIMIT, "too many distinct symbols"); }
                dense[c] = q->nsym; q->symrow[q->nsym++] = c;
            }
            codes[(size_t)p] = (uint8_t)dense[c];
        }
        for (int i = q->nsym; i < MAX_SYM; i++) q->symrow[i] = 0;
        require_device();
        q->dsq.upload(codes, q->device); q->doff.upload(q->off, q->device); q->dlen.upload(q->len, q->device);
        CUDA_TRY(cudaDeviceSynchronize());
        *out = q;
        return WITCH_OK;
    } catch (const std::exception &ex) {
        delete q;
        return fail(WITCH_ERR_CUDA, ex.what());
    }
}
extern "C" void witch_queries_destroy(witch_queries *q) {
    if (!q) return;
    cudaSetDevice(q->device);
    delete q;
}
extern "C" int witch_queries_count(const witch_queries *q) { return q ? q->n : 0; }

// ----------------------------------------------------------------------------------------------------------
// Family S launch plumbing
struct SClass { int C, T; std::vector<int> hmms; };

static int parser_generation();
// shared memory of the generation-6 parser classes (128 threads x C columns: emission rows + the parameter rows kept in smem)
static size_t parser6_smem(int nsym, int C, int pmode) { return ((size_t)(nsym + p2_smem_rows(pmode)) * 128 * C + PARSER_RED_ROWS * S_RED) * sizeof(float); }
static std::vector<SClass> s_classes(const witch_ehmm *e, const std::vector<int> &hmms, int nsym, bool allow6 = true) {
    std::map<std::pair<int, int>, std::vector<int>> m;
    // generation 6 needs two resident CTAs per SM to pay: 2 x (tables + 1 KB reserved) <= 227 KB. C = 13 (hybrid: 3 parameter
    // rows in smem): up to 13 distinct query symbols; C = 16 (all 9 parameter rows in smem): the 4 canonical ones only
    const bool g6 = allow6 && parser_generation() == 6;
    const bool gen6_13 = g6 && 2 * (parser6_smem(nsym, 13, 2) + 1024) <= 227 * 1024;
    const bool gen6_16 = g6 && 2 * (parser6_smem(nsym, 16, 1) + 1024) <= 227 * 1024;
    for (int h : hmms) {
        const int M = e->M[h];
        if (gen6_13 && M > 1024 && M <= 13 * 128) { m[{13, 128}].push_back(h); continue; }
        // models of 2,049-3,328 nodes (the root of a 16S-sized decomposition): the same kernel with 256 threads, one CTA per SM
        if (g6 && parser6_smem(nsym, 13, 2) * 2 + 1024 <= 227 * 1024 && M > 2048 && M <= 13 * 256) { m[{13, 256}].push_back(h); continue; }
        if (gen6_16 && M > 13 * 128 && M <= 2048) { m[{16, 128}].push_back(h); continue; }
        if (M > 8192) throw std::runtime_error("model longer than 8192 nodes is not supported");   // (check_limits refuses earlier)
        // C = 16 (3,841 .. 8,192 nodes, e.g. the root of a 16S-sized decomposition): the parameter set no longer fits the
        // register file and spills to local memory -- a slow class for the few models that long, not a fast path
        int C = (M <= 1024) ? 4 : (M <= 3072) ? 8 : (M <= 3840) ? 12 : 16;
        int T = ((M + C - 1) / C + 31) / 32 * 32;
        if (T < 64) T = 64;
        m[{C, T}].push_back(h);
    }
    std::vector<SClass> out;
    for (auto &kv : m) out.push_back({kv.first.first, kv.first.second, kv.second});
    return out;
}

template <int C, int MAXT, int MINB>
static void launch_parser(const witch_ehmm *e, const witch_queries *q, int T, ParserWork wk, cudaStream_t st, int maxgrid) {
    const size_t smem = ((size_t)q->nsym * T * C + PARSER_RED_ROWS * S_RED) * sizeof(float);
    if (smem > 200 * 1024) throw std::runtime_error("emission table does not fit shared memory (too many symbols x model length)");
    auto kern = mh_parser_kernel<C, MAXT, MINB>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    if (occ < 1) throw std::runtime_error("parser kernel cannot be resident (registers/shared memory)");
    const long long nitems = (long long)wk.nh * wk.nq;
    const int grid = (int)std::min<long long>(std::min<long long>(nitems, (long long)e->num_sms * occ), maxgrid);
    WITCH_LAUNCH(kern, grid, T, smem, st)(e->view(), q->view(), wk);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
}

template <int C, int MAXT, int MINB, int PMODE, bool FIXT = false>
static void launch_parser2(const witch_ehmm *e, const witch_queries *q, int T, ParserWork wk, cudaStream_t st, int maxgrid) {
    const size_t smem = ((size_t)(q->nsym + p2_smem_rows(PMODE)) * T * C + PARSER_RED_ROWS * S_RED) * sizeof(float);
    if (smem > 220 * 1024) throw std::runtime_error("emission + parameter tables do not fit shared memory");
    auto kern = mh_parser2_kernel<C, MAXT, MINB, PMODE, FIXT>;
    if (FIXT) T = MAXT;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    if (occ < 1) throw std::runtime_error("parser kernel cannot be resident (registers/shared memory)");
    const long long nitems = (long long)wk.nh * ((wk.nq + 1) / 2);
    const int grid = (int)std::min<long long>(std::min<long long>(nitems, (long long)e->num_sms * occ), maxgrid);
    WITCH_LAUNCH(kern, grid, T, smem, st)(e->view(), q->view(), wk);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
}

// Parser generation (WITCH_PARSER=n overrides the default):
//   1 = one query per CTA everywhere (parser_kernel.cuh; the round-1 kernel: 316-330 Gcell/s on truncated c2);
//   5 = two packed queries per CTA (parser2_kernel.cuh, parameters in registers) for the C = 8, T <= 224 class (models of
//       1,025-1,792 nodes) -- capped at 128 registers by its 14 warps/SM, it spills: 330 Gcell/s;
//   6 (default) = two packed queries per CTA in 128-thread CTAs that own MORE columns per thread, 2 CTAs/SM at 245
//       registers without spills (s_classes): 13 columns per thread with the six on-chain parameter sets in registers and
//       MI/II/entry in shared memory for models of 1,025-1,664 nodes (458 Gcell/s on truncated c2, 475 on full c2) and, with
//       256 threads, of 2,049-3,328 nodes; 16 columns per thread with all parameters in shared memory for 1,665-2,048
//       nodes (348 Gcell/s; plain-ACGT query sets only); the C = 4 classes (<= 1,024 nodes) run the register variant
//       with two packed queries (c4 sample: 248 vs 193 Gcell/s). Whatever does not fit falls back to 5, then 1.
// Measured and dropped (round 2): packed pairs for every C = 8 class with parameters in registers (305) or in shared
// memory (302), 12 columns x 160 threads (294: 168-register cap, spills), 9 columns x 192 threads (417), the local D chain
// with one dependent FFMA per link (no change), four partial sums for the row sums (no change, kept).
// Generations 1 and 5 give bit-identical results; generation 6 sums a row's columns in another association (13 or 16
// per thread instead of 8): scores differ by <= 2.5e-4 bits.
#ifndef WITCH_PARSER_DEFAULT
#define WITCH_PARSER_DEFAULT 6
#endif
static int parser_generation() {
    static const int g = [] { const char *s = getenv("WITCH_PARSER"); return s ? atoi(s) : WITCH_PARSER_DEFAULT; }();
    return g;
}

// Runs the multihit parser for all (query in qsel) x (hmm in hsel); results in e->parse [n_queries*H].
static void run_parser(witch_ehmm *e, witch_queries *q, const std::vector<int> &qsel, const std::vector<int> &hsel,
                       float *d_dbg_bwd, cudaStream_t st) {
    if (qsel.empty() || hsel.empty()) return;
    std::vector<int> qorder(qsel);
    std::stable_sort(qorder.begin(), qorder.end(), [&](int a, int b) { return q->len[a] > q->len[b]; });
    e->parse.alloc((size_t)q->n * e->H);
    e->i1.upload(qorder, st);
    const int Lcap = q->maxlen + 1;
    int maxgrid = e->num_sms * 8;
    e->scratch.alloc((size_t)maxgrid * 2 * PARSER_SCRATCH_ROWS * (size_t)((Lcap + 4) & ~3) + 64);   // (two queries per CTA in generation 2/3)
    e->counter.alloc(64);
    auto classes = s_classes(e, hsel, q->nsym, /*allow6=*/d_dbg_bwd == nullptr && qsel.size() >= 2);
    std::vector<int> allh;
    std::vector<int> hoff;
    for (auto &c : classes) {
        std::stable_sort(c.hmms.begin(), c.hmms.end(), [&](int a, int b) { return e->M[a] > e->M[b]; });
        hoff.push_back((int)allh.size());
        allh.insert(allh.end(), c.hmms.begin(), c.hmms.end());
    }
    e->i2.upload(allh, st);
    CUDA_TRY(cudaMemsetAsync(e->counter.p, 0, 64 * sizeof(unsigned), st));
    double cells = 0;
    {
        double sl = 0, sm = 0;
        for (int x : qsel) sl += q->len[x];
        for (int h : hsel) sm += e->M[h];
        cells = sl * sm;
    }
    ScopedTimer tm(0, st, cells);
    for (size_t ci = 0; ci < classes.size(); ci++) {
        if (ci >= 64) throw std::runtime_error("too many launch classes");
        ParserWork wk;
        wk.hmms = e->i2.p + hoff[ci]; wk.nh = (int)classes[ci].hmms.size();
        wk.qorder = e->i1.p; wk.nq = (int)qorder.size();
        wk.Lcap = Lcap; wk.scratch = e->scratch.p; wk.counter = e->counter.p + ci; wk.out = e->parse.p;
        wk.dbg_bwd = d_dbg_bwd;
        const int T = classes[ci].T;
        const int gen = d_dbg_bwd ? 1 : parser_generation();
        if (gen == 6 && classes[ci].C == 13 && T == 256 && wk.nq >= 2) { launch_parser2<13, 256, 1, 2, true>(e, q, T, wk, st, maxgrid); continue; }
        if (gen == 6 && classes[ci].C == 13 && T == 128 && wk.nq >= 2) { launch_parser2<13, 128, 2, 2, true>(e, q, T, wk, st, maxgrid); continue; }
        if (gen == 6 && classes[ci].C == 16 && T == 128 && wk.nq >= 2) { launch_parser2<16, 128, 2, 1, true>(e, q, T, wk, st, maxgrid); continue; }
        if (gen == 5 || gen == 6) {
            if (classes[ci].C == 8 && T <= 224 && wk.nq >= 2) { launch_parser2<8, 224, 2, 0>(e, q, T, wk, st, maxgrid); continue; }
            // models of at most 1,024 nodes (protein families, short markers): two packed queries per CTA, parameters in
            // registers (128 registers, no spills; c4 sample: 248 vs 193 Gcell/s, bit-identical)
            if (gen == 6 && classes[ci].C == 4 && T <= 256 && wk.nq >= 2) { launch_parser2<4, 256, 2, 0>(e, q, T, wk, st, maxgrid); continue; }
        }
        switch (classes[ci].C) {
            case 4: if (T <= 256) launch_parser<4, 256, 2>(e, q, T, wk, st, maxgrid); else launch_parser<4, 512, 1>(e, q, T, wk, st, maxgrid); break;
            case 8: if (T <= 256) launch_parser<8, 256, 2>(e, q, T, wk, st, maxgrid); else launch_parser<8, 384, 1>(e, q, T, wk, st, maxgrid); break;
            case 12: launch_parser<12, 320, 1>(e, q, T, wk, st, maxgrid); break;
            case 16: launch_parser<16, 512, 1>(e, q, T, wk, st, maxgrid); break;
            default: throw std::runtime_error("bad class");
        }
    }
}

#include "abi_stages.inl"
