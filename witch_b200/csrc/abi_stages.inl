// Stage orchestration behind the C ABI (included by witch_abi.cu).
#include <chrono>
struct HostClock {   // WITCH_TIMING=1: wall-clock of the host-side phases of a call, to stderr
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    bool on = getenv("WITCH_TIMING") != nullptr;
    void lap(const char *what) {
        if (!on) return;
        cudaDeviceSynchronize();
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[witch timing] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Wavefront launches. One launch = one bucket (length class x model-size class) of a SORTED device item list; the group
// list and the {first, end} group pair of the bucket live on the device (built by worklist_kernels.cuh for the score
// stage, uploaded from the host for the align stage), so a sequence of launches needs no synchronisation in between.
struct WaveLaunch {
    int item_end = 0;        // items of the bucket end here
    int Lcap = 0, maxM = 0;  // longest envelope / model of the bucket
    long long nitems = 0;
    double cells = 0;
    const int *grange = nullptr;  // device {first group, end group}
    unsigned *counter = nullptr;  // device work counter of this launch (zeroed by the caller)
};

// envelope-mode launch shape: warps per CTA x resident CTAs per SM. 4 x 3 (12 warps/SM, 168 registers) is the measured
// optimum: 6 x 2 -3 %, 2 x 6 -35 % (six emission tables per SM), 8 x 2 (128 registers, spills) -38 % -- DESIGN.md section 8
#ifndef WITCH_ENV_WARPS
#define WITCH_ENV_WARPS 4
#endif
#ifndef WITCH_ENV_MINB
#define WITCH_ENV_MINB 3
#endif
template <bool ALIGN, int C, int WAVE_WARPS, int MINB, int RING, bool LANE_EXP>
static void run_wave_c(witch_ehmm *e, witch_queries *q, const WaveItem *d_items, const int *d_group_first,
                       const std::vector<WaveLaunch> &launches, float *d_envsc, float *d_domcorr, int *d_cols,
                       const long long *d_coloff, float *d_dbg_fwd, float *d_dbg_bwd, cudaStream_t st) {
    const int SW = 32 * C;
    constexpr bool ROW16 = W_ROW16 != 0 && !ALIGN && !LANE_EXP;
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const double budget = std::min<double>(64.0e9, 0.5 * (double)(free_b + e->bytes.n));
    auto kern = wave_kernel<C, ALIGN, WAVE_WARPS, MINB, RING, LANE_EXP>;
    struct Plan { WaveLayout lay; int max_strips, emis_floats, res_cap; size_t smem; long long grid; };
    std::vector<Plan> plans;
    size_t need = 0;
    for (auto &L : launches) {   // plan every launch first: the scratch must not be reallocated while launches are in flight
        Plan P;
        P.max_strips = (L.maxM + SW - 1) / SW;
        P.lay = wave_layout(L.Lcap, P.max_strips, C, ALIGN, LANE_EXP);
        P.emis_floats = q->nsym * P.max_strips * SW;
        P.res_cap = (L.Lcap + 1 + 15) / 16 * 16;
        P.smem = (size_t)P.emis_floats * sizeof(float) + (size_t)WAVE_WARPS * P.res_cap +
                 (size_t)WAVE_WARPS * RING * (wave_ring_stage_bytes(C, ALIGN, ROW16) + 8) + (size_t)WAVE_WARPS * wave_bnd_ring_bytes();
        if (P.smem > 220 * 1024) throw std::runtime_error("emission table + residue staging do not fit shared memory");
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
        int occ = 1;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WAVE_WARPS * 32, P.smem));
        if (occ < 1) throw std::runtime_error("wave kernel cannot be resident");
        // (every item may be its own group in the worst case; the kernel reads the true group range from the device)
        long long grid = std::min<long long>(L.nitems, (long long)e->num_sms * occ);
        const long long max_slots = (long long)(budget / (double)P.lay.total);
        if (max_slots < WAVE_WARPS) throw std::runtime_error("not enough device memory for the wave scratch");
        P.grid = std::max<long long>(1, std::min<long long>(grid, max_slots / WAVE_WARPS));
        need = std::max<size_t>(need, (size_t)((long long)P.grid * WAVE_WARPS * P.lay.total));
        plans.push_back(P);
    }
    if (need > e->bytes.n) { CUDA_TRY(cudaStreamSynchronize(st)); e->bytes.alloc(need); }   // grows during warm-up only
    for (size_t z = 0; z < launches.size(); z++) {
        const WaveLaunch &L = launches[z];
        const Plan &P = plans[z];
        WaveWork wk;
        wk.items = d_items; wk.group_first = d_group_first; wk.grange = L.grange; wk.item_end = L.item_end;
        wk.counter = L.counter; wk.scratch = (char *)e->bytes.p; wk.slot_bytes = P.lay.total; wk.Lcap = L.Lcap;
        wk.max_strips = P.max_strips; wk.emis_floats = P.emis_floats; wk.res_cap = P.res_cap; wk.envsc = d_envsc; wk.domcorr = d_domcorr;
        wk.cols = d_cols; wk.col_off = d_coloff; wk.dbg_fwd = d_dbg_fwd; wk.dbg_bwd = d_dbg_bwd;
        ScopedTimer tm(ALIGN ? 2 : 1, st, L.cells);
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
        WITCH_LAUNCH(kern, (int)P.grid, WAVE_WARPS * 32, P.smem, st)(e->view(), q->view(), wk);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
}

template <bool ALIGN>
static void run_wave(witch_ehmm *e, witch_queries *q, const WaveItem *d_items, const int *d_group_first,
                     const std::vector<WaveLaunch> &launches, float *d_envsc, float *d_domcorr, int *d_cols,
                     const long long *d_coloff, float *d_dbg_fwd, float *d_dbg_bwd, cudaStream_t st) {
    if (launches.empty()) return;
#define WV_ARGS e, q, d_items, d_group_first, launches, d_envsc, d_domcorr, d_cols, d_coloff, d_dbg_fwd, d_dbg_bwd, st
    const bool lane_exp = e->alph == ALPH_AMINO;  // per-lane scaling exponents (see wave_kernels.cuh)
    if (ALIGN) { if (lane_exp) run_wave_c<true, 8, 4, 2, 3, true>(WV_ARGS); else run_wave_c<true, 8, 4, 2, 3, false>(WV_ARGS); }
    else { if (lane_exp) run_wave_c<false, 8, WITCH_ENV_WARPS, WITCH_ENV_MINB, 3, true>(WV_ARGS); else run_wave_c<false, 8, WITCH_ENV_WARPS, WITCH_ENV_MINB, 3, false>(WV_ARGS); }
#undef WV_ARGS
}

// Host-built work list (align stage, debug hooks: the pairs come from the caller as host arrays): bucket by length class
// and model-size class, order each bucket by (model rank, longer first) with two stable counting passes, cut groups,
// upload everything into handle-owned buffers.
template <bool ALIGN>
static void run_wave_host_items(witch_ehmm *e, witch_queries *q, const std::vector<WaveItem> &in, float *d_envsc, float *d_domcorr,
                                int *d_cols, const long long *d_coloff, float *d_dbg_fwd, float *d_dbg_bwd, cudaStream_t st) {
    if (in.empty()) return;
    const int GW = 4;   // warps per CTA of every wave kernel instantiation used above
    std::vector<std::vector<WaveItem>> buckets(WL_BUCKETS);
    // align stage: a launch lasts as long as its longest item (one warp walks a whole (query, HMM) pair: ~65 ms for
    // 1,500 x 1,600 cells) and a step has about one wave of pairs, so the length classes up to 2,048 residues share ONE
    // launch (scratch is per resident warp, sized by the longest item either way): short pairs fill in next to the long ones
    auto bucket_of = [&](const WaveItem &it) {
        const int b = wl_bucket(it.Ls, e->M[it.h]);
        return (ALIGN && it.Ls <= 2048) ? 2 * 3 + (b & 1) : b;
    };
    for (auto &it : in) buckets[bucket_of(it)].push_back(it);
    std::vector<WaveItem> items;
    std::vector<int> gfirst, grange(2 * 16, 0);
    std::vector<WaveLaunch> launches;
    items.reserve(in.size());
    for (int b = 0; b < WL_BUCKETS; b++) {
        auto &bk = buckets[b];
        if (bk.empty()) continue;
        int Lcap = 0;
        for (auto &it : bk) Lcap = std::max(Lcap, it.Ls);
        std::vector<WaveItem> tmp(bk.size());
        std::vector<size_t> cnt((size_t)Lcap + 2, 0);
        for (auto &it : bk) cnt[Lcap - it.Ls + 1]++;                 // key 2: Ls descending
        for (size_t k = 1; k < cnt.size(); k++) cnt[k] += cnt[k - 1];
        for (auto &it : bk) tmp[cnt[Lcap - it.Ls]++] = it;
        cnt.assign((size_t)e->H + 1, 0);
        for (auto &it : tmp) cnt[e->hrank[it.h] + 1]++;              // key 1: model rank
        for (size_t k = 1; k < cnt.size(); k++) cnt[k] += cnt[k - 1];
        for (auto &it : tmp) bk[cnt[e->hrank[it.h]]++] = it;
        WaveLaunch L;
        L.Lcap = Lcap; L.nitems = (long long)bk.size();
        grange[2 * b] = (int)gfirst.size();
        const size_t base = items.size();
        for (size_t i = 0; i < bk.size();) {
            size_t j = i;
            while (j < bk.size() && bk[j].h == bk[i].h && j - i < (size_t)GW) j++;
            gfirst.push_back((int)(base + i));
            i = j;
        }
        grange[2 * b + 1] = (int)gfirst.size();
        for (auto &it : bk) { L.maxM = std::max(L.maxM, e->M[it.h]); L.cells += (double)it.Ls * e->M[it.h]; items.push_back(it); }
        L.item_end = (int)items.size();
        L.grange = (const int *)(intptr_t)b;   // bucket id for now: the device pointers exist after the upload below
        launches.push_back(L);
    }
    // (the previous call's launches may still read these buffers: reuse is ordered by the stream, growth synchronises)
    if (items.size() > e->items.n || gfirst.size() > e->group_first.n) CUDA_TRY(cudaStreamSynchronize(st));
    e->items.upload(items, st);
    e->group_first.upload(gfirst, st);
    e->grange.upload(grange, st);
    e->counter.alloc(64);
    CUDA_TRY(cudaMemsetAsync(e->counter.p, 0, 64 * sizeof(unsigned), st));
    for (auto &L : launches) {
        const int b = (int)(intptr_t)L.grange;
        L.grange = e->grange.p + 2 * b;
        L.counter = e->counter.p + b;
    }
    run_wave<ALIGN>(e, q, e->items.p, e->group_first.p, launches, d_envsc, d_domcorr, d_cols, d_coloff, d_dbg_fwd, d_dbg_bwd, st);
}

struct LimitError : std::runtime_error { using std::runtime_error::runtime_error; };

// Shared-memory limits, checked BEFORE any work of a stage call is enqueued: the kernels stage one emission row per
// symbol that occurs in the query set (always the K canonical ones) for the whole model, so symbols x model length is
// bounded: parser nsym*M*4 B <= ~200 KB, wave kernels nsym*ceil(M/256)*1 KB + residue staging <= ~220 KB.
// DNA/RNA: M <= 8192 with up to 6 distinct symbols; amino (20-25 symbols): M <= ~2,000.
static void check_limits(const witch_ehmm *e, const witch_queries *q) {
    for (int h = 0; h < e->H; h++) {
        const int M = e->M[h];
        if (M > 8192) throw LimitError("model longer than 8192 nodes is not supported");
        const int C = (M <= 1024) ? 4 : (M <= 3072) ? 8 : (M <= 3840) ? 12 : 16;
        int T = ((M + C - 1) / C + 31) / 32 * 32;
        if (T < 64) T = 64;
        const size_t ps = ((size_t)q->nsym * T * C + PARSER_RED_ROWS * S_RED) * sizeof(float);
        const size_t ws = (size_t)q->nsym * ((M + 255) / 256) * 256 * sizeof(float) + 4 * (size_t)((q->maxlen + 16) / 16 * 16) + 4 * 3 * 2056 + 4 * 800;
        if (ps > 200 * 1024 || ws > 220 * 1024)
            throw LimitError("model " + std::to_string(h) + " (" + std::to_string(M) + " nodes) x " + std::to_string(q->nsym) +
                             " distinct query symbols does not fit the kernels' shared-memory emission tables (limits: " +
                             "symbols x nodes <= ~50,000, i.e. 8192 nodes for plain DNA/RNA, ~2,000 nodes for protein)");
    }
}

static void check_handles(witch_ehmm *e, witch_queries *q) {
    if (!e || !q) throw std::invalid_argument("null handle");
    if (e->alph != q->alph) throw std::invalid_argument("queries were digitised for another alphabet");
    if (e->device != q->device) throw std::invalid_argument("queries live on another device than the eHMM");
    CUDA_TRY(cudaSetDevice(e->device));
    check_limits(e, q);
}

// exclusive sum / key-value sort / running maximum on the device (CUB; the host simulation build substitutes loops)
static void dev_exclusive_sum(witch_ehmm *e, const int *in, int *out, long long n, cudaStream_t st) {
#ifdef WITCH_HOST_SIM
    int acc = 0;
    for (long long i = 0; i < n; i++) { const int v = in[i]; out[i] = acc; acc += v; }
#else
    size_t tmp = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, st));
    e->cubtmp.alloc(tmp);
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(e->cubtmp.p, tmp, in, out, n, st));
#endif
}
static void dev_inclusive_max(witch_ehmm *e, const int *in, int *out, long long n, cudaStream_t st) {
#ifdef WITCH_HOST_SIM
    int acc = 0;
    for (long long i = 0; i < n; i++) { acc = std::max(acc, in[i]); out[i] = acc; }
#else
    size_t tmp = 0;
    CUDA_TRY(cub::DeviceScan::InclusiveScan(nullptr, tmp, in, out, cub::Max(), n, st));
    e->cubtmp.alloc(tmp);
    CUDA_TRY(cub::DeviceScan::InclusiveScan(e->cubtmp.p, tmp, in, out, cub::Max(), n, st));
#endif
}
static void dev_sort_items(witch_ehmm *e, const unsigned long long *kin, unsigned long long *kout, const WaveItem *vin, WaveItem *vout,
                           long long n, cudaStream_t st) {
#ifdef WITCH_HOST_SIM
    std::vector<long long> ix(n);
    std::iota(ix.begin(), ix.end(), 0);
    std::stable_sort(ix.begin(), ix.end(), [&](long long a, long long b) { return kin[a] < kin[b]; });
    for (long long i = 0; i < n; i++) { kout[i] = kin[ix[i]]; vout[i] = vin[ix[i]]; }
#else
    size_t tmp = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, n, 0, 36, st));
    e->cubtmp.alloc(tmp);
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(e->cubtmp.p, tmp, kin, kout, vin, vout, n, 0, 36, st));
#endif
}

// The multi-domain branch for the nmd flagged regions listed in e->mdregs -> e->mdout, enqueued on `st`.
// Three kernels per batch: the region Forward matrices (one warp per region), the stochastic traces (one THREAD per region:
// a region's 200 traces share one random-number stream and are sequential, so the parallelism is across regions), the
// clustering (one warp per region). Scratch slots are sized per region and packed; the host orders the regions by size
// and cuts batches that fit the scratch budget (it needs the region list for that: one read-back of 16 B per region).
static void run_md(witch_ehmm *e, witch_queries *q, int nmd, cudaStream_t st_list, cudaStream_t st) {
    if (nmd <= 0) return;
    std::vector<MdRegion> regs(nmd);
    CUDA_TRY(cudaMemcpyAsync(regs.data(), e->mdregs.p, (size_t)nmd * sizeof(MdRegion), cudaMemcpyDeviceToHost, st_list));
    CUDA_TRY(cudaStreamSynchronize(st_list));
    const int nsp_cap = 4096;
    std::vector<long long> bytes(nmd);
    std::vector<int> order(nmd);
    int Qcap = 2;
    for (int r = 0; r < nmd; r++) {
        const int h = regs[r].h, Lr = regs[r].j0 - regs[r].i0 + 1, Qh = std::max(2, (e->M[h] - 1) / 4 + 1);
        bytes[r] = md_layout(Lr, Qh, e->M[h], nsp_cap).total;
        Qcap = std::max(Qcap, Qh);
        order[r] = r;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bytes[a] > bytes[b]; });
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const long long budget_all = (long long)std::min<double>(110.0e9, 0.6 * (double)(free_b + e->mdbytes.n));
    if (bytes[order[0]] > budget_all) throw std::runtime_error("not enough device memory for the multi-domain scratch of one region");
    long long total_bytes = 0;
    for (int r = 0; r < nmd; r++) total_bytes += bytes[r];
    // a list that does not fit one batch is processed by two lanes (streams) on the two halves of the scratch
    const int nlanes = (total_bytes > budget_all && 2 * bytes[order[0]] <= budget_all) ? 2 : 1;
    const long long budget = (budget_all / nlanes) & ~255LL;
    std::vector<long long> slot_off(nmd);
    std::vector<int> batch_begin{0};
    long long used = 0, need = 0;
    for (int j = 0; j < nmd; j++) {
        const long long b = bytes[order[j]];
        if (used + b > budget) { batch_begin.push_back(j); used = 0; }
        slot_off[j] = used;
        used += b;
        need = std::max(need, used);
    }
    batch_begin.push_back(nmd);
    const int nbatch = (int)batch_begin.size() - 1;
    if (nlanes == 2) need = budget * 2;
    if ((size_t)need > e->mdbytes.n) { CUDA_TRY(cudaDeviceSynchronize()); e->mdbytes.alloc((size_t)need); }
    if ((size_t)nmd > e->mdorder.n || (size_t)2 * nbatch + 2 > e->mdcounter.n) { CUDA_TRY(cudaStreamSynchronize(st)); CUDA_TRY(cudaStreamSynchronize(e->aux2)); }
    e->mdorder.upload(order, st);
    e->mdslot.upload(slot_off, st);
    e->mdcounter.alloc((size_t)2 * nbatch + 2);
    CUDA_TRY(cudaMemsetAsync(e->mdcounter.p, 0, ((size_t)2 * nbatch + 2) * sizeof(unsigned), st));
    const size_t smem = (size_t)MD_WARPS * 8 * Qcap * sizeof(float);   // per warp: M and D vectors of the current row
    if (smem > 200 * 1024) throw std::runtime_error("model too long for the multi-domain branch's row staging");
    CUDA_TRY(cudaFuncSetAttribute(md_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ1 = 1, occ3 = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, md_forward_kernel, MD_WARPS * 32, smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, md_cluster_kernel, MD_WARPS * 32, 0));
    ScopedTimer tm(3, st, 0.0);
    if (nlanes == 2) {   // the second lane starts once the lists are uploaded
        CUDA_TRY(cudaEventRecord(e->ev_md, st));
        CUDA_TRY(cudaStreamWaitEvent(e->aux2, e->ev_md, 0));
    }
    for (int b = 0; b < nbatch; b++) {
        cudaStream_t sb = (nlanes == 2 && (b & 1)) ? e->aux2 : st;
        MdWork W;
        W.regions = e->mdregs.p; W.order = e->mdorder.p; W.slot_off = e->mdslot.p; W.begin = batch_begin[b]; W.end = batch_begin[b + 1];
        W.scratch = (char *)e->mdbytes.p + ((nlanes == 2 && (b & 1)) ? budget : 0); W.Qcap = Qcap; W.nsp_cap = nsp_cap; W.out = e->mdout.p; W.spread = 1;
        const int nb = W.end - W.begin;
        W.counter = e->mdcounter.p + 2 * b;
        int grid = (int)std::min<long long>((nb + MD_WARPS - 1) / MD_WARPS, (long long)e->num_sms * std::max(occ1, 1));
        WITCH_LAUNCH(md_forward_kernel, grid, MD_WARPS * 32, smem, sb)(e->view(), q->view(), W);
        {   // walkers per warp: as few as the batch size allows while the GPU still holds every walker at once
            static const int forced = [] { const char *s = getenv("WITCH_MD_SPREAD"); return s ? atoi(s) : 0; }();
            const long long cap = (long long)e->num_sms * 512;   // (measured: c4-sized batches want 32 walkers per warp, a c2 slab's 155 regions one each)
            int spread = 32;
            while (spread > 1 && (long long)nb * spread > cap) spread >>= 1;
            if (forced >= 1 && forced <= 32 && (forced & (forced - 1)) == 0) spread = forced;
            W.spread = spread;
            const unsigned tg = (unsigned)(((long long)nb * spread + 127) / 128);
            if (e->Kp == 29) WITCH_LAUNCH(md_trace_kernel<20>, tg, 128, 0, sb)(e->view(), q->view(), W);
            else WITCH_LAUNCH(md_trace_kernel<4>, tg, 128, 0, sb)(e->view(), q->view(), W);
        }
        W.counter = e->mdcounter.p + 2 * b + 1;
        grid = (int)std::min<long long>((nb + MD_WARPS - 1) / MD_WARPS, (long long)e->num_sms * std::max(occ3, 1));
        WITCH_LAUNCH(md_cluster_kernel, grid, MD_WARPS * 32, 0, sb)(e->view(), q->view(), W);
        g_launches += 3;
        CUDA_TRY(cudaGetLastError());
    }
    if (nlanes == 2) {   // join the second lane
        CUDA_TRY(cudaEventRecord(e->ev_md2, e->aux2));
        CUDA_TRY(cudaStreamWaitEvent(st, e->ev_md2, 0));
    }
}

extern "C" int witch_score_dev(witch_ehmm *e, witch_queries *q, float *d_scores, uint8_t *d_reported, float *d_pre,
                               uint8_t *d_flags, void *stream) {
    try {
        check_handles(e, q);
        if (!d_scores || !d_reported) return fail(WITCH_ERR_ARG, "witch_score: scores/reported must not be NULL");
        cudaStream_t st = (cudaStream_t)stream;
        const int nq = q->n, H = e->H;
        if (nq == 0) return WITCH_OK;
        const long long np = (long long)nq * H;
        if (np * (long long)MAX_ENV >= (1LL << 31)) return fail(WITCH_ERR_LIMIT, "more than 2^31 envelope slots: split the query set");
        std::vector<int> qs(nq), hs(H);
        std::iota(qs.begin(), qs.end(), 0);
        std::iota(hs.begin(), hs.end(), 0);
        HostClock hc;
        run_parser(e, q, qs, hs, nullptr, st);
        hc.lap("parser");
        // ---- region counts, scans, totals: the host learns the sizes of the two lists and the bucket shapes of the first ----
        const unsigned nb = (unsigned)((np + 255) / 256);
        e->cntA.alloc(np); e->cntB.alloc(np); e->baseA.alloc(np); e->baseB.alloc(np); e->desc.alloc(2);
        e->counter.alloc(64);
        CUDA_TRY(cudaMemsetAsync(e->counter.p, 0, 64 * sizeof(unsigned), st));
        CUDA_TRY(cudaMemsetAsync(e->desc.p, 0, 2 * sizeof(WlDesc), st));
        WITCH_LAUNCH(region_count_kernel, nb, 256, 0, st)(e->parse.p, np, H, e->dM.p, e->cntA.p, e->cntB.p, e->desc.p);
        dev_exclusive_sum(e, e->cntA.p, e->baseA.p, np, st);
        dev_exclusive_sum(e, e->cntB.p, e->baseB.p, np, st);
        WITCH_LAUNCH(scan_totals_kernel, 1, 32, 0, st)(e->cntA.p, e->baseA.p, e->cntB.p, e->baseB.p, np, e->desc.p);
        g_launches += 2;
        CUDA_TRY(cudaMemcpyAsync(e->hdesc, e->desc.p, sizeof(WlDesc), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));   // sizing step 1
        const int nA = e->hdesc->totals[0], nmd = e->hdesc->totals[1];
        const long long NB = (long long)nmd * MD_MAXC, NI = (long long)nA + NB, NL = std::max<long long>(nA, NB);
        hc.lap("region lists");
        e->f1.alloc((size_t)NI + 1); e->f2.alloc((size_t)NI + 1);
        e->items.alloc((size_t)NL + 1); e->items2.alloc((size_t)NL + 1); e->keys.alloc((size_t)NL + 1); e->keys2.alloc((size_t)NL + 1);
        e->runhead.alloc((size_t)NL + 1); e->runstart.alloc((size_t)NL + 1); e->gflag.alloc((size_t)NL + 1); e->gid.alloc((size_t)NL + 1);
        e->group_first.alloc((size_t)NL + 1); e->grange.alloc(64);
        e->mdregs.alloc((size_t)nmd + 1); e->mdout.alloc((size_t)nmd + 1);
        int maxM_small = 0, maxM_big = 0;
        for (int h = 0; h < H; h++) { if (e->M[h] > 13 * 256) maxM_big = std::max(maxM_big, e->M[h]); else maxM_small = std::max(maxM_small, e->M[h]); }
        // sorts the n items in e->items/e->keys, cuts groups, returns the launches described by `hd` (counters from cbase)
        auto build_and_launch = [&](long long n, const WlDesc *hd, int cbase, int *grange) {
            dev_sort_items(e, e->keys.p, e->keys2.p, e->items.p, e->items2.p, n, st);
            const unsigned nbi = (unsigned)((n + 255) / 256);
            CUDA_TRY(cudaMemsetAsync(grange, 0, 32 * sizeof(int), st));
            WITCH_LAUNCH(group_head_kernel, nbi, 256, 0, st)(e->keys2.p, (int)n, e->runhead.p);
            dev_inclusive_max(e, e->runhead.p, e->runstart.p, n, st);
            WITCH_LAUNCH(group_flag_kernel, nbi, 256, 0, st)(e->keys2.p, e->runstart.p, (int)n, WITCH_ENV_WARPS, e->gflag.p);
            dev_exclusive_sum(e, e->gflag.p, e->gid.p, n, st);
            WITCH_LAUNCH(group_scatter_kernel, nbi, 256, 0, st)(e->keys2.p, e->gflag.p, e->gid.p, (int)n, e->group_first.p, grange);
            g_launches += 3;
            std::vector<WaveLaunch> launches;
            long long off = 0;
            for (int b = 0; b < WL_BUCKETS; b++) {
                const int cnt = hd->count[b];
                if (cnt == 0) continue;
                WaveLaunch L;
                off += cnt;
                L.item_end = (int)off; L.nitems = cnt; L.Lcap = hd->maxLs[b]; L.maxM = (b & 1) ? maxM_big : maxM_small;
                L.cells = hd->cells[b]; L.grange = grange + 2 * b; L.counter = e->counter.p + cbase + b;
                launches.push_back(L);
            }
            run_wave<false>(e, q, e->items2.p, e->group_first.p, launches, e->f1.p, e->f2.p, nullptr, nullptr, nullptr, nullptr, st);
        };
        // ---- the multi-domain branch runs on the handle's side stream, next to the envelope pass of everything else ----
        if (nmd > 0) {
            e->ensure_aux();
            WITCH_LAUNCH(md_list_kernel, nb, 256, 0, st)(e->parse.p, np, H, e->baseB.p, e->mdregs.p);
            g_launches++;
            CUDA_TRY(cudaEventRecord(e->ev_fork, st));
            CUDA_TRY(cudaStreamWaitEvent(e->aux, e->ev_fork, 0));
            run_md(e, q, nmd, st, e->aux);   // (reads the region list back on `st`, then enqueues on the side stream)
            CUDA_TRY(cudaEventRecord(e->ev_join, e->aux));
        }
        if (nA > 0) {
            WITCH_LAUNCH(items_sd_kernel, nb, 256, 0, st)(e->parse.p, np, H, e->baseA.p, e->dhrank.p, e->dM.p, e->items.p, e->keys.p);
            g_launches++;
            const WlDesc hd1 = *e->hdesc;
            build_and_launch(nA, &hd1, 0, e->grange.p);
            hc.lap("envelope pass (single-domain regions) + multi-domain branch");
        }
        if (nmd > 0) {
            CUDA_TRY(cudaStreamWaitEvent(st, e->ev_join, 0));
            // (the first list's buffers are reused: its launches are complete once the host has the second descriptor)
            CUDA_TRY(cudaStreamSynchronize(st));
            if (hc.on) {   // WITCH_TIMING: where the branch spends its time (device clock ticks per region)
                std::vector<MdOut> mo(nmd);
                CUDA_TRY(cudaMemcpy(mo.data(), e->mdout.p, (size_t)nmd * sizeof(MdOut), cudaMemcpyDeviceToHost));
                double s0 = 0, s1 = 0, s2 = 0; long long m0 = 0, m1 = 0, m2 = 0, worst = 0; int wi = 0;
                for (int i = 0; i < nmd; i++) {
                    s0 += mo[i].clk[0]; s1 += mo[i].clk[1]; s2 += mo[i].clk[2];
                    m0 = std::max(m0, mo[i].clk[0]); m1 = std::max(m1, mo[i].clk[1]); m2 = std::max(m2, mo[i].clk[2]);
                    const long long tot = mo[i].clk[0] + mo[i].clk[1] + mo[i].clk[2];
                    if (tot > worst) { worst = tot; wi = i; }
                }
                fprintf(stderr, "[witch timing] md regions %d: Mticks mean fwd %.1f traces %.1f cluster %.1f | max fwd %.1f traces %.1f cluster %.1f | worst region Lr %d M %d: %.1f %.1f %.1f\n",
                        nmd, s0 / nmd / 1e6, s1 / nmd / 1e6, s2 / nmd / 1e6, m0 / 1e6, m1 / 1e6, m2 / 1e6, (int)(mo[wi].clk[3] >> 32), (int)(mo[wi].clk[3] & 0xffffffff),
                        mo[wi].clk[0] / 1e6, mo[wi].clk[1] / 1e6, mo[wi].clk[2] / 1e6);
            }
            WITCH_LAUNCH(items_md_kernel, (unsigned)((NB + 255) / 256), 256, 0, st)(e->mdregs.p, e->mdout.p, nmd, nA, e->dhrank.p, e->dM.p,
                                                                                e->items.p, e->keys.p, e->desc.p + 1);
            g_launches++;
            CUDA_TRY(cudaMemcpyAsync(e->hdesc, e->desc.p + 1, sizeof(WlDesc), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));   // sizing step 2 (only when some region took the multi-domain branch)
            const WlDesc hd2 = *e->hdesc;
            build_and_launch(NB, &hd2, 16, e->grange.p + 32);
            hc.lap("envelope pass (multi-domain envelopes)");
        }
        WITCH_LAUNCH(finalize_scores_kernel, (unsigned)((np + 255) / 256), 256, 0, st)(e->parse.p, q->dlen.p, nq, H, e->baseA.p, e->baseB.p,
                                                                           e->mdout.p, nA, e->f1.p, e->f2.p, d_scores, d_reported,
                                                                           d_pre, d_flags);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        return WITCH_OK;
    } catch (const LimitError &ex) {
        return fail(WITCH_ERR_LIMIT, ex.what());
    } catch (const std::invalid_argument &ex) {
        return fail(WITCH_ERR_ARG, ex.what());
    } catch (const std::exception &ex) {
        return fail(WITCH_ERR_CUDA, ex.what());
    }
}

extern "C" int witch_score(witch_ehmm *e, witch_queries *q, float *scores, uint8_t *reported, float *pre, uint8_t *flags) {
    try {
        check_handles(e, q);
        if (!scores || !reported) return fail(WITCH_ERR_ARG, "witch_score: scores/reported must not be NULL");
        const size_t np = (size_t)q->n * e->H;
        if (np == 0) return WITCH_OK;
        DevBuf<float> ds, dp;
        DevBuf<uint8_t> dr, df;
        ds.alloc(np); dp.alloc(np); dr.alloc(np); df.alloc(np);
        int rc = witch_score_dev(e, q, ds.p, dr.p, dp.p, df.p, nullptr);
        if (rc != WITCH_OK) return rc;
        CUDA_TRY(cudaMemcpy(scores, ds.p, np * sizeof(float), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(reported, dr.p, np, cudaMemcpyDeviceToHost));
        if (pre) CUDA_TRY(cudaMemcpy(pre, dp.p, np * sizeof(float), cudaMemcpyDeviceToHost));
        if (flags) CUDA_TRY(cudaMemcpy(flags, df.p, np, cudaMemcpyDeviceToHost));
        return WITCH_OK;
    } catch (const LimitError &ex) {
        return fail(WITCH_ERR_LIMIT, ex.what());
    } catch (const std::invalid_argument &ex) {
        return fail(WITCH_ERR_ARG, ex.what());
    } catch (const std::exception &ex) {
        return fail(WITCH_ERR_CUDA, ex.what());
    }
}

extern "C" int witch_weights_topk_dev(const witch_ehmm *e, const float *d_scores, const uint8_t *d_reported, int nq, int k,
                                      int round_decimals, int32_t *d_idx, double *d_w, int32_t *d_count, void *stream) {
    try {
        if (!e || !d_scores || !d_reported || !d_idx || !d_w || !d_count || k <= 0 || nq < 0)
            return fail(WITCH_ERR_ARG, "witch_weights_topk: bad arguments");
        CUDA_TRY(cudaSetDevice(e->device));
        if (nq == 0) return WITCH_OK;
        const int warps = 4;
        WITCH_LAUNCH(weights_topk_kernel, (nq + warps - 1) / warps, warps * 32, 0, (cudaStream_t)stream)(
            d_scores, d_reported, e->dnseq.p, nq, e->H, k, round_decimals, d_idx, d_w, d_count);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        return WITCH_OK;
    } catch (const std::exception &ex) {
        return fail(WITCH_ERR_CUDA, ex.what());
    }
}

extern "C" int witch_weights_topk(const witch_ehmm *e, const float *scores, const uint8_t *reported, int nq, int k,
                                  int round_decimals, int32_t *idx, double *w, int32_t *count) {
    try {
        if (!e || !scores || !reported || !idx || !w || !count || k <= 0 || nq < 0)
            return fail(WITCH_ERR_ARG, "witch_weights_topk: bad arguments");
        CUDA_TRY(cudaSetDevice(e->device));
        if (nq == 0) return WITCH_OK;
        const size_t np = (size_t)nq * e->H;
        DevBuf<float> ds; DevBuf<uint8_t> dr; DevBuf<int> di, dc; DevBuf<double> dw;
        ds.alloc(np); dr.alloc(np); di.alloc((size_t)nq * k); dw.alloc((size_t)nq * k); dc.alloc(nq);
        CUDA_TRY(cudaMemcpy(ds.p, scores, np * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(dr.p, reported, np, cudaMemcpyHostToDevice));
        int rc = witch_weights_topk_dev(e, ds.p, dr.p, nq, k, round_decimals, di.p, dw.p, dc.p, nullptr);
        if (rc != WITCH_OK) return rc;
        CUDA_TRY(cudaMemcpy(idx, di.p, (size_t)nq * k * sizeof(int), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(w, dw.p, (size_t)nq * k * sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(count, dc.p, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost));
        return WITCH_OK;
    } catch (const std::exception &ex) {
        return fail(WITCH_ERR_CUDA, ex.what());
    }
}

static std::vector<WaveItem> align_items(witch_ehmm *e, witch_queries *q, int n_pairs, const int32_t *qidx,
                                         const int32_t *hidx) {
    std::vector<WaveItem> items;
    items.reserve(n_pairs);
    for (int p = 0; p < n_pairs; p++) {
        if (qidx[p] < 0 || qidx[p] >= q->n || hidx[p] < 0 || hidx[p] >= e->H)
            throw std::invalid_argument("witch_align: pair index out of range");
        if (q->len[qidx[p]] <= 0) continue;
        WaveItem it;
        it.q = qidx[p]; it.h = hidx[p]; it.i0 = 1; it.Ls = q->len[qidx[p]]; it.pair = p;
        items.push_back(it);
    }
    return items;
}

extern "C" int witch_align_dev(witch_ehmm *e, witch_queries *q, int n_pairs, const int32_t *qidx, const int32_t *hidx,
                               const int64_t *col_offsets, int32_t *d_cols, void *stream) {
    try {
        check_handles(e, q);
        if (n_pairs < 0 || (n_pairs > 0 && (!qidx || !hidx || !col_offsets || !d_cols)))
            return fail(WITCH_ERR_ARG, "witch_align: bad arguments");
        if (n_pairs == 0) return WITCH_OK;
        cudaStream_t st = (cudaStream_t)stream;
        std::vector<WaveItem> items = align_items(e, q, n_pairs, qidx, hidx);
        std::vector<long long> co(n_pairs);
        for (int p = 0; p < n_pairs; p++) {
            if (col_offsets[p] < 0) return fail(WITCH_ERR_ARG, "witch_align: negative column offset");
            co[p] = col_offsets[p];
        }
        if ((size_t)n_pairs > e->coloff.n) CUDA_TRY(cudaStreamSynchronize(st));
        e->coloff.upload(co, st);
        run_wave_host_items<true>(e, q, items, nullptr, nullptr, d_cols, e->coloff.p, nullptr, nullptr, st);
        return WITCH_OK;
    } catch (const LimitError &ex) {
        return fail(WITCH_ERR_LIMIT, ex.what());
    } catch (const std::invalid_argument &ex) {
        return fail(WITCH_ERR_ARG, ex.what());
    } catch (const std::exception &ex) {
        return fail(WITCH_ERR_CUDA, ex.what());
    }
}

extern "C" int witch_align(witch_ehmm *e, witch_queries *q, int n_pairs, const int32_t *qidx, const int32_t *hidx,
                           const int64_t *col_offsets, int32_t *cols) {
    try {
        check_handles(e, q);
        if (n_pairs < 0 || (n_pairs > 0 && (!qidx || !hidx || !col_offsets || !cols)))
            return fail(WITCH_ERR_ARG, "witch_align: bad arguments");
        if (n_pairs == 0) return WITCH_OK;
        long long total = 0;
        for (int p = 0; p < n_pairs; p++) {
            if (qidx[p] < 0 || qidx[p] >= q->n) return fail(WITCH_ERR_ARG, "witch_align: pair index out of range");
            total = std::max<long long>(total, col_offsets[p] + q->len[qidx[p]]);
        }
        DevBuf<int> dc; dc.alloc((size_t)total);
        CUDA_TRY(cudaMemset(dc.p, 0xff, (size_t)total * sizeof(int)));
        int rc = witch_align_dev(e, q, n_pairs, qidx, hidx, col_offsets, dc.p, nullptr);
        if (rc != WITCH_OK) return rc;
        CUDA_TRY(cudaMemcpy(cols, dc.p, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost));
        return WITCH_OK;
    } catch (const LimitError &ex) {
        return fail(WITCH_ERR_LIMIT, ex.what());
    } catch (const std::invalid_argument &ex) {
        return fail(WITCH_ERR_ARG, ex.what());
    } catch (const std::exception &ex) {
        return fail(WITCH_ERR_CUDA, ex.what());
    }
}

extern "C" int witch_debug_fwdbwd(witch_ehmm *e, witch_queries *q, int n_pairs, const int32_t *qidx, const int32_t *hidx,
                                  int mode, float *fwd_nats, float *bwd_nats) {
    try {
        check_handles(e, q);
        if (n_pairs <= 0 || !qidx || !hidx || !fwd_nats || !bwd_nats) return fail(WITCH_ERR_ARG, "bad arguments");
        for (int p = 0; p < n_pairs; p++)
            if (qidx[p] < 0 || qidx[p] >= q->n || hidx[p] < 0 || hidx[p] >= e->H) return fail(WITCH_ERR_ARG, "witch_debug_fwdbwd: pair index out of range");
        if (mode == 1) {
            // multihit parser: run per distinct (q, h) through Family S
            std::vector<int> qs(qidx, qidx + n_pairs), hs(hidx, hidx + n_pairs);
            std::sort(qs.begin(), qs.end()); qs.erase(std::unique(qs.begin(), qs.end()), qs.end());
            std::sort(hs.begin(), hs.end()); hs.erase(std::unique(hs.begin(), hs.end()), hs.end());
            DevBuf<float> db; db.alloc((size_t)q->n * e->H);
            run_parser(e, q, qs, hs, db.p, nullptr);
            std::vector<PairParse> parse((size_t)q->n * e->H);
            std::vector<float> bw((size_t)q->n * e->H);
            CUDA_TRY(cudaMemcpy(parse.data(), e->parse.p, parse.size() * sizeof(PairParse), cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(bw.data(), db.p, bw.size() * sizeof(float), cudaMemcpyDeviceToHost));
            for (int p = 0; p < n_pairs; p++) {
                const size_t ix = (size_t)qidx[p] * e->H + hidx[p];
                fwd_nats[p] = parse[ix].fwd_bits * 0.69314718056f;
                bwd_nats[p] = bw[ix];
            }
        } else {
            std::vector<WaveItem> items = align_items(e, q, n_pairs, qidx, hidx);
            DevBuf<float> df, db, d1, d2;
            df.alloc(n_pairs); db.alloc(n_pairs); d1.alloc(n_pairs); d2.alloc(n_pairs);
            CUDA_TRY(cudaMemset(df.p, 0, n_pairs * sizeof(float)));
            CUDA_TRY(cudaMemset(db.p, 0, n_pairs * sizeof(float)));
            run_wave_host_items<false>(e, q, items, d1.p, d2.p, nullptr, nullptr, df.p, db.p, nullptr);
            CUDA_TRY(cudaDeviceSynchronize());
            CUDA_TRY(cudaMemcpy(fwd_nats, df.p, n_pairs * sizeof(float), cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(bwd_nats, db.p, n_pairs * sizeof(float), cudaMemcpyDeviceToHost));
        }
        return WITCH_OK;
    } catch (const LimitError &ex) {
        return fail(WITCH_ERR_LIMIT, ex.what());
    } catch (const std::invalid_argument &ex) {
        return fail(WITCH_ERR_ARG, ex.what());
    } catch (const std::exception &ex) {
        return fail(WITCH_ERR_CUDA, ex.what());
    }
}

extern "C" double witch_measure_fp32_peak(double ms_target) {
    try {
        require_device();
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
        DevBuf<float> out; out.alloc(4);
        const int grid = prop.multiProcessorCount * 8, block = 256;
        cudaEvent_t a, b;
        CUDA_TRY(cudaEventCreate(&a)); CUDA_TRY(cudaEventCreate(&b));
        int iters = 1 << 14;
        double best = 0;
        for (int rep = 0; rep < 6; rep++) {
            CUDA_TRY(cudaEventRecord(a));
            WITCH_LAUNCH(ffma_peak_kernel, grid, block)(out.p, iters);
            g_launches++;
            CUDA_TRY(cudaEventRecord(b));
            CUDA_TRY(cudaEventSynchronize(b));
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
            const double tf = 2.0 * 8.0 * (double)iters * grid * block / (ms * 1e-3) / 1e12;
            if (rep > 0) best = std::max(best, tf);
            if (ms < ms_target && iters < (1 << 24)) iters *= 2;
        }
        cudaEventDestroy(a); cudaEventDestroy(b);
        return best;
    } catch (const std::exception &ex) {
        g_err = ex.what();
        return -1.0;
    }
}

extern "C" int witch_graph_align(witch_ehmm *e, int nq, const int32_t *qlen, const int64_t *res_off, const char *residues,
                                 const int32_t *pair_begin, const int32_t *pair_hmm, const double *pair_w,
                                 const int64_t *col_off, const int32_t *cols, int n_hmm, const int64_t *hmm_off,
                                 const int32_t *retained, const int32_t *nongaps, int backbone_length,
                                 const int64_t *row_off, char *rows, int32_t *row_len) {
    try {
        if (!e || nq < 0 || (nq > 0 && (!qlen || !res_off || !residues || !pair_begin || !hmm_off || !retained || !nongaps ||
                                        !row_off || !rows || !row_len)) || backbone_length <= 0 || n_hmm <= 0)
            return fail(WITCH_ERR_ARG, "witch_graph_align: bad arguments");
        if (nq == 0) return WITCH_OK;
        if (n_hmm != e->H) return fail(WITCH_ERR_ARG, "witch_graph_align: n_hmm differs from the eHMM");
        const int np = pair_begin[nq];
        if (np < 0 || (np > 0 && (!pair_hmm || !pair_w || !col_off || !cols))) return fail(WITCH_ERR_ARG, "witch_graph_align: missing pair arrays");
        for (int q = 0; q < nq; q++)
            if (pair_begin[q + 1] < pair_begin[q] || qlen[q] < 0 || res_off[q] < 0 || row_off[q] < 0) return fail(WITCH_ERR_ARG, "witch_graph_align: bad offsets");
        for (int p = 0; p < np; p++) if (col_off[p] < 0) return fail(WITCH_ERR_ARG, "witch_graph_align: negative column offset");
        CUDA_TRY(cudaSetDevice(e->device));
        int Lcap = 1;
        long long res_total = 0, cols_total = 0, rows_total = 0;
        for (int q = 0; q < nq; q++) {
            Lcap = std::max(Lcap, qlen[q]);
            res_total = std::max<long long>(res_total, res_off[q] + qlen[q]);
            rows_total = std::max<long long>(rows_total, row_off[q] + 2LL * backbone_length + qlen[q] + 2);
            if (pair_begin[q + 1] - pair_begin[q] > GRAPH_KMAX) return fail(WITCH_ERR_LIMIT, "more than 16 included HMMs for one query");
            for (int p = pair_begin[q]; p < pair_begin[q + 1]; p++) {
                if (pair_hmm[p] < 0 || pair_hmm[p] >= n_hmm) return fail(WITCH_ERR_ARG, "witch_graph_align: HMM index out of range");
                cols_total = std::max<long long>(cols_total, col_off[p] + qlen[q]);
            }
        }
        auto up = [&](auto &buf, const auto *src, size_t n) {
            buf.alloc(n);
            if (n) CUDA_TRY(cudaMemcpy(buf.p, src, n * sizeof(*src), cudaMemcpyHostToDevice));
        };
        DevBuf<int> d_qlen, d_pb, d_ph, d_cols, d_ret, d_ng, d_rl;
        DevBuf<long long> d_ro, d_co, d_ho, d_rowoff;
        DevBuf<double> d_pw;
        DevBuf<char> d_res, d_rows;
        std::vector<long long> ro(res_off, res_off + nq), co(col_off, col_off + std::max(np, 0)), ho(hmm_off, hmm_off + n_hmm + 1),
            rwo(row_off, row_off + nq);
        up(d_qlen, qlen, nq); up(d_ro, ro.data(), nq); up(d_res, residues, (size_t)res_total);
        up(d_pb, pair_begin, nq + 1); up(d_ph, pair_hmm, np); up(d_pw, pair_w, np); up(d_co, co.data(), np);
        up(d_cols, cols, (size_t)cols_total); up(d_ho, ho.data(), n_hmm + 1);
        up(d_ret, retained, (size_t)hmm_off[n_hmm]); up(d_ng, nongaps, (size_t)hmm_off[n_hmm]);
        up(d_rowoff, rwo.data(), nq);
        d_rows.alloc((size_t)rows_total); d_rl.alloc(nq);
        const int Wcap = backbone_length + 3;
        const long long nblocks = Lcap / 32 + 2, nw16 = (Wcap + 31) / 16 + 2;
        long long slot = (long long)Wcap * 8 + (long long)Lcap * GRAPH_KMAX * 12 + (long long)Lcap * 4 +
                         ((Lcap + Wcap + 15) / 16) * 16 + nblocks * nw16 * 32 * 4;
        slot = (slot + 255) / 256 * 256;
        int occ = 1;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, graph_dp_kernel, 128, 0));
        long long grid = std::min<long long>(((long long)nq + 3) / 4, (long long)e->num_sms * std::max(occ, 1));
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        grid = std::max<long long>(1, std::min<long long>(grid, (long long)(0.5 * free_b / (4.0 * slot))));
        DevBuf<char> scratch; scratch.alloc((size_t)grid * 4 * slot);
        e->counter.alloc(64);
        CUDA_TRY(cudaMemset(e->counter.p, 0, sizeof(unsigned)));
        GraphWork G;
        G.nq = nq; G.qlen = d_qlen.p; G.res_off = d_ro.p; G.residues = d_res.p; G.pair_begin = d_pb.p; G.pair_hmm = d_ph.p;
        G.pair_w = d_pw.p; G.col_off = d_co.p; G.cols = d_cols.p; G.hmm_off = d_ho.p; G.retained = d_ret.p; G.nongaps = d_ng.p;
        G.backbone_length = backbone_length; G.row_off = d_rowoff.p; G.rows = d_rows.p; G.row_len = d_rl.p;
        G.counter = e->counter.p; G.scratch = scratch.p; G.slot_bytes = slot; G.Lcap = Lcap;
        WITCH_LAUNCH(graph_dp_kernel, (int)grid, 128)(G);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpy(rows, d_rows.p, (size_t)rows_total, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(row_len, d_rl.p, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost));
        return WITCH_OK;
    } catch (const std::exception &ex) {
        return fail(WITCH_ERR_CUDA, ex.what());
    }
}

extern "C" int witch_merge_rows(witch_ehmm *e, int nrows, const int64_t *row_off, const int32_t *row_len,
                                const uint8_t *is_backbone, const char *rows, int backbone_length, int32_t *gap_width,
                                int64_t *out_width, char *merged, int64_t merged_cap_width, char *masked) {
    try {
        if (!e || nrows < 0 || backbone_length <= 0 || !out_width || (nrows > 0 && (!row_off || !row_len || !is_backbone || !rows)))
            return fail(WITCH_ERR_ARG, "witch_merge_rows: bad arguments");
        CUDA_TRY(cudaSetDevice(e->device));
        const int B = backbone_length;
        long long total = 0;
        for (int r = 0; r < nrows; r++) {
            if (row_len[r] < 0 || row_off[r] < 0) return fail(WITCH_ERR_ARG, "witch_merge_rows: negative row offset/length");
            total = std::max<long long>(total, row_off[r] + row_len[r]);
        }
        DevBuf<long long> d_ro, d_gs;
        DevBuf<int> d_rl, d_w;
        DevBuf<uint8_t> d_bb;
        DevBuf<char> d_rows, d_out, d_msk;
        std::vector<long long> ro(row_off, row_off + nrows);
        auto up = [&](auto &buf, const auto *src, size_t n) {
            buf.alloc(n);
            if (n) CUDA_TRY(cudaMemcpy(buf.p, src, n * sizeof(*src), cudaMemcpyHostToDevice));
        };
        up(d_ro, ro.data(), nrows); up(d_rl, row_len, nrows); up(d_bb, is_backbone, nrows); up(d_rows, rows, (size_t)total);
        d_w.alloc(B + 1); d_gs.alloc(B + 2);
        CUDA_TRY(cudaMemset(d_w.p, 0, (size_t)(B + 1) * sizeof(int)));
        MergeWork W;
        W.nrows = nrows; W.row_off = d_ro.p; W.row_len = d_rl.p; W.is_backbone = d_bb.p; W.rows = d_rows.p; W.B = B;
        W.width = d_w.p; W.gap_start = d_gs.p; W.out = nullptr; W.masked = nullptr; W.out_width = 0;
        const int grid = std::max(1, (nrows + 3) / 4);
        if (nrows > 0) {
            WITCH_LAUNCH(merge_rows_kernel<false>, grid, 128)(W);
            g_launches++;
        }
        WITCH_LAUNCH(merge_scan_kernel, 1, 1024)(d_w.p, B, d_gs.p);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        long long width_total = 0;
        CUDA_TRY(cudaMemcpy(&width_total, d_gs.p + (B + 1), sizeof(long long), cudaMemcpyDeviceToHost));
        *out_width = width_total;
        if (gap_width) CUDA_TRY(cudaMemcpy(gap_width, d_w.p, (size_t)(B + 1) * sizeof(int), cudaMemcpyDeviceToHost));
        if (!merged && !masked) return WITCH_OK;   // plan only: the caller sizes its buffers from *out_width
        if (merged && merged_cap_width < width_total) return fail(WITCH_ERR_ARG, "witch_merge_rows: merged buffer narrower than the merged width");
        if (nrows == 0) return WITCH_OK;
        d_out.alloc((size_t)nrows * (size_t)width_total);
        if (masked) d_msk.alloc((size_t)nrows * (size_t)B);
        if (masked) CUDA_TRY(cudaMemset(d_msk.p, '-', (size_t)nrows * (size_t)B));
        W.out = d_out.p; W.masked = masked ? d_msk.p : nullptr; W.out_width = width_total;
        WITCH_LAUNCH(merge_rows_kernel<true>, grid, 128)(W);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        if (merged)
            CUDA_TRY(cudaMemcpy2D(merged, (size_t)merged_cap_width, d_out.p, (size_t)width_total, (size_t)width_total, (size_t)nrows,
                                  cudaMemcpyDeviceToHost));
        if (masked) CUDA_TRY(cudaMemcpy(masked, d_msk.p, (size_t)nrows * (size_t)B, cudaMemcpyDeviceToHost));
        return WITCH_OK;
    } catch (const std::exception &ex) {
        return fail(WITCH_ERR_CUDA, ex.what());
    }
}
