// Fiber scheduler of the host SIMT simulation (tools/sim/simt.h). TEST TOOLING ONLY.
#include "simt.h"

namespace simt {
Block *g_blk = nullptr;
Fiber *g_cur = nullptr;
ucontext_t g_sched;
uint3 g_blockIdx{0, 0, 0}, g_blockDim{1, 1, 1}, g_gridDim{1, 1, 1};
alignas(256) char g_smem_arena[256 * 1024];
unsigned long long g_switches = 0;
static const std::function<void()> *g_body = nullptr;
static const size_t STACK = 192 * 1024;

static void fiber_main() {
    (*g_body)();
    Fiber *f = g_cur;
    Block &b = *g_blk;
    f->done = true;
    Warp &w = b.warps[f->tIdx.x >> 5];
    // an exited thread no longer takes part in barriers: release the ones that are now complete
    w.live--; b.live--;
    if (w.live > 0 && w.count >= w.live) { w.count = 0; w.gen = w.gen + 1; }
    if (b.live > 0 && b.bar_count >= b.live) { b.bar_count = 0; b.bar_gen = b.bar_gen + 1; }
    swapcontext(&f->ctx, &g_sched);
}

void run_grid(unsigned grid, unsigned block, const std::function<void()> &body) {
    g_body = &body;
    g_gridDim = uint3{grid, 1, 1};
    g_blockDim = uint3{block, 1, 1};
    std::vector<char *> stacks(block);
    for (auto &s : stacks) s = (char *)malloc(STACK);
    for (unsigned bi = 0; bi < grid; bi++) {
        Block b;
        b.fibers.resize(block);
        b.warps.resize((block + 31) / 32);
        b.live = (int)block;
        for (unsigned t = 0; t < block; t++) b.warps[t >> 5].live++;
        g_blk = &b;
        g_blockIdx = uint3{bi, 0, 0};
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = b.fibers[t];
            f.tIdx = uint3{t, 0, 0};
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = stacks[t];
            f.ctx.uc_stack.ss_size = STACK;
            f.ctx.uc_link = &g_sched;
            makecontext(&f.ctx, fiber_main, 0);
        }
        int remaining = (int)block;
        while (remaining > 0) {
            bool progressed = false;
            for (unsigned t = 0; t < block; t++) {
                Fiber &f = b.fibers[t];
                if (f.done) continue;
                if (f.wait_ptr && *f.wait_ptr == f.wait_val) continue;
                g_cur = &f;
                swapcontext(&g_sched, &f.ctx);
                progressed = true;
                if (f.done) remaining--;
            }
            if (!progressed) throw std::runtime_error("SIMT simulation: deadlock (a collective was not reached by every thread)");
        }
    }
    for (auto s : stacks) free(s);
    g_blk = nullptr; g_cur = nullptr;
}
}  // namespace simt
