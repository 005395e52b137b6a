// Fiber scheduler of the host SIMT simulation (tools/sim/simt.h). TEST TOOLING ONLY.
#include "simt.h"

#ifdef SIMT_FAST_SWITCH
asm(R"(
.text
.globl simt_swap
.type simt_swap,@function
simt_swap:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_swap,.-simt_swap
)");
#endif

namespace simt {
void *g_sched_sp = nullptr;
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
#ifdef SIMT_FAST_SWITCH
    simt_swap(&f->sp, g_sched_sp);
    abort();   // a finished fiber is never resumed
#else
    swapcontext(&f->ctx, &g_sched);
#endif
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
#ifdef SIMT_FAST_SWITCH
            {   // initial frame: six callee-saved slots, then the entry point as the return address of simt_swap
                uintptr_t top = ((uintptr_t)stacks[t] + STACK) & ~(uintptr_t)15;
                void **sp = (void **)top;
                *--sp = nullptr;                  // fake return address of fiber_main (never used)
                *--sp = (void *)&fiber_main;
                for (int k = 0; k < 6; k++) *--sp = nullptr;
                f.sp = sp;
            }
#else
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = stacks[t];
            f.ctx.uc_stack.ss_size = STACK;
            f.ctx.uc_link = &g_sched;
            makecontext(&f.ctx, fiber_main, 0);
#endif
        }
        int remaining = (int)block;
        while (remaining > 0) {
            bool progressed = false;
            for (unsigned t = 0; t < block; t++) {
                Fiber &f = b.fibers[t];
                if (f.done) continue;
                if (f.wait_ptr && *f.wait_ptr == f.wait_val) continue;
                g_cur = &f;
#ifdef SIMT_FAST_SWITCH
                simt_swap(&g_sched_sp, f.sp);
#else
                swapcontext(&g_sched, &f.ctx);
#endif
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
