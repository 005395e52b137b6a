// tools/sim/simt.h -- a small SIMT simulator so that the CUDA sources of libwitch_b200 (kernels + C ABI) can be compiled
// with g++ and executed on the CPU, thread by thread, for TESTS when no GPU is at hand.
//
// TEST TOOLING ONLY. It is never loaded by the product path (witch_b200/_lib.py knows nothing about it; only
// tools/sim/sim_check.py points the ctypes binding at tools/sim/libwitch_sim.so) and it is far too slow to be a fallback:
// every CUDA thread is a user-level fiber; warp shuffles, __syncwarp and __syncthreads are rendez-vous points between
// fibers; blocks of a grid run one after the other; TMA bulk copies complete at issue time (so mbarrier waits are no-ops and a
// ring slot that is refilled too early is caught as wrong data, never hidden); shared-memory "addresses" are 32-bit
// offsets from one static arena. What it checks: the arithmetic, indexing, ring/boundary bookkeeping and the work
// distribution of the kernels exactly as written. What it cannot check: memory-model races, PTX semantics, performance.
#pragma once
#ifndef WITCH_HOST_SIM
#define WITCH_HOST_SIM 1
#endif
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

// ---------------------------------------------------------------- language shims
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __builtin_assume(x) ((void)0)
#define __isGlobal(p) true
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a) : x(a) {} };
static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
static inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }
template <class A, class B> static inline auto min(A a, B b) -> decltype(a + b) { return a < b ? a : b; }
template <class A, class B> static inline auto max(A a, B b) -> decltype(a + b) { return a > b ? a : b; }

namespace simt {
#if defined(__x86_64__) && !defined(SIMT_USE_UCONTEXT)
#define SIMT_FAST_SWITCH 1   // hand-written context switch (callee-saved registers + stack pointer): no sigprocmask system calls
extern "C" void simt_swap(void **save_sp, void *load_sp);
#endif
struct Fiber {
    ucontext_t ctx;
    void *sp = nullptr;
    char *stack = nullptr;
    uint3 tIdx{0, 0, 0};
    bool done = false;
    volatile unsigned *wait_ptr = nullptr;
    unsigned wait_val = 0;
};
struct Warp { volatile unsigned gen = 0; int count = 0, live = 0; uint64_t xbuf[32]; };
struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    int live = 0, bar_count = 0;
    volatile unsigned bar_gen = 0;
};
extern Block *g_blk;
extern Fiber *g_cur;
extern ucontext_t g_sched;
extern void *g_sched_sp;
extern uint3 g_blockIdx, g_blockDim, g_gridDim;
extern char g_smem_arena[];
extern unsigned long long g_switches;
void run_grid(unsigned grid, unsigned block, const std::function<void()> &body);
#ifdef SIMT_FAST_SWITCH
inline void yield() { g_switches++; simt_swap(&g_cur->sp, g_sched_sp); }
#else
inline void yield() { g_switches++; swapcontext(&g_cur->ctx, &g_sched); }
#endif
inline void wait_change(volatile unsigned *p, unsigned seen) {
    g_cur->wait_ptr = p; g_cur->wait_val = seen;
    while (*p == seen) yield();
    g_cur->wait_ptr = nullptr;
}
inline Warp &my_warp() { return g_blk->warps[g_cur->tIdx.x >> 5]; }
inline void warp_sync() {
    Warp &w = my_warp();
    const unsigned gen = w.gen;
    if (++w.count >= w.live) { w.count = 0; w.gen = gen + 1; } else wait_change(&w.gen, gen);
}
inline void block_sync() {
    Block &b = *g_blk;
    const unsigned gen = b.bar_gen;
    if (++b.bar_count >= b.live) { b.bar_count = 0; b.bar_gen = gen + 1; } else wait_change(&b.bar_gen, gen);
}
template <class T> inline T shfl(T v, int src, bool valid) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    Warp &w = my_warp();
    const int lane = g_cur->tIdx.x & 31;
    w.xbuf[lane] = 0;
    std::memcpy(&w.xbuf[lane], &v, sizeof(T));
    warp_sync();
    T r = v;
    if (valid) std::memcpy(&r, &w.xbuf[src], sizeof(T));
    warp_sync();
    return r;
}
template <class K, class... A> struct Launch {
    K k; unsigned grid, block;
    template <class... B> void operator()(B... args) const {
        K kk = k;
        run_grid(grid, block, [=]() { kk(args...); });
    }
};
template <class K, class G, class B> Launch<K> launcher(K k, G g, B b, size_t = 0, void * = nullptr) { return Launch<K>{k, (unsigned)g, (unsigned)b}; }
inline char *sptr(unsigned a) { return g_smem_arena + (intptr_t)(int32_t)a; }
}  // namespace simt

#define threadIdx (simt::g_cur->tIdx)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)
#define WITCH_LAUNCH(kernel, ...) simt::launcher(kernel, __VA_ARGS__)
#define WITCH_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(simt::g_smem_arena)

// ---------------------------------------------------------------- intrinsics
static inline void __syncthreads() { simt::block_sync(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_sync(); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { const int l = threadIdx.x & 31; return simt::shfl(v, l - d, l - d >= 0); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { const int l = threadIdx.x & 31; return simt::shfl(v, l + d, l + d < 32); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { const int l = threadIdx.x & 31; return simt::shfl(v, (l ^ m) & 31, true); }
template <class T> static inline T __shfl_sync(unsigned, T v, int s) { return simt::shfl(v, s & 31, true); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    simt::Warp &w = simt::my_warp();
    const int lane = threadIdx.x & 31;
    w.xbuf[lane] = pred ? 1 : 0;
    simt::warp_sync();
    unsigned r = 0;
    const int nl = std::min<int>(32, (int)blockDim.x - (int)(threadIdx.x & ~31u));
    for (int l = 0; l < nl; l++) r |= (unsigned)(w.xbuf[l] & 1) << l;
    simt::warp_sync();
    return r;
}
static inline long long clock64() { return 0; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
template <class T, class U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = o + (T)v; return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { T o = *p; *p = o | (T)v; return o; }

// ---------------------------------------------------------------- host versions of the PTX helpers of the kernels
namespace witch {
static inline unsigned smem_u32(const void *p) { return (unsigned)(int32_t)((const char *)p - simt::g_smem_arena); }
static inline float4 lds_f4(unsigned a) { return *reinterpret_cast<const float4 *>(simt::sptr(a)); }
static inline float4 lds_f4v(unsigned a) { return lds_f4(a); }
static inline uint4 lds_u4v(unsigned a) { return *reinterpret_cast<const uint4 *>(simt::sptr(a)); }
static inline float lds_f1(unsigned a) { return *reinterpret_cast<const float *>(simt::sptr(a)); }
static inline float lds_f1v(unsigned a) { return lds_f1(a); }
static inline void sts_f1(unsigned a, float v) { *reinterpret_cast<float *>(simt::sptr(a)) = v; }
static inline int lds_u8(unsigned a) { return *reinterpret_cast<const uint8_t *>(simt::sptr(a)); }
static inline int lds_i1v(unsigned a) { return *reinterpret_cast<const int *>(simt::sptr(a)); }
static inline void mbar_init(unsigned, int) {}
static inline void mbar_expect_tx(unsigned, unsigned) {}
static inline void mbar_wait(unsigned, unsigned) {}
static inline void tma_load_1d(unsigned dst, const void *src, unsigned bytes, unsigned) { std::memcpy(simt::sptr(dst), src, bytes); }
static inline void fence_mbar_init() {}
static inline void fence_proxy_async() {}
static inline bool wave_elect_one() { return (threadIdx.x & 31) == 0; }
}  // namespace witch
#define PIN32(x) ((void)0)
#define PIN64(x) ((void)0)

// ---------------------------------------------------------------- CUDA runtime stand-ins (host memory, synchronous)
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
struct cudaDeviceProp { int multiProcessorCount = 2; char name[64] = "host SIMT simulation"; };
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) {
    *p = (T *)aligned_alloc(256, (n + 255) / 256 * 256 + 256);
    return *p ? 0 : 2;
}
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)calloc(1, n); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind) {
    for (size_t r = 0; r < h; r++) std::memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
    return 0;
}
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char *cudaGetErrorString(cudaError_t) { return "host simulation error"; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); return 0; }
static inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)4 << 30; return 0; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *o, F, int, size_t) { *o = 1; return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (void *)1; return 0; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, int) { *s = (void *)1; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, int) { *e = (void *)1; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
