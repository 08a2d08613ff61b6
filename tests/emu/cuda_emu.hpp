// cuda_emu.hpp — TEST INFRASTRUCTURE: a host stand-in for the slice of the CUDA execution model
// that the on-chip kernels (kernels_warp.cuh, kernels_cta.cuh, fill.cuh, sort_*.cuh) use, so that
// the very same device templates can be compiled with g++ and run on the CPU against the oracle
// (tests/test_kernels_emulated_cpu.py). One CTA runs at a time; every CUDA thread is a host
// thread; warp collectives and barriers are real rendezvous points, so code that forgets a
// __syncwarp()/__syncthreads() misbehaves here too (threads are NOT in lockstep).
// Nothing of this is linked into the product; it proves nothing about timing.
#pragma once
#include <pthread.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <vector_types.h>        // uint2, uint4, int4, uint3, dim3 (plain structs when not compiling with nvcc)
#include <vector_functions.h>    // make_uint4, ...
#include <driver_types.h>        // cudaStream_t, cudaError_t (declarations in types.cuh)

#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static          // one CTA at a time: a function-local static is that CTA's shared variable
#define __align__(n) alignas(n)

namespace emu {

constexpr int MAX_THREADS = 1024;

struct Cta {
    int n_threads = 0;
    pthread_barrier_t cta_bar;
    pthread_barrier_t warp_bar[MAX_THREADS / 32];
    uint64_t slot[MAX_THREADS];        // per-thread exchange word for shuffles / reductions
    int pred[MAX_THREADS];
    alignas(16) unsigned char dyn_smem[232 * 1024];
};

inline Cta& cta() { static Cta c; return c; }

struct Ids { uint3 thread, block; dim3 grid, blockdim; };
inline Ids& ids() { thread_local Ids i; return i; }

inline int tid() { return (int)ids().thread.x; }
inline void warp_sync() { pthread_barrier_wait(&cta().warp_bar[tid() >> 5]); }
inline void cta_sync() { pthread_barrier_wait(&cta().cta_bar); }

// publish v, let `read` look at the 32 slots of this warp, withdraw
template <class T, class F>
inline auto warp_exchange(T v, F read) {
    static_assert(sizeof(T) <= 8, "exchange slots are 64-bit");
    Cta& c = cta();
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    c.slot[tid()] = raw;
    warp_sync();
    const uint64_t* base = c.slot + (tid() & ~31);
    auto r = read([&](int lane) { T x; std::memcpy(&x, &base[lane & 31], sizeof(T)); return x; });
    warp_sync();
    return r;
}

// run one kernel launch: grid x block threads, CTAs one after the other
template <class Kernel>
void launch(int grid, int block, Kernel k) {
    Cta& c = cta();
    c.n_threads = block;
    for (int b = 0; b < grid; ++b) {
        std::memset(c.dyn_smem, 0xCD, sizeof c.dyn_smem);      // a CTA must not rely on what a previous one left behind
        pthread_barrier_init(&c.cta_bar, nullptr, (unsigned)block);
        for (int w = 0; w < (block + 31) / 32; ++w)
            pthread_barrier_init(&c.warp_bar[w], nullptr, (unsigned)std::min(32, block - 32 * w));
        std::vector<std::thread> ts;
        for (int t = 0; t < block; ++t) {
            ts.emplace_back([=] {
                Ids& i = ids();
                i.thread = uint3{(unsigned)t, 0, 0};
                i.block = uint3{(unsigned)b, 0, 0};
                i.grid = dim3((unsigned)grid, 1, 1);
                i.blockdim = dim3((unsigned)block, 1, 1);
                k();
            });
        }
        for (auto& t : ts) t.join();
        pthread_barrier_destroy(&c.cta_bar);
        for (int w = 0; w < (block + 31) / 32; ++w) pthread_barrier_destroy(&c.warp_bar[w]);
    }
}

}  // namespace emu

#define threadIdx (emu::ids().thread)
#define blockIdx (emu::ids().block)
#define gridDim (emu::ids().grid)
#define blockDim (emu::ids().blockdim)

// ---- synchronisation ------------------------------------------------------------------------
inline void __syncthreads() { emu::cta_sync(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_sync(); }
inline int __syncthreads_or(int p) {
    emu::Cta& c = emu::cta();
    c.pred[emu::tid()] = p;
    emu::cta_sync();
    int r = 0;
    for (int t = 0; t < c.n_threads; ++t) r |= c.pred[t] != 0;
    emu::cta_sync();
    return r;
}
inline int __syncthreads_and(int p) {
    emu::Cta& c = emu::cta();
    c.pred[emu::tid()] = p;
    emu::cta_sync();
    int r = 1;
    for (int t = 0; t < c.n_threads; ++t) r &= c.pred[t] != 0;
    emu::cta_sync();
    return r;
}

// ---- warp collectives (full masks only, as the kernels use them) -------------------------------
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    const int me = emu::tid() & 31;
    return emu::warp_exchange(v, [&](auto at) { return at(me ^ lane_mask); });
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
    const int me = emu::tid() & 31;
    return emu::warp_exchange(v, [&](auto at) { return (me + (int)delta < 32) ? at(me + (int)delta) : at(me); });
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) {
    return emu::warp_exchange(v, [&](auto at) { return at(src); });
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    const int me = emu::tid() & 31;
    return emu::warp_exchange(v, [&](auto at) { return (me - (int)delta >= 0) ? at(me - (int)delta) : at(me); });
}
inline unsigned __match_any_sync(unsigned, unsigned v) {
    return emu::warp_exchange(v, [&](auto at) { unsigned m = 0; for (int l = 0; l < 32; ++l) m |= (unsigned)(at(l) == v) << l; return m; });
}
inline unsigned __ballot_sync(unsigned, int p) {
    return emu::warp_exchange(p, [&](auto at) { unsigned m = 0; for (int l = 0; l < 32; ++l) m |= (unsigned)(at(l) != 0) << l; return m; });
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    return emu::warp_exchange(v, [&](auto at) { unsigned r = 0; for (int l = 0; l < 32; ++l) r |= at(l); return r; });
}
inline unsigned __reduce_and_sync(unsigned, unsigned v) {
    return emu::warp_exchange(v, [&](auto at) { unsigned r = ~0u; for (int l = 0; l < 32; ++l) r &= at(l); return r; });
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    return emu::warp_exchange(v, [&](auto at) { unsigned r = ~0u; for (int l = 0; l < 32; ++l) r = at(l) < r ? at(l) : r; return r; });
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    return emu::warp_exchange(v, [&](auto at) { unsigned r = 0; for (int l = 0; l < 32; ++l) r = at(l) > r ? at(l) : r; return r; });
}
inline int __any_sync(unsigned, int p) {
    return emu::warp_exchange(p, [&](auto at) { int r = 0; for (int l = 0; l < 32; ++l) r |= at(l) != 0; return r; });
}
inline int __all_sync(unsigned, int p) {
    return emu::warp_exchange(p, [&](auto at) { int r = 1; for (int l = 0; l < 32; ++l) r &= at(l) != 0; return r; });
}

// ---- memory and arithmetic intrinsics ------------------------------------------------------------
namespace emu {
// the hardware faults on a misaligned vector access; x86 would not: check it here
inline void check_aligned(const void* p, size_t n) {
    if (reinterpret_cast<uintptr_t>(p) % n != 0) {
        std::fprintf(stderr, "cuda_emu: %zu-byte access at misaligned address %p\n", n, p);
        std::abort();
    }
}
}  // namespace emu
template <class T> inline T __ldg(const T* p) { emu::check_aligned(p, sizeof(T)); return *p; }
template <class T> inline T __ldcs(const T* p) { emu::check_aligned(p, sizeof(T)); return *p; }
template <class T, class V> inline void __stcs(T* p, V v) { emu::check_aligned(p, sizeof(T)); *p = (T)v; }
inline void __stcs(int4* p, int4 v) { emu::check_aligned(p, 16); *p = v; }
// atomics: the CUDA threads are concurrent host threads, so these must be real atomics
template <class T> inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> inline T atomicMax(T* p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __clz(unsigned x) { return x == 0 ? 32 : __builtin_clz(x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
template <class T> inline T min(T a, T b) { return b < a ? b : a; }
template <class T> inline T max(T a, T b) { return a < b ? b : a; }
inline int64_t min(int64_t a, int b) { return a < b ? a : b; }
inline int64_t max(int64_t a, int b) { return a > b ? a : b; }
inline int min(int a, unsigned b) { return a < (int)b ? a : (int)b; }
