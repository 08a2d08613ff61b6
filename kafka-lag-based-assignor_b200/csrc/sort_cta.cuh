// sort_cta.cuh — CTA-wide bitonic sorts of (key, pay) elements.
//  * cta_bitonic_sort(): n = 2^m elements in SHARED MEMORY (used by the round fill when a
//    topic has more than 4096 consumers). Strides below 64 are done by each warp on a private
//    64-element window between two __syncwarp()s (no CTA barrier); larger strides take one
//    __syncthreads() per stage.
//  * cta_bitonic_sort_regs(): elements in REGISTERS (k_cta_topics, k_big_fill), see below.
#pragma once
#include "common.cuh"
#include "sort_warp.cuh"

namespace lagassign {

__device__ __forceinline__ void smem_cmpx(uint64_t* key, uint32_t* pay, int i, int p, bool asc) {
    uint64_t ka = key[i], kb = key[p];
    uint32_t pa = pay[i], pb = pay[p];
    bool swap = asc ? elem_less(kb, pb, ka, pa) : elem_less(ka, pa, kb, pb);
    if (swap) { key[i] = kb; key[p] = ka; pay[i] = pb; pay[p] = pa; }
}

// All threads of the CTA must call this (blockDim.x a multiple of 32, n >= 64, n a
// power of two). Sorts ascending by (key, pay). Ends with a __syncthreads().
__device__ __forceinline__ void cta_bitonic_sort(uint64_t* key, uint32_t* pay, int n) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    __syncthreads();
    for (int k = 2; k <= n; k <<= 1) {
        int j = k >> 1;
        // strides >= 64: one barrier per stage
        for (; j >= 64; j >>= 1) {
            for (int p = tid; p < (n >> 1); p += nt) {
                int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                smem_cmpx(key, pay, i, i | j, (i & k) == 0);
            }
            __syncthreads();
        }
        // strides < 64: each warp owns windows of 64 consecutive elements
        for (int w = warp; w < (n >> 6); w += nwarps) {
            const int base = w << 6;
            for (int jj = j; jj > 0; jj >>= 1) {
                int i = base + (((lane & ~(jj - 1)) << 1) | (lane & (jj - 1)));
                smem_cmpx(key, pay, i, i | jj, (i & k) == 0);
                __syncwarp();
            }
        }
        // phases up to k = 32 only touch a warp's own 64-element windows
        if (k >= 64) __syncthreads();
    }
    __syncthreads();
}

// ---- register-resident variant -----------------------------------------------------------
// n = NT*EPT elements of W words (see sort_warp.cuh; W = 1, 2 or 3) held in registers by a
// CTA of NT threads, logical index i = tid*EPT + r. Strides below EPT are in-register, below
// 32*EPT warp shuffles, and only the log2(NT/32) highest index bits cross warps: those
// stages exchange through shared memory (one __syncthreads() each, two alternating
// buffers). A 1024-element sort is 49 barrier-free stages + 6 exchanges instead of the 20
// barriers of cta_bitonic_sort().
//
// xbuf: 2 * W * n words of shared memory. All NT threads must call.
template <int W>
__device__ __forceinline__ bool words_less(const uint32_t (&a)[W], const uint32_t (&b)[W]) {
    if (W == 1) return a[0] < b[0];
    if (W == 2) return lt2(a[1], a[0], b[1], b[0]);
    return lt3(a[W - 1], a[W - 2], a[0], b[W - 1], b[W - 2], b[0]);
}

template <int W, int EPT, int NT>
__device__ __forceinline__ void cta_bitonic_sort_regs(uint32_t (&w)[W][EPT], uint32_t* xbuf) {
    constexpr int N = NT * EPT;
    constexpr int WARP_SPAN = 32 * EPT;
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    int flip = 0;
    // exchange through shared memory with thread tid ^ tmask; mirror: the partner's slot EPT-1-r
    auto cross_warp = [&](unsigned tmask, bool mirror, bool keep_max) {
        uint32_t* buf = xbuf + flip * (W * N);
        flip ^= 1;
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
#pragma unroll
            for (int x = 0; x < W; ++x) buf[(x * EPT + r) * NT + tid] = w[x][r];
        }
        __syncthreads();
        const unsigned pt = tid ^ tmask;
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            uint32_t o[W];
#pragma unroll
            for (int x = 0; x < W; ++x) o[x] = buf[(x * EPT + (mirror ? EPT - 1 - r : r)) * NT + pt];
            cx_remote<W, EPT>(w, r, o, keep_max);
        }
    };
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
        // mirror stage of the size-k merge (see sort_warp.cuh): i against i ^ (k-1)
        if (k <= EPT) {
#pragma unroll
            for (int r = 0; r < EPT; ++r)
                if (r < (r ^ (k - 1))) cx_local<W, EPT>(w, r, r ^ (k - 1));
        } else if (k <= WARP_SPAN) {
            const unsigned lm = (unsigned)(k / EPT - 1);
            const bool keep_max = (lane & (unsigned)(k / EPT / 2)) != 0;
            uint32_t o[EPT][W];
#pragma unroll
            for (int r = 0; r < EPT; ++r)
#pragma unroll
                for (int x = 0; x < W; ++x) o[r][x] = __shfl_xor_sync(0xffffffffu, w[x][EPT - 1 - r], lm);
#pragma unroll
            for (int r = 0; r < EPT; ++r) cx_remote<W, EPT>(w, r, o[r], keep_max);
        } else {
            cross_warp((unsigned)(k / EPT - 1), true, (tid & (unsigned)(k / EPT / 2)) != 0);
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            if (j >= WARP_SPAN) {
                cross_warp((unsigned)(j / EPT), false, (tid & (unsigned)(j / EPT)) != 0);
            } else if (j >= EPT) {
                const unsigned lj = (unsigned)(j / EPT);
                const bool keep_max = (lane & lj) != 0;
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    uint32_t o[W];
#pragma unroll
                    for (int x = 0; x < W; ++x) o[x] = __shfl_xor_sync(0xffffffffu, w[x][r], lj);
                    cx_remote<W, EPT>(w, r, o, keep_max);
                }
            } else {
#pragma unroll
                for (int r = 0; r < EPT; ++r)
                    if ((r & j) == 0) cx_local<W, EPT>(w, r, r | j);
            }
        }
    }
}

}  // namespace lagassign
