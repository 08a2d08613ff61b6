// sort_cta.cuh — bitonic sort of n = 2^m (key, pay) elements in shared memory by
// one CTA. Strides below 64 are done by each warp on a private 64-element window
// between two __syncwarp()s (no CTA barrier); larger strides take one
// __syncthreads() per stage.
#pragma once
#include "common.cuh"

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

}  // namespace lagassign
