// sort_warp.cuh — bitonic sort of 32*EPT elements held in the registers of one
// warp. An element is W 32-bit words compared as one unsigned integer (word W-1
// most significant): W=3 is the general (64-bit key, 32-bit payload) element,
// W=2 / W=1 are the packed forms used when a topic's keys share a common prefix
// (see kernels_small.cuh). Elements must be pairwise distinct, except padding
// elements, which are identical to each other.
//
// Logical index i = lane*EPT + r ("blocked"): the low log2(EPT) index bits are
// register bits (compare-exchange without communication), the five high bits are
// lane bits (shuffle). The low bits are the ones a bitonic network visits most
// often, so blocked costs 15 shuffle stages at EPT=8 where a striped layout would
// cost 30. Input placement is arbitrary (the payload remembers where an element
// came from), so loads stay coalesced.
//
// Everything is branch-free: compare-exchange is a borrow chain / 64-bit compare /
// min-max plus selects, so all 32 lanes stay converged.
#pragma once
#include "common.cuh"

namespace lagassign {

// a < b for W-word unsigned integers (scalars passed most significant first).
__device__ __forceinline__ bool lt3(uint32_t a2, uint32_t a1, uint32_t a0, uint32_t b2, uint32_t b1, uint32_t b0) {
#ifdef LAGASSIGN_HOST_EMULATION
    return a2 != b2 ? a2 < b2 : (a1 != b1 ? a1 < b1 : a0 < b0);
#else
    uint32_t borrow;
    asm("{\n\t.reg .u32 t;\n\t"
        "sub.cc.u32 t, %1, %4;\n\t"
        "subc.cc.u32 t, %2, %5;\n\t"
        "subc.cc.u32 t, %3, %6;\n\t"
        "subc.u32 %0, 0, 0;\n\t}"
        : "=r"(borrow) : "r"(a0), "r"(a1), "r"(a2), "r"(b0), "r"(b1), "r"(b2));
    return borrow != 0;
#endif
}
__device__ __forceinline__ bool lt2(uint32_t a1, uint32_t a0, uint32_t b1, uint32_t b0) {
    return (((uint64_t)a1 << 32) | a0) < (((uint64_t)b1 << 32) | b0);
}

// Compare-exchange building blocks shared with the CTA-wide sort (sort_cta.cuh).
// Local pair, ascending: w[.][r] <- smaller, w[.][q] <- larger.
template <int W, int EPT>
__device__ __forceinline__ void cx_local(uint32_t (&w)[W][EPT], int r, int q) {
    if (W == 1) {
        const uint32_t lo = min(w[0][r], w[0][q]), hi = max(w[0][r], w[0][q]);
        w[0][r] = lo;
        w[0][q] = hi;
    } else {
        bool swap;   // w[q] < w[r]
        if (W == 2) swap = lt2(w[1][q], w[0][q], w[1][r], w[0][r]);
        else swap = lt3(w[W - 1][q], w[W - 2][q], w[0][q], w[W - 1][r], w[W - 2][r], w[0][r]);
#pragma unroll
        for (int x = 0; x < W; ++x) {
            const uint32_t a = w[x][r], b = w[x][q];
            w[x][r] = swap ? b : a;
            w[x][q] = swap ? a : b;
        }
    }
}
// Remote partner element o[]: keep the larger of (mine, other) if keep_max else the smaller.
template <int W, int EPT>
__device__ __forceinline__ void cx_remote(uint32_t (&w)[W][EPT], int r, const uint32_t (&o)[W], bool keep_max) {
    if (W == 1) {
        const uint32_t lo = min(w[0][r], o[0]), hi = max(w[0][r], o[0]);
        w[0][r] = keep_max ? hi : lo;       // SASS: VIMNMX + predicated VIMNMX
    } else {
        bool lt;   // other < mine
        if (W == 2) lt = lt2(o[1], o[0], w[1][r], w[0][r]);
        else lt = lt3(o[W - 1], o[W - 2], o[0], w[W - 1][r], w[W - 2][r], w[0][r]);
        const bool take = (lt != keep_max);   // equal elements are identical: either choice is the same
#pragma unroll
        for (int x = 0; x < W; ++x) w[x][r] = take ? o[x] : w[x][r];
    }
}

// The network is the direction-free form of the bitonic sorter: the merge of two ascending
// runs of k/2 starts with the MIRROR stage (i against i ^ (k-1)) and continues with the
// half-cleaners i against i ^ j, j = k/4 .. 1; every compare-exchange puts the smaller element
// at the smaller index. With no per-block direction a local pair costs two min/max (one word)
// and no select, and the per-lane predicates are plain lane bits.
template <int W, int EPT>
__device__ __forceinline__ void warp_bitonic_sort_words(uint32_t (&w)[W][EPT]) {
    constexpr int N = 32 * EPT;
    const unsigned lane = lane_id();
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
        if (k <= EPT) {
#pragma unroll
            for (int r = 0; r < EPT; ++r)
                if (r < (r ^ (k - 1))) cx_local<W, EPT>(w, r, r ^ (k - 1));
        } else {
            // mirror across lanes: partner lane ^ (k/EPT - 1), register slot EPT-1-r
            const unsigned lm = (unsigned)(k / EPT - 1);
            const bool keep_max = (lane & (unsigned)(k / EPT / 2)) != 0;
            uint32_t o[EPT][W];
#pragma unroll
            for (int r = 0; r < EPT; ++r)
#pragma unroll
                for (int x = 0; x < W; ++x) o[r][x] = __shfl_xor_sync(0xffffffffu, w[x][EPT - 1 - r], lm);
#pragma unroll
            for (int r = 0; r < EPT; ++r) cx_remote<W, EPT>(w, r, o[r], keep_max);
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            if (j >= EPT) {
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

// (64-bit key, 32-bit payload) interface over the 3-word sort.
template <int EPT>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t (&key)[EPT], uint32_t (&pay)[EPT]) {
    uint32_t w[3][EPT];
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        w[0][r] = pay[r];
        w[1][r] = (uint32_t)key[r];
        w[2][r] = (uint32_t)(key[r] >> 32);
    }
    warp_bitonic_sort_words<3, EPT>(w);
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        pay[r] = w[0][r];
        key[r] = ((uint64_t)w[2][r] << 32) | w[1][r];
    }
}

}  // namespace lagassign
