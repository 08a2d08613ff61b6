// kernels_warp.cuh — topics of up to 256 partitions: the whole of assignTopic
// (LagBasedPartitionAssignor.java:204-277) for one topic in ONE WARP, partitions in
// registers. Lags are read from HBM once and the two result arrays written once: 16 B per
// partition (20 B with explicit partition ids). Larger topics: kernels_cta.cuh (one CTA,
// up to 4096 partitions), radix.cuh (device-wide).
#pragma once
#include "common.cuh"
#include "types.cuh"
#include "fill.cuh"
#include "sort_cta.cuh"
#include "sort_warp.cuh"

namespace lagassign {

#ifndef LAGASSIGN_WARP_MINBLOCKS
#define LAGASSIGN_WARP_MINBLOCKS 4   // CTAs per SM the register allocation aims for
#endif
constexpr int SMALL_WARPS = 8;   // warps (= topics in flight) per CTA of k_warp_topics

struct __align__(16) WarpScratch {
    uint64_t lag_key[WARP_MAX_P];   // sorted keys, for the round fill / totals
    uint32_t out_a[WARP_MAX_P];     // out_partition staging (blocked -> coalesced)
    uint32_t out_b[WARP_MAX_P];     // out_consumer staging
    uint32_t part_tab[WARP_MAX_P];  // partition ids in ascending order (unordered-input path)
};

// Staging of a warp's 32*EPT result words: written in sort order ("blocked": lane
// owns EPT consecutive words) and read back striped so the global store is
// coalesced. 16-byte chunks are XOR-swizzled so neither side has bank conflicts.
template <int EPT>
__device__ __forceinline__ int stage_addr(int i) {
    if (EPT < 4) return i;
    const int chunk = i >> 2;
    return ((chunk ^ ((chunk >> 3) & 1)) << 2) | (i & 3);
}
template <int EPT>
__device__ __forceinline__ void stage_store(uint32_t* buf, unsigned lane, const uint32_t (&v)[EPT]) {
    if constexpr (EPT >= 4) {
#pragma unroll
        for (int c = 0; c < EPT / 4; ++c) {
            const int i = (int)lane * EPT + c * 4;
            *reinterpret_cast<uint4*>(buf + stage_addr<EPT>(i)) = make_uint4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
        }
    } else if constexpr (EPT == 2) {
        *reinterpret_cast<uint2*>(buf + lane * 2) = make_uint2(v[0], v[1]);
    } else {
        buf[lane] = v[0];
    }
}

template <int EPT> struct Log2Slots;
template <> struct Log2Slots<1> { static constexpr int value = 5; };
template <> struct Log2Slots<2> { static constexpr int value = 6; };
template <> struct Log2Slots<4> { static constexpr int value = 7; };
template <> struct Log2Slots<8> { static constexpr int value = 8; };

// Sort (key, pay) ascending, pay < 32*EPT distinct for live elements. Padding elements
// carry pay = PAY_SENTINEL and key = ref, the key of some live element (so they are neutral
// in the width analysis below; pay all-ones makes every packed form of a padding element
// all-ones without a select). Keys of one topic usually share most of their high bits (real
// lags are small numbers), so the sort runs on the narrowest element that still orders them:
// the s significant key bits above the B payload bits, in one 32-bit word (s + B <= 32), one
// 64-bit word (s + B <= 64) or the general three words. The choice is uniform across the warp.
// The general three-word sort and the multi-round fill are kept out of line so
// that their register needs do not set the occupancy of the common path.
template <int EPT>
static __device__ __noinline__ void warp_sort_general(uint64_t (&key)[EPT], uint32_t (&pay)[EPT]) {
#pragma unroll
    for (int r = 0; r < EPT; ++r) key[r] = (pay[r] != PAY_SENTINEL) ? key[r] : KEY_SENTINEL;
    warp_bitonic_sort<EPT>(key, pay);
}

// Returns true when key[] holds the sorted keys on return; false when only pay[] (the
// sorted positions) is valid, which happens on the truncated path below. On return slot r of
// lane l is sorted position l*EPT + r; slots at positions >= P hold padding (pay then is NOT
// the sentinel any more: test the position, not pay).
template <int EPT>
__device__ __forceinline__ bool warp_sort_adaptive(uint64_t (&key)[EPT], uint32_t (&pay)[EPT], uint64_t ref, int P,
                                                   bool allow_trunc) {
    constexpr int B = Log2Slots<EPT>::value;
    constexpr uint32_t PAYMASK = (1u << B) - 1u;
    const unsigned lane = lane_id();
    uint32_t x_lo = 0, x_hi = 0;
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        x_lo |= (uint32_t)key[r] ^ (uint32_t)ref;
        x_hi |= (uint32_t)(key[r] >> 32) ^ (uint32_t)(ref >> 32);
    }
    const uint32_t xlo = __reduce_or_sync(0xffffffffu, x_lo);
    const uint32_t xhi = __reduce_or_sync(0xffffffffu, x_hi);
    const int s = xhi ? 64 - __clz(xhi) : (xlo ? 32 - __clz(xlo) : 0);   // bits in which keys differ
    if (s == 0) {
        // every lag equal (an idle topic, typically all zero): list order is the sorted order
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            pay[r] = (uint32_t)((int)lane * EPT + r);
            key[r] = ref;
        }
        return true;
    }
    if (s + B <= 32) {
        const uint32_t mask = 0xFFFFFFFFu >> (32 - s);   // 1 <= s <= 27
        uint32_t w[1][EPT];
#pragma unroll
        for (int r = 0; r < EPT; ++r) w[0][r] = (((uint32_t)key[r] & mask) << B) | pay[r];   // padding: all ones
        warp_bitonic_sort_words<1, EPT>(w);
        const uint64_t hi = ref & ~(uint64_t)mask;
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            pay[r] = w[0][r] & PAYMASK;
            key[r] = hi | (uint64_t)(w[0][r] >> B);
        }
        return true;
    }
    if (allow_trunc) {
        // Optimistic one-word sort on the TOP 32-B of the s differing key bits. If no two
        // neighbours of the result share that truncated key, the order is already the exact
        // (key, position) order: for high-entropy lags (random 40-bit values, 256 per topic)
        // that is ~99.8 % of topics, at a third of the instructions of the 64-bit sort.
        // Otherwise fall through to the exact sort; key[] / pay[] are still untouched.
        const int sh = s - (32 - B);                 // s > 32 - B here
        uint32_t w[1][EPT];
#pragma unroll
        for (int r = 0; r < EPT; ++r) w[0][r] = (((uint32_t)(key[r] >> sh)) << B) | pay[r];   // bits >= s drop out
        warp_bitonic_sort_words<1, EPT>(w);
        bool dup = false;
#pragma unroll
        for (int r = 0; r + 1 < EPT; ++r) {
            const int i = (int)lane * EPT + r;
            dup |= (i + 1 < P) && ((w[0][r] ^ w[0][r + 1]) <= PAYMASK);
        }
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, w[0][0], 1);
        dup |= (lane < 31) && ((int)lane * EPT + EPT < P) && ((w[0][EPT - 1] ^ nxt) <= PAYMASK);
        if (!__any_sync(0xffffffffu, dup)) {
#pragma unroll
            for (int r = 0; r < EPT; ++r) pay[r] = w[0][r] & PAYMASK;
            return false;
        }
    }
    if (s + B <= 64) {
        const uint64_t mask = ~0ull >> (64 - s);   // 1 <= s <= 59
        uint32_t w[2][EPT];
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const uint64_t v = ((key[r] & mask) << B) | (uint64_t)(int64_t)(int32_t)pay[r];   // padding: all ones
            w[0][r] = (uint32_t)v;
            w[1][r] = (uint32_t)(v >> 32);
        }
        warp_bitonic_sort_words<2, EPT>(w);
        const uint64_t hi = ref & ~mask;
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const uint64_t v = ((uint64_t)w[1][r] << 32) | w[0][r];
            pay[r] = (uint32_t)v & PAYMASK;
            key[r] = hi | (v >> B);
        }
    } else {
        // Hand the out-of-line sort COPIES, so that key[] / pay[] are not address-taken and stay in
        // registers on the common path (they cost ~22 local-memory stores per topic at config 4 otherwise:
        // 0.137 -> 0.121 ms on B200).
        uint64_t k3[EPT];
        uint32_t p3[EPT];
#pragma unroll
        for (int r = 0; r < EPT; ++r) { k3[r] = key[r]; p3[r] = pay[r]; }
        warp_sort_general<EPT>(k3, p3);
#pragma unroll
        for (int r = 0; r < EPT; ++r) { key[r] = k3[r]; pay[r] = p3[r]; }
    }
    return true;
}

static __device__ __noinline__ void warp_rounds_path(int P, int C, const int32_t* __restrict__ cset, WarpScratch& ws,
                                                     int64_t* tot, int32_t* cnt);

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// PLAIN: the batch has no explicit partition ids and no totals/counts outputs (the common
// call); the paths for those are compiled out.
// LAGMODE: where the lags come from (types.cuh): with raw offsets, computePartitionLag (Main:376-404) is evaluated
// right here, as the keys are built, and the lags never exist in memory.
template <int EPT, bool PLAIN, int LAGMODE>
__device__ __forceinline__ void warp_topic(const DevBatch& b, int t, WarpScratch& ws) {
    const unsigned lane = lane_id();
    const int64_t off = b.topic_off[t];
    const int P = (int)(b.topic_off[t + 1] - off);
    const int s = b.topic_cset[t];
    const int64_t coff = b.cset_off[s];
    const int64_t C64 = b.cset_off[s + 1] - coff;
    const int32_t* __restrict__ cset = b.cset_rank + coff;
    auto lag_at = [&](int m) -> uint64_t { return (uint64_t)lag_load<LAGMODE>(b, off + m); };

    uint64_t key[EPT];
    uint32_t pay[EPT];
    // Coalesced load; which register slot an element lands in does not matter,
    // the payload carries its list position (Main:228 sorts the whole list).
    const uint64_t ref = lag_at(0) ^ LAG_KEY_MASK;   // element 0 is live (P >= 1)
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        const int m = r * 32 + (int)lane;
        const bool live = m < P;
        key[r] = live ? (lag_at(m) ^ LAG_KEY_MASK) : ref;
        pay[r] = live ? (uint32_t)m : PAY_SENTINEL;
    }

    bool use_tab = false;
    if constexpr (!PLAIN) {
        if (b.part != nullptr) {
            // Ties on lag break on partition id ascending (Main:230-231). A list that
            // is already ascending in partition id lets the position stand in for it;
            // otherwise order it by partition id first (stable by position).
            const int32_t* __restrict__ partp = b.part + off;
            bool bad = false;
#pragma unroll
            for (int r = 0; r < EPT; ++r) {
                const int m = r * 32 + (int)lane;
                if (m + 1 < P) bad |= partp[m] > partp[m + 1];
            }
            if (__any_sync(0xffffffffu, bad)) {
                uint64_t k2[EPT];
                uint32_t p2[EPT];
                const uint64_t ref2 = (uint64_t)((uint32_t)partp[0] ^ 0x80000000u);
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    const int m = r * 32 + (int)lane;
                    const bool live = m < P;
                    k2[r] = live ? (uint64_t)((uint32_t)partp[m] ^ 0x80000000u) : ref2;
                    p2[r] = live ? (uint32_t)m : PAY_SENTINEL;
                }
                warp_sort_adaptive<EPT>(k2, p2, ref2, P, /*allow_trunc=*/false);
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    const int i = (int)lane * EPT + r;
                    const bool live = i < P;
                    if (live) ws.part_tab[i] = (uint32_t)k2[r] ^ 0x80000000u;
                    key[r] = live ? (lag_at((int)p2[r]) ^ LAG_KEY_MASK) : ref;
                    pay[r] = live ? (uint32_t)i : PAY_SENTINEL;
                }
                use_tab = true;
                __syncwarp();
            }
        }
    }

    // Main:228-235 (pay is the list position here unless the ids had to be ordered first)
    const bool keys_exact = warp_sort_adaptive<EPT>(key, pay, ref, P, /*allow_trunc=*/!use_tab);

    // out_partition: lane l now holds sorted positions l*EPT .. l*EPT+EPT-1
    uint32_t v[EPT];
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        v[r] = pay[r];
        if constexpr (!PLAIN) {
            if (b.part != nullptr) {
                const int i = (int)lane * EPT + r;
                v[r] = 0;
                if (i < P) v[r] = use_tab ? ws.part_tab[pay[r]] : (uint32_t)b.part[off + pay[r]];
            }
        }
    }
    const bool rounds = (int64_t)P > C64;
    const bool need_keys = rounds || (!PLAIN && b.out_total != nullptr);
    if (need_keys) {
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int i = (int)lane * EPT + r;
            if (i < P) ws.lag_key[i] = keys_exact ? key[r] : (lag_at((int)pay[r]) ^ LAG_KEY_MASK);
        }
    }
    int32_t* __restrict__ outp = b.out_part + off;
    int32_t* __restrict__ outc = b.out_cons + off;
    // 16-byte stores straight from the registers when the topic's slices are aligned (always,
    // when partition counts are multiples of 4); otherwise staged through shared memory so
    // that the 4-byte stores are coalesced.
    const bool vec = EPT >= 4 && aligned16(outp) && aligned16(outc) && aligned16(cset);
    if (EPT == 1) {
        if ((int)lane < P) __stcs(outp + lane, (int32_t)v[0]);
    } else if (vec) {
#pragma unroll
        for (int c = 0; c < EPT / 4; ++c) {
            const int i = (int)lane * EPT + c * 4;
            if (i + 4 <= P) {
                __stcs(reinterpret_cast<int4*>(outp + i), make_int4((int)v[c * 4], (int)v[c * 4 + 1], (int)v[c * 4 + 2], (int)v[c * 4 + 3]));
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (i + e < P) __stcs(outp + i + e, (int32_t)v[c * 4 + e]);
            }
        }
    } else {
        stage_store<EPT>(ws.out_a, lane, v);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int m = r * 32 + (int)lane;
            if (m < P) __stcs(outp + m, (int32_t)ws.out_a[stage_addr<EPT>(m)]);
        }
    }

    int64_t* tot = nullptr;
    int32_t* cnt = nullptr;
    if constexpr (!PLAIN) {
        if (b.out_total != nullptr) {
            tot = b.out_total + b.tc_off[t];
            cnt = b.out_count + b.tc_off[t];
        }
    }
    if (!rounds) {
        // P <= C: one partial round on all-zero totals, so the k-th partition in
        // sorted order goes to the k-th member in id order.
        if (vec) {
#pragma unroll
            for (int c = 0; c < EPT / 4; ++c) {
                const int i = (int)lane * EPT + c * 4;
                if (i + 4 <= P) {
                    __stcs(reinterpret_cast<int4*>(outc + i), __ldg(reinterpret_cast<const int4*>(cset + i)));
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (i + e < P) __stcs(outc + i + e, cset[i + e]);
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < EPT; ++r) {
                const int m = r * 32 + (int)lane;
                if (m < P) __stcs(outc + m, cset[m]);
            }
        }
        if constexpr (!PLAIN) {
            if (tot != nullptr) {
                __syncwarp();      // lag_key[] written above in sort order
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    const int m = r * 32 + (int)lane;
                    if (m < P) { tot[m] = (int64_t)(ws.lag_key[m] ^ LAG_KEY_MASK); cnt[m] = 1; }
                }
            }
        }
        // members beyond P keep total 0 / count 0 (buffers are zero-filled)
    } else {
        __syncwarp();
        warp_rounds_path(P, (int)C64, cset, ws, tot, cnt);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int m = r * 32 + (int)lane;
            if (m < P) __stcs(outc + m, (int32_t)ws.out_b[m]);
        }
    }
    __syncwarp();
}

static __device__ __noinline__ void warp_rounds_path(int P, int C, const int32_t* __restrict__ cset, WarpScratch& ws,
                                                     int64_t* tot, int32_t* cnt) {
    auto lag_at = [&](int64_t k) { return (int64_t)(ws.lag_key[k] ^ LAG_KEY_MASK); };
    auto put = [&](int64_t k, int32_t rank) { ws.out_b[k] = (uint32_t)rank; };
    warp_round_fill_dispatch(P, C, cset, lag_at, put, tot, cnt);
}

// One warp per topic of `list` (topics with 1 <= P <= 256 and a non-empty consumer set).
template <bool PLAIN, int LAGMODE>
__global__ void __launch_bounds__(SMALL_WARPS * 32, LAGASSIGN_WARP_MINBLOCKS)
k_warp_topics(DevBatch b, const int32_t* __restrict__ list, int n_list) {
    __shared__ WarpScratch scratch[SMALL_WARPS];
    const int warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * SMALL_WARPS + warp;
    if (idx >= n_list) return;
    const int t = list[idx];
    const int P = (int)(b.topic_off[t + 1] - b.topic_off[t]);
    WarpScratch& ws = scratch[warp];
    if (P <= 32) warp_topic<1, PLAIN, LAGMODE>(b, t, ws);
    else if (P <= 64) warp_topic<2, PLAIN, LAGMODE>(b, t, ws);
    else if (P <= 128) warp_topic<4, PLAIN, LAGMODE>(b, t, ws);
    else warp_topic<8, PLAIN, LAGMODE>(b, t, ws);
}


}  // namespace lagassign
