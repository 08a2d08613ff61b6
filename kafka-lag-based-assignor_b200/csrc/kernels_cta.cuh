// kernels_cta.cuh — topics of 257..4096 partitions: the whole of assignTopic
// (LagBasedPartitionAssignor.java:204-277) for one topic in ONE CTA, partitions in the
// registers of its 256 threads. Same I/O as the warp kernel (kernels_warp.cuh): lags read
// once, the two result arrays written once.
#pragma once
#include "common.cuh"
#include "types.cuh"
#include "fill.cuh"
#include "sort_cta.cuh"
#include "sort_warp.cuh"

namespace lagassign {

// ---- one CTA per topic, 256 < P <= 4096 ----------------------------------------------
// 256 threads, the topic's partitions in REGISTERS (EPT per thread, N = 256*EPT slots, one
// kernel instantiation per size class N = 512 / 1024 / 2048 / 4096 so that shared memory and
// occupancy match the class). Same ideas as the warp kernel: branch-free register sort
// (shuffles inside a warp, shared-memory exchanges only for the three index bits that cross
// warps), adaptive element width with the optimistic one-word sort, staged coalesced stores.
// (Not yet ported from the warp kernel: neutral padding and 16-byte stores from registers.)

constexpr int CTA_THREADS = 256;

template <int EPT>
struct CtaRegSmem {
    static constexpr int N = CTA_THREADS * EPT;
    static constexpr size_t xbuf_off = 0;                                 // sort exchanges / round fill / output staging
    static constexpr size_t skey_off = xbuf_off + (size_t)2 * 3 * N * 4;  // sorted keys, for the fill and the totals
    static constexpr size_t tab_off = skey_off + (size_t)N * 8;           // partition ids in ascending order (unordered input)
    static constexpr size_t red_off = tab_off + (size_t)N * 4;            // reductions, first words
    static constexpr size_t total = red_off + (size_t)(CTA_THREADS + 64) * 4;
};

__device__ __forceinline__ int pow2_ceil(int x) {
    return x <= 1 ? 1 : 1 << (32 - __clz(x - 1));
}

// Sort the CTA's (key, pay) pairs (position-resident afterwards: slot r of thread t is sorted
// position t*EPT + r). Returns true when key[] is exact on return (see warp_sort_adaptive).
template <int EPT>
__device__ __forceinline__ bool cta_sort_adaptive(uint64_t (&key)[EPT], uint32_t (&pay)[EPT], int P, bool allow_trunc,
                                                  uint32_t* xbuf, uint32_t* red) {
    constexpr int NT = CTA_THREADS;
    constexpr int B = 8 + Log2E<EPT>::value;          // log2(N)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // bits in which the live keys differ: OR ^ AND over the CTA
    uint32_t o_lo = 0, o_hi = 0, a_lo = ~0u, a_hi = ~0u;
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        if (pay[r] != PAY_SENTINEL) {
            o_lo |= (uint32_t)key[r]; o_hi |= (uint32_t)(key[r] >> 32);
            a_lo &= (uint32_t)key[r]; a_hi &= (uint32_t)(key[r] >> 32);
        }
    }
    o_lo = __reduce_or_sync(0xffffffffu, o_lo); o_hi = __reduce_or_sync(0xffffffffu, o_hi);
    a_lo = __reduce_and_sync(0xffffffffu, a_lo); a_hi = __reduce_and_sync(0xffffffffu, a_hi);
    __syncthreads();
    if (lane == 0) { red[warp * 4] = o_lo; red[warp * 4 + 1] = o_hi; red[warp * 4 + 2] = a_lo; red[warp * 4 + 3] = a_hi; }
    __syncthreads();
    o_lo = o_hi = 0; a_lo = a_hi = ~0u;
#pragma unroll
    for (int w2 = 0; w2 < NT / 32; ++w2) {
        o_lo |= red[w2 * 4]; o_hi |= red[w2 * 4 + 1]; a_lo &= red[w2 * 4 + 2]; a_hi &= red[w2 * 4 + 3];
    }
    const uint64_t diff = (((uint64_t)(o_hi ^ a_hi)) << 32) | (o_lo ^ a_lo);
    const int s = bits_of(diff);
    const uint64_t common = (((uint64_t)a_hi) << 32) | a_lo;      // bits shared by all live keys (above s)
    if (s == 0) {
        // every lag equal (an idle topic, typically all zero): list order is the sorted order
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int i = tid * EPT + r;
            pay[r] = i < P ? (uint32_t)i : PAY_SENTINEL;
            key[r] = i < P ? common : KEY_SENTINEL;
        }
        return true;
    }
    if (s + B <= 32 || allow_trunc) {
        const int sh = max(0, s + B - 32);
        uint32_t w[1][EPT];
#pragma unroll
        for (int r = 0; r < EPT; ++r)
            w[0][r] = (pay[r] != PAY_SENTINEL)
                          ? ((((uint32_t)(key[r] >> sh)) & ((1u << (32 - B)) - 1u)) << B) | pay[r] : 0xFFFFFFFFu;
        cta_bitonic_sort_regs<1, EPT, NT>(w, xbuf);
        int dup = 0;
        if (sh > 0) {      // truncated: accept only if no two neighbours share the truncated key
#pragma unroll
            for (int r = 0; r + 1 < EPT; ++r)
                dup |= (tid * EPT + r + 1 < P) && ((w[0][r] >> B) == (w[0][r + 1] >> B));
            red[64 + tid] = w[0][0];
            __syncthreads();
            if (tid + 1 < NT) dup |= (tid * EPT + EPT < P) && ((w[0][EPT - 1] >> B) == (red[64 + tid + 1] >> B));
            dup = __syncthreads_or(dup);
        }
        if (!dup) {
            const uint64_t mask = s >= 64 ? ~0ull : ((1ull << s) - 1ull);
#pragma unroll
            for (int r = 0; r < EPT; ++r) {
                pay[r] = w[0][r] & ((1u << B) - 1u);
                if (sh == 0) key[r] = (common & ~mask) | (uint64_t)(w[0][r] >> B);
            }
            return sh == 0;
        }
    }
    if (s + B <= 64) {
        const uint64_t mask = ~0ull >> (64 - s);       // s >= 1 here
        uint32_t w[2][EPT];
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const uint64_t v = (pay[r] != PAY_SENTINEL) ? (((key[r] & mask) << B) | pay[r]) : ~0ull;
            w[0][r] = (uint32_t)v; w[1][r] = (uint32_t)(v >> 32);
        }
        cta_bitonic_sort_regs<2, EPT, NT>(w, xbuf);
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const uint64_t v = ((uint64_t)w[1][r] << 32) | w[0][r];
            pay[r] = (uint32_t)v & ((1u << B) - 1u);
            key[r] = (common & ~mask) | (v >> B);
        }
    } else {
        uint32_t w[3][EPT];
#pragma unroll
        for (int r = 0; r < EPT; ++r) { w[0][r] = pay[r]; w[1][r] = (uint32_t)key[r]; w[2][r] = (uint32_t)(key[r] >> 32); }
        cta_bitonic_sort_regs<3, EPT, NT>(w, xbuf);
#pragma unroll
        for (int r = 0; r < EPT; ++r) { pay[r] = w[0][r]; key[r] = ((uint64_t)w[2][r] << 32) | w[1][r]; }
    }
    return true;
}

template <int EPT>
__global__ void __launch_bounds__(CTA_THREADS, (EPT >= 16) ? 1 : 2)
k_cta_topics(DevBatch b, const int32_t* __restrict__ list, int n_list) {
#ifdef LAGASSIGN_HOST_EMULATION
    unsigned char* smem_raw = emu::cta().dyn_smem;
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    using L = CtaRegSmem<EPT>;
    constexpr int NT = CTA_THREADS, N = L::N;
    uint32_t* xbuf = reinterpret_cast<uint32_t*>(smem_raw + L::xbuf_off);
    uint64_t* skey = reinterpret_cast<uint64_t*>(smem_raw + L::skey_off);
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem_raw + L::tab_off);
    uint32_t* red = reinterpret_cast<uint32_t*>(smem_raw + L::red_off);
    const int tid = threadIdx.x;

    for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
        const int t = list[li];
        const int64_t off = b.topic_off[t];
        const int P = (int)(b.topic_off[t + 1] - off);
        const int s = b.topic_cset[t];
        const int64_t coff = b.cset_off[s];
        const int64_t C64 = b.cset_off[s + 1] - coff;
        const int32_t* __restrict__ cset = b.cset_rank + coff;

        uint64_t key[EPT];
        uint32_t pay[EPT];
        // coalesced load; which slot an element lands in does not matter (the payload is its list position)
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int m = r * NT + tid;
            const bool live = m < P;
            key[r] = live ? ((uint64_t)lag_load<LAG_EITHER>(b, off + m) ^ LAG_KEY_MASK) : KEY_SENTINEL;
            pay[r] = live ? (uint32_t)m : PAY_SENTINEL;
        }
        bool use_tab = false;
        if (b.part != nullptr) {
            // ties on lag break on partition id ascending (Main:230-231): if the list is not already in
            // that order, order it by partition id first (stable by position)
            int bad = 0;
#pragma unroll
            for (int r = 0; r < EPT; ++r) {
                const int m = r * NT + tid;
                if (m + 1 < P) bad |= b.part[off + m] > b.part[off + m + 1];
            }
            if (__syncthreads_or(bad)) {
                uint64_t k2[EPT];
                uint32_t p2[EPT];
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    const int m = r * NT + tid;
                    const bool live = m < P;
                    k2[r] = live ? (uint64_t)((uint32_t)b.part[off + m] ^ 0x80000000u) : KEY_SENTINEL;
                    p2[r] = live ? (uint32_t)m : PAY_SENTINEL;
                }
                cta_sort_adaptive<EPT>(k2, p2, P, /*allow_trunc=*/false, xbuf, red);
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    const int q = tid * EPT + r;
                    const bool live = q < P;
                    if (live) tab[q] = (uint32_t)k2[r] ^ 0x80000000u;
                    key[r] = live ? ((uint64_t)lag_load<LAG_EITHER>(b, off + p2[r]) ^ LAG_KEY_MASK) : KEY_SENTINEL;
                    pay[r] = live ? (uint32_t)q : PAY_SENTINEL;
                }
                use_tab = true;
                __syncthreads();
            }
        }
        const bool keys_exact = cta_sort_adaptive<EPT>(key, pay, P, /*allow_trunc=*/!use_tab, xbuf, red);   // Main:228-235

        // out_partition: staged through shared memory so that the global store is coalesced
        __syncthreads();
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int q = tid * EPT + r;
            uint32_t v = 0;
            if (q < P) {
                if (b.part == nullptr) v = pay[r];
                else if (use_tab) v = tab[pay[r]];
                else v = (uint32_t)b.part[off + pay[r]];
            }
            xbuf[q] = v;
        }
        const bool rounds = (int64_t)P > C64;
        const bool need_keys = rounds || b.out_total != nullptr;
        if (need_keys) {
#pragma unroll
            for (int r = 0; r < EPT; ++r) {
                const int q = tid * EPT + r;
                if (q < P) skey[q] = keys_exact ? key[r] : ((uint64_t)lag_load<LAG_EITHER>(b, off + pay[r]) ^ LAG_KEY_MASK);
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int m = r * NT + tid;
            if (m < P) __stcs(b.out_part + off + m, (int32_t)xbuf[m]);
        }

        int64_t* tot = b.out_total ? b.out_total + b.tc_off[t] : nullptr;
        int32_t* cnt = b.out_count ? b.out_count + b.tc_off[t] : nullptr;
        auto lag_at = [&](int64_t k) { return (int64_t)(skey[k] ^ LAG_KEY_MASK); };
        auto put = [&](int64_t k, int32_t rank) { b.out_cons[off + k] = rank; };
        if (!rounds) {
            // P <= C: one partial round on all-zero totals: k-th sorted partition -> k-th member in id order
#pragma unroll
            for (int r = 0; r < EPT; ++r) {
                const int m = r * NT + tid;
                if (m < P) {
                    __stcs(b.out_cons + off + m, cset[m]);
                    if (tot) { tot[m] = lag_at(m); cnt[m] = 1; }
                }
            }
        } else {
            __syncthreads();      // the staging reads of xbuf are done before the fill reuses it
            const int C = (int)C64;
            if (C <= FILL_WARP_MAX_C) {
                if (tid < 32) warp_round_fill_dispatch(P, C, cset, lag_at, put, tot, cnt);
            } else if (C <= 512) {
                cta_round_fill_regs<2, NT>(P, C, cset, xbuf, lag_at, put, tot, cnt);
            } else if (C <= 1024) {
                if constexpr (EPT >= 4) cta_round_fill_regs<4, NT>(P, C, cset, xbuf, lag_at, put, tot, cnt);
            } else if (C <= 2048) {
                if constexpr (EPT >= 8) cta_round_fill_regs<8, NT>(P, C, cset, xbuf, lag_at, put, tot, cnt);
            } else {
                if constexpr (EPT >= 16) cta_round_fill_regs<16, NT>(P, C, cset, xbuf, lag_at, put, tot, cnt);
            }
        }
        __syncthreads();   // shared arrays are reused by the next topic
    }
}

}  // namespace lagassign
