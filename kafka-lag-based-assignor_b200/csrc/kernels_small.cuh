// kernels_small.cuh — topics whose partitions fit on chip: the whole of
// assignTopic (LagBasedPartitionAssignor.java:204-277) for one topic in one warp
// (P <= 256, registers) or one CTA (P <= 4096, shared memory). Lags are read from
// HBM once and the two result arrays written once: 16 B per partition (20 B with
// explicit partition ids).
#pragma once
#include "common.cuh"
#include "fill.cuh"
#include "sort_cta.cuh"
#include "sort_warp.cuh"

namespace lagassign {

struct DevBatch {
    const int64_t* topic_off;   // [T+1]
    const int64_t* lag;         // [N]
    const int32_t* part;        // [N] or null
    const int32_t* topic_cset;  // [T]
    const int64_t* cset_off;    // [S+1]
    const int32_t* cset_rank;   // [R]
    const int64_t* tc_off;      // [T+1] offsets of the totals/counts blocks (or null)
    int32_t* out_part;          // [N]
    int32_t* out_cons;          // [N]
    int64_t* out_total;         // null unless enabled
    int32_t* out_count;
};

constexpr int SMALL_WARPS = 8;   // warps (= topics in flight) per CTA of k_warp_topics

struct WarpScratch {
    uint64_t lag_key[WARP_MAX_P];   // sorted keys, for the round fill / totals
    uint32_t out_a[WARP_MAX_P];     // out_partition staging (blocked -> coalesced)
    uint32_t out_b[WARP_MAX_P];     // out_consumer staging
    uint32_t part_tab[WARP_MAX_P];  // partition ids in ascending order (unordered-input path)
};

template <int EPT>
__device__ __forceinline__ void warp_topic(const DevBatch& b, int t, WarpScratch& ws) {
    const unsigned lane = lane_id();
    const int64_t off = b.topic_off[t];
    const int P = (int)(b.topic_off[t + 1] - off);
    const int s = b.topic_cset[t];
    const int64_t coff = b.cset_off[s];
    const int64_t C64 = b.cset_off[s + 1] - coff;
    const int32_t* __restrict__ cset = b.cset_rank + coff;

    uint64_t key[EPT];
    uint32_t pay[EPT];
    // Coalesced load; which register slot an element lands in does not matter,
    // the payload carries its list position (Main:228 sorts the whole list).
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        int m = r * 32 + (int)lane;
        bool live = m < P;
        key[r] = live ? ((uint64_t)ld_stream(b.lag + off + m) ^ LAG_KEY_MASK) : KEY_SENTINEL;
        pay[r] = live ? (uint32_t)m : PAY_SENTINEL;
    }

    bool use_tab = false;
    if (b.part != nullptr) {
        // Ties on lag break on partition id ascending (Main:230-231). A list that
        // is already ascending in partition id lets the position stand in for it;
        // otherwise order it by partition id first (stable by position).
        bool bad = false;
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            int m = r * 32 + (int)lane;
            if (m + 1 < P) bad |= b.part[off + m] > b.part[off + m + 1];
        }
        if (__any_sync(0xffffffffu, bad)) {
            uint64_t k2[EPT];
            uint32_t p2[EPT];
#pragma unroll
            for (int r = 0; r < EPT; ++r) {
                int m = r * 32 + (int)lane;
                bool live = m < P;
                k2[r] = live ? (uint64_t)((uint32_t)b.part[off + m] ^ 0x80000000u) : KEY_SENTINEL;
                p2[r] = live ? (uint32_t)m : PAY_SENTINEL;
            }
            warp_bitonic_sort<EPT>(k2, p2);
#pragma unroll
            for (int r = 0; r < EPT; ++r) {
                int i = (int)lane * EPT + r;
                bool live = i < P;
                if (live) ws.part_tab[i] = (uint32_t)k2[r] ^ 0x80000000u;
                key[r] = live ? ((uint64_t)b.lag[off + p2[r]] ^ LAG_KEY_MASK) : KEY_SENTINEL;
                pay[r] = live ? (uint32_t)i : PAY_SENTINEL;
            }
            use_tab = true;
            __syncwarp();
        }
    }

    warp_bitonic_sort<EPT>(key, pay);

    // out_partition, staged through shared memory so the global store is coalesced
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        int i = (int)lane * EPT + r;
        if (i < P) {
            uint32_t v;
            if (b.part == nullptr) v = pay[r];
            else if (use_tab) v = ws.part_tab[pay[r]];
            else v = (uint32_t)b.part[off + pay[r]];
            ws.out_a[i] = v;
        }
    }
    const bool rounds = (int64_t)P > C64;
    const bool need_keys = rounds || b.out_total != nullptr;
    if (need_keys) {
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            int i = (int)lane * EPT + r;
            if (i < P) ws.lag_key[i] = key[r];
        }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        int m = r * 32 + (int)lane;
        if (m < P) __stcs(b.out_part + off + m, (int32_t)ws.out_a[m]);
    }

    int64_t* tot = b.out_total ? b.out_total + b.tc_off[t] : nullptr;
    int32_t* cnt = b.out_count ? b.out_count + b.tc_off[t] : nullptr;
    if (!rounds) {
        // P <= C: one partial round on all-zero totals, so the k-th partition in
        // sorted order goes to the k-th member in id order.
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            int m = r * 32 + (int)lane;
            if (m < P) {
                __stcs(b.out_cons + off + m, cset[m]);
                if (tot) { tot[m] = (int64_t)(ws.lag_key[m] ^ LAG_KEY_MASK); cnt[m] = 1; }
            }
        }
        // members beyond P keep total 0 / count 0 (buffers are zero-filled)
    } else {
        const int C = (int)C64;
        auto lag_at = [&](int64_t k) { return (int64_t)(ws.lag_key[k] ^ LAG_KEY_MASK); };
        auto put = [&](int64_t k, int32_t rank) { ws.out_b[k] = (uint32_t)rank; };
        warp_round_fill_dispatch(P, C, cset, lag_at, put, tot, cnt);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            int m = r * 32 + (int)lane;
            if (m < P) __stcs(b.out_cons + off + m, (int32_t)ws.out_b[m]);
        }
    }
    __syncwarp();
}

// One warp per topic of `list` (topics with 1 <= P <= 256 and a non-empty consumer set).
__global__ void __launch_bounds__(SMALL_WARPS * 32)
k_warp_topics(DevBatch b, const int32_t* __restrict__ list, int n_list) {
    __shared__ WarpScratch scratch[SMALL_WARPS];
    const int warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * SMALL_WARPS + warp;
    if (idx >= n_list) return;
    const int t = list[idx];
    const int P = (int)(b.topic_off[t + 1] - b.topic_off[t]);
    WarpScratch& ws = scratch[warp];
    if (P <= 32) warp_topic<1>(b, t, ws);
    else if (P <= 64) warp_topic<2>(b, t, ws);
    else if (P <= 128) warp_topic<4>(b, t, ws);
    else warp_topic<8>(b, t, ws);
}

// ---- one CTA per topic, 256 < P <= 4096 ----------------------------------------------

constexpr int CTA_THREADS = 512;

// dynamic shared memory layout of k_cta_topics (sizes for n = CTA_MAX_P slots)
struct CtaSmemLayout {
    static constexpr size_t key_off = 0;
    static constexpr size_t pay_off = key_off + sizeof(uint64_t) * CTA_MAX_P;
    static constexpr size_t tab_off = pay_off + sizeof(uint32_t) * CTA_MAX_P;
    static constexpr size_t ckey_off = tab_off + sizeof(uint32_t) * CTA_MAX_P;
    static constexpr size_t cidx_off = ckey_off + sizeof(uint64_t) * CTA_MAX_P;
    static constexpr size_t total = cidx_off + sizeof(uint32_t) * CTA_MAX_P;
};

__device__ __forceinline__ int pow2_ceil(int x) {
    return x <= 1 ? 1 : 1 << (32 - __clz(x - 1));
}

__global__ void __launch_bounds__(CTA_THREADS)
k_cta_topics(DevBatch b, const int32_t* __restrict__ list, int n_list) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* key = reinterpret_cast<uint64_t*>(smem_raw + CtaSmemLayout::key_off);
    uint32_t* pay = reinterpret_cast<uint32_t*>(smem_raw + CtaSmemLayout::pay_off);
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem_raw + CtaSmemLayout::tab_off);
    uint64_t* ckey = reinterpret_cast<uint64_t*>(smem_raw + CtaSmemLayout::ckey_off);
    uint32_t* cidx = reinterpret_cast<uint32_t*>(smem_raw + CtaSmemLayout::cidx_off);

    const int tid = threadIdx.x, nt = blockDim.x;
    for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
        const int t = list[li];
        const int64_t off = b.topic_off[t];
        const int P = (int)(b.topic_off[t + 1] - off);
        const int s = b.topic_cset[t];
        const int64_t coff = b.cset_off[s];
        const int64_t C64 = b.cset_off[s + 1] - coff;
        const int32_t* __restrict__ cset = b.cset_rank + coff;
        const int n = max(64, pow2_ceil(P));

        bool use_tab = false;
        if (b.part != nullptr) {
            int bad = 0;
            for (int i = tid; i + 1 < P; i += nt) bad |= b.part[off + i] > b.part[off + i + 1];
            if (__syncthreads_or(bad)) {
                for (int i = tid; i < n; i += nt) {
                    bool live = i < P;
                    key[i] = live ? (uint64_t)((uint32_t)b.part[off + i] ^ 0x80000000u) : KEY_SENTINEL;
                    pay[i] = live ? (uint32_t)i : PAY_SENTINEL;
                }
                cta_bitonic_sort(key, pay, n);
                for (int i = tid; i < n; i += nt) {
                    bool live = i < P;
                    uint32_t orig = pay[i];
                    if (live) tab[i] = (uint32_t)key[i] ^ 0x80000000u;
                    key[i] = live ? ((uint64_t)b.lag[off + orig] ^ LAG_KEY_MASK) : KEY_SENTINEL;
                    pay[i] = live ? (uint32_t)i : PAY_SENTINEL;
                }
                use_tab = true;
            }
        }
        if (!use_tab) {
            for (int i = tid; i < n; i += nt) {
                bool live = i < P;
                key[i] = live ? ((uint64_t)ld_stream(b.lag + off + i) ^ LAG_KEY_MASK) : KEY_SENTINEL;
                pay[i] = live ? (uint32_t)i : PAY_SENTINEL;
            }
        }
        cta_bitonic_sort(key, pay, n);   // Main:228-235

        for (int i = tid; i < P; i += nt) {
            uint32_t v;
            if (b.part == nullptr) v = pay[i];
            else if (use_tab) v = tab[pay[i]];
            else v = (uint32_t)b.part[off + pay[i]];
            __stcs(b.out_part + off + i, (int32_t)v);
        }

        int64_t* tot = b.out_total ? b.out_total + b.tc_off[t] : nullptr;
        int32_t* cnt = b.out_count ? b.out_count + b.tc_off[t] : nullptr;
        auto lag_at = [&](int64_t k) { return (int64_t)(key[k] ^ LAG_KEY_MASK); };
        auto put = [&](int64_t k, int32_t rank) { b.out_cons[off + k] = rank; };
        if ((int64_t)P <= C64) {
            for (int i = tid; i < P; i += nt) {
                __stcs(b.out_cons + off + i, cset[i]);
                if (tot) { tot[i] = lag_at(i); cnt[i] = 1; }
            }
        } else if (C64 <= FILL_WARP_MAX_C) {
            if (tid < 32) warp_round_fill_dispatch(P, (int)C64, cset, lag_at, put, tot, cnt);
        } else {
            const int C = (int)C64;
            cta_round_fill<(CTA_MAX_P + CTA_THREADS - 1) / CTA_THREADS>(
                P, C, max(64, pow2_ceil(C)), cset, ckey, cidx, lag_at, put, tot, cnt);
        }
        __syncthreads();   // shared arrays are reused by the next topic
    }
}

// Topics with an empty consumer set are never visited by the reference (Main:177
// iterates consumersPerTopic; Main:211-213): partitions stay in input order,
// unassigned.
__global__ void __launch_bounds__(256)
k_unassigned_topics(DevBatch b, const int32_t* __restrict__ list, int n_list) {
    for (int li = blockIdx.y; li < n_list; li += gridDim.y) {
        const int t = list[li];
        const int64_t off = b.topic_off[t];
        const int64_t P = b.topic_off[t + 1] - off;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
            b.out_part[off + i] = b.part ? b.part[off + i] : (int32_t)i;
            b.out_cons[off + i] = -1;
        }
    }
}

}  // namespace lagassign
