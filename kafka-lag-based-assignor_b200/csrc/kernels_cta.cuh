// kernels_cta.cuh — topics whose partitions fit on chip: the whole of
// assignTopic (LagBasedPartitionAssignor.java:204-277) for one topic in one warp
// (P <= 256, registers) or one CTA (P <= 4096, shared memory). Lags are read from
// HBM once and the two result arrays written once: 16 B per partition (20 B with
// explicit partition ids).
#pragma once
#include "common.cuh"
#include "types.cuh"
#include "fill.cuh"
#include "sort_cta.cuh"
#include "sort_warp.cuh"

namespace lagassign {

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
