// radix.cuh — device-wide path for topics too large for one CTA (P > 4096):
// a segmented, stable, least-significant-digit radix sort (8-bit digits, 64-bit
// key + 32-bit payload): one histogram kernel for all digit positions (digit
// bases and skip decisions), then per digit a counting kernel, two small scans
// and a scatter kernel, followed by the finalize / round-fill kernels.
//
// Segmentation: a tile never straddles topics, histograms and offset scans are
// per topic, and scatter destinations stay inside the topic's own range, so any
// number of large topics share one sequence of launches with no topic digit.
//
// Pass skipping is decided on the device per (topic, pass): if one digit value
// holds all P elements of a topic the pass cannot reorder anything and its tiles
// return at once; the per-topic source/destination buffer choice follows. With
// all-zero lags (BASELINE config 5) every pass is skipped and the sorted order is
// the input order, as the reference's stable sort gives (Main:228-235).
//
// Pass order (LSD): 4 passes over the partition id (only when the ids are given
// and not already ascending inside the topic: they are the tie-break, Main:231),
// then 8 passes over key = lag ^ 0x7FFF...F (descending signed lag, Main:234).
#pragma once
#include "common.cuh"
#include "types.cuh"
#include "fill.cuh"
#include <cstdint>
#include <cstdio>

namespace lagassign {

__device__ __forceinline__ uint32_t digit_of(uint64_t key, int lag_pass) {
    return (uint32_t)(key >> (8 * lag_pass)) & 0xFFu;
}

// ---- 1. are the partition ids ascending inside each large topic? -----------------------
__global__ void __launch_bounds__(256)
k_big_order_check(BigTopics bt, const int32_t* __restrict__ part) {
    for (int g = blockIdx.x; g < bt.n_tiles; g += gridDim.x) {
        const int b = bt.tile_bt[g];
        const int64_t P = bt.P[b];
        const int64_t t0 = (int64_t)(g - bt.tile0[b]) * RT_TILE;
        const int32_t* p = part + bt.off[b];
        int bad = 0;
        for (int i = threadIdx.x; i < RT_TILE; i += blockDim.x) {
            int64_t e = t0 + i;
            if (e + 1 < P) bad |= p[e] > p[e + 1];
        }
        if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(&bt.need_tie[b], 1);
    }
}

// ---- 2. value range of every topic: which digit passes can be skipped ----------------------
// A pass cannot reorder a topic whose elements all carry the same digit. That is decided from the
// bitwise OR and AND of the topic's keys (a digit is constant iff OR and AND agree on its eight bits),
// so this kernel is one streaming read of the lags (8 B/element) with no histogram: the digit bases of
// a pass come from that pass's own per-tile counts (4a-4c). An earlier version kept eight shared-memory
// histograms here (one atomic per element and digit): 0.47 ms for 2^26 elements, 57 % of BASELINE
// config 5's step. Also written here: the digit byte of the first lag pass (input of its counting
// kernel 4a) and, with raw offsets, the lags themselves (computePartitionLag, Main:376-404, fused).
//   range[b] = { OR of keys, OR of ~keys (= ~AND), OR of id keys, OR of ~id keys }   (zero-initialised)
template <bool HAS_PART>
__global__ void __launch_bounds__(RT_THREADS)
k_big_range(BigTopics bt, int64_t* __restrict__ lag, LagOffsets offs, const int32_t* __restrict__ part) {
    __shared__ unsigned long long red[RT_WARPS][4];
    // contiguous range of tiles per CTA so that a topic's words are touched by few atomics
    const int per = (bt.n_tiles + gridDim.x - 1) / gridDim.x;
    const int g0 = blockIdx.x * per, g1 = min(bt.n_tiles, g0 + per);
    const bool from_offsets = offs.end != nullptr;
    uint64_t k_or = 0, k_nor = 0, i_or = 0, i_nor = 0;
    int cur_b = -1;
    auto flush = [&](int b) {
        if (b < 0) return;
        const uint32_t v[8] = {(uint32_t)k_or, (uint32_t)(k_or >> 32), (uint32_t)k_nor, (uint32_t)(k_nor >> 32),
                               (uint32_t)i_or, (uint32_t)(i_or >> 32), (uint32_t)i_nor, (uint32_t)(i_nor >> 32)};
        uint32_t r[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q] = __reduce_or_sync(0xffffffffu, v[q]);
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) red[threadIdx.x >> 5][q] = ((unsigned long long)r[2 * q + 1] << 32) | r[2 * q];
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            unsigned long long x = 0;
#pragma unroll
            for (int w = 0; w < RT_WARPS; ++w) x |= red[w][threadIdx.x];
            if (x) atomicOr(reinterpret_cast<unsigned long long*>(bt.range) + (size_t)b * 4 + threadIdx.x, x);
        }
        __syncthreads();
        k_or = k_nor = i_or = i_nor = 0;
    };
    for (int g = g0; g < g1; ++g) {
        const int b = bt.tile_bt[g];
        if (b != cur_b) { flush(cur_b); cur_b = b; }
        const int64_t P = bt.P[b];
        const int64_t t0 = (int64_t)(g - bt.tile0[b]) * RT_TILE;
        int64_t* lg = lag + bt.off[b];
        const bool tie = HAS_PART && bt.need_tie[b] != 0;
        uint8_t* dg = bt.dig + bt.boff[b];
        // all of the tile's loads first (the stores below may alias them as far as the compiler knows)
        int64_t lv[RT_IPT];
#pragma unroll
        for (int k = 0; k < RT_IPT; ++k) {
            const int64_t e = t0 + k * RT_THREADS + threadIdx.x;
            lv[k] = 0;
            if (e < P) lv[k] = from_offsets ? lag_from_offsets(offs, bt.off[b] + e) : ld_stream(lg + e);
        }
#pragma unroll
        for (int k = 0; k < RT_IPT; ++k) {
            const int64_t e = t0 + k * RT_THREADS + threadIdx.x;
            if (e < P) {
                if (from_offsets) lg[e] = lv[k];
                const uint64_t key = (uint64_t)lv[k] ^ LAG_KEY_MASK;
                k_or |= key; k_nor |= ~key;
                dg[e] = (uint8_t)key;      // digit of the first lag pass, for its counting kernel (4a) when it is the topic's first pass
                if (tie) {
                    const uint64_t pk = (uint32_t)part[bt.off[b] + e] ^ 0x80000000u;
                    i_or |= pk; i_nor |= ~pk;
                }
            }
        }
    }
    flush(cur_b);
}

// ---- 3. per topic: skip decisions, buffer ping-pong, one record per (pass, tile) ----------------
__global__ void __launch_bounds__(256)
k_big_scan(BigTopics bt, int first_pass) {
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        const uint64_t k_or = bt.range[(size_t)b * 4 + 0], k_and = ~bt.range[(size_t)b * 4 + 1];
        const uint64_t i_or = bt.range[(size_t)b * 4 + 2], i_and = ~bt.range[(size_t)b * 4 + 3];
        const uint64_t k_diff = k_or ^ k_and;                 // bits in which the topic's keys differ
        const uint32_t i_diff = (uint32_t)(i_or ^ i_and);
        const bool tie = bt.need_tie[b] != 0;
        int cur = SEL_RAW, run = 0;
        for (int p = 0; p < N_PASS; ++p) {
            bool skip;
            if (p < N_TIE_PASS) skip = p < first_pass || !tie || ((i_diff >> (8 * p)) & 0xFFu) == 0;
            else skip = ((k_diff >> (8 * (p - N_TIE_PASS))) & 0xFFull) == 0;
            if (skip) {
                bt.sel_src[b * N_PASS + p] = SEL_SKIP;
                bt.sel_dst[b * N_PASS + p] = SEL_SKIP;
            } else {
                const int dst = (cur == SEL_A) ? SEL_B : SEL_A;
                bt.sel_src[b * N_PASS + p] = (int8_t)cur;
                bt.sel_dst[b * N_PASS + p] = (int8_t)dst;
                cur = dst;
                run += 1;
                atomicOr(&bt.pass_any[p], 1u);
            }
        }
        bt.sel_final[b] = (int8_t)cur;
        bt.all_equal[b] = (int8_t)(k_diff == 0);
        atomicMax(bt.passes_run, run);
        // the next executed pass after each pass (its digit byte is written by that pass's scatter)
        int nxt = -1;
        for (int p = N_PASS - 1; p >= 0; --p) {
            bt.nxt_pass[b * N_PASS + p] = (int8_t)nxt;
            if (bt.sel_src[b * N_PASS + p] != SEL_SKIP) nxt = p;
        }
    }
    __syncthreads();
    // What the pass kernels need to know about a tile in a pass, as ONE word next to the tile descriptor, so
    // that they do not chase tile -> topic -> three per-topic arrays with dependent loads:
    //   byte 0 = source buffer, byte 1 = destination buffer, byte 2 = the topic's next executed pass (0xFF: none)
    const int n_tiles_topic = (int)((bt.P[b] + RT_TILE - 1) / RT_TILE);
    const int tile0 = bt.tile0[b];
    for (int p = 0; p < N_PASS; ++p) {
        const uint32_t rec = (uint32_t)(uint8_t)bt.sel_src[b * N_PASS + p] | ((uint32_t)(uint8_t)bt.sel_dst[b * N_PASS + p] << 8) |
                             ((uint32_t)(uint8_t)bt.nxt_pass[b * N_PASS + p] << 16);
        for (int k = tid; k < n_tiles_topic; k += blockDim.x) bt.pass_rec[(size_t)p * bt.n_tiles + tile0 + k] = rec;
    }
}

// ---- 4. one digit pass = four kernels, no inter-CTA waiting ------------------------------
//   4a k_big_tile_hist   digit counts of every tile          (reads 1 B/element: the digit bytes the previous
//                                                             pass left behind; 8 B/element in a topic's first pass)
//   4b k_big_seg_scan    exclusive scan over the tiles of a segment (RT_SEG tiles), per digit
//   4c k_big_topic_scan  exclusive scan over a topic's segments, seeded with the digit base
//   4d k_big_scatter     rank inside the tile, scatter       (reads 12, writes 12 + 1 B/element)
// An earlier version of this pass was a single "onesweep" kernel whose tiles obtained
// their offsets by decoupled look-back (one- and two-level, windowed). On B200 it stayed
// near 25 % of HBM peak regardless of tile shape, prefetch depth or look-back window:
// with ~450 tiles in flight, per-tile cycle counters showed a third of the time spent
// waiting for predecessor tiles (the chain forces completion in tile order, so every
// tile runs at the pace of the slowest CTA ahead of it). A separate counting kernel removes
// every dependency between CTAs; so that it does not have to read the 8-byte keys again,
// the scatter of pass p also writes, next to every key, the ONE byte the counting kernel
// of the topic's next executed pass needs (BigTopics::dig).
#ifndef LAGASSIGN_RT_STAGES
#define LAGASSIGN_RT_STAGES 2
#endif
constexpr int RT_STAGES = LAGASSIGN_RT_STAGES;   // tiles in flight per CTA (the one being processed included)
static_assert(RT_STAGES >= 2 && RT_STAGES <= 4, "ring depth");
static_assert(RT_THREADS == 256, "the pass kernels map one thread to one digit value");

// digit of an element in `pass`: passes below N_TIE_PASS sort by partition id (pv = the id),
// the others by the key
__device__ __forceinline__ uint32_t pass_digit(int pass, uint64_t key, uint32_t pv) {
    return pass < N_TIE_PASS ? (((pv ^ 0x80000000u) >> (8 * pass)) & 0xFFu) : digit_of(key, pass - N_TIE_PASS);
}

// ---- bulk asynchronous copies (TMA, cp.async.bulk: SASS UBLKCP) completing on an mbarrier ----------
#ifdef LAGASSIGN_HOST_EMULATION   // tests/emu: the copy happens at issue; the barrier word = {completed phases, pending bytes}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
__device__ __forceinline__ void fence_mbar_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    __atomic_store_n(reinterpret_cast<uint32_t*>(bar) + 1, bytes, __ATOMIC_RELAXED);
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    emu::check_aligned(smem_dst, 16); emu::check_aligned(gsrc, 16);
    if (bytes % 16 != 0) { std::fprintf(stderr, "cuda_emu: bulk copy of %u bytes\n", bytes); std::abort(); }
    std::memcpy(smem_dst, gsrc, bytes);
    uint32_t* w = reinterpret_cast<uint32_t*>(bar);
    if (__atomic_sub_fetch(w + 1, bytes, __ATOMIC_RELAXED) == 0) __atomic_add_fetch(w, 1u, __ATOMIC_RELEASE);   // phase complete
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(bar);
    while ((__atomic_load_n(w, __ATOMIC_ACQUIRE) & 1u) == parity) std::this_thread::yield();
}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (TMA) ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 16-byte aligned source, destination and size
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
#endif

// 4a ------------------------------------------------------------------------------------------
// One CTA per tile, grid-stride. What bounds this kernel is the chain of dependent loads per tile
// (descriptor -> buffer choice -> data), so the next tile's descriptor and digit bytes are fetched
// while the current tile is counted (warp-private counters, shared-memory atomics).
template <bool HAS_PART>
__global__ void __launch_bounds__(RT_THREADS)
k_big_tile_hist(BigTopics bt, int pass, const int64_t* __restrict__ lag, const int32_t* __restrict__ part) {
    __shared__ uint32_t whist[RT_WARPS][256];
    const int tid = threadIdx.x, warp = tid >> 5;
    if (bt.pass_any[pass] == 0) return;
    static_assert(RT_IPT == 16, "one uint4 of digit bytes per thread");
    struct T { int src, n_valid; int64_t buf_off, raw_off; uint4 v; };
    auto fetch = [&](int g) {
        T t;
        t.src = SEL_SKIP; t.n_valid = 0; t.buf_off = 0; t.raw_off = 0; t.v = make_uint4(0, 0, 0, 0);
        if (g < bt.n_tiles) {
            const int4 a = __ldg(bt.tile_desc + 2 * (size_t)g), c = __ldg(bt.tile_desc + 2 * (size_t)g + 1);
            t.n_valid = a.z;
            t.buf_off = ((int64_t)(uint32_t)c.y << 32) | (uint32_t)c.x;
            t.raw_off = ((int64_t)(uint32_t)c.w << 32) | (uint32_t)c.z;
            t.src = (int)(int8_t)(__ldg(bt.pass_rec + (size_t)pass * bt.n_tiles + g) & 0xFFu);
            // the digit bytes are only meaningful when src is a sort buffer; reading them anyway keeps the load
            // independent of the buffer choice (the buffer is allocated and padded for every tile)
            if (tid * 16 < t.n_valid) t.v = *reinterpret_cast<const uint4*>(bt.dig + t.buf_off + tid * 16);
        }
        return t;
    };
#pragma unroll
    for (int w = 0; w < RT_WARPS; ++w) whist[w][tid] = 0;
    T cur = fetch((int)blockIdx.x);
    __syncthreads();
    for (int g = blockIdx.x; g < bt.n_tiles; g += gridDim.x) {
        const T nxt = fetch(g + (int)gridDim.x);
        if (cur.src != SEL_SKIP) {                 // uniform over the CTA
            const int n_valid = cur.n_valid;
            if (cur.src != SEL_RAW || pass == N_TIE_PASS) {
                // the scatter of the topic's previous pass left this pass's digit of every element in bt.dig
                // (for a topic whose first pass is the first lag pass, k_big_range did)
                const int i0 = tid * 16;
                const uint32_t wds[4] = {cur.v.x, cur.v.y, cur.v.z, cur.v.w};
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (i0 + q < n_valid) atomicAdd(&whist[warp][(wds[q >> 2] >> (8 * (q & 3))) & 0xFFu], 1u);
            } else {
                // the topic's first executed pass: digits from the caller's arrays
                const int64_t* lg = lag + cur.raw_off;
                const int32_t* pt = HAS_PART ? part + cur.raw_off : nullptr;
#pragma unroll
                for (int k = 0; k < RT_IPT; ++k) {
                    const int i = k * RT_THREADS + tid;
                    if (i < n_valid) {
                        const uint32_t pv = (HAS_PART && pass < N_TIE_PASS) ? (uint32_t)pt[i] : 0u;
                        const uint64_t key = (!HAS_PART || pass >= N_TIE_PASS) ? ((uint64_t)ld_stream(lg + i) ^ LAG_KEY_MASK) : 0ull;
                        atomicAdd(&whist[warp][pass_digit(pass, key, pv)], 1u);
                    }
                }
            }
            __syncthreads();
            uint32_t c = 0;
#pragma unroll
            for (int w = 0; w < RT_WARPS; ++w) { c += whist[w][tid]; whist[w][tid] = 0; }   // read and clear for the next tile
            bt.tile_hist[(size_t)g * 256 + tid] = c;
            __syncthreads();
        }
        cur = nxt;
    }
}

// 4b: one CTA per segment; tile_hist becomes the exclusive prefix inside the segment ------------
__global__ void __launch_bounds__(256)
k_big_seg_scan(BigTopics bt, int pass) {
    const int seg = blockIdx.x, tid = threadIdx.x;
    if (bt.pass_any[pass] == 0) return;
    const int b = bt.seg_bt[seg];
    if (bt.sel_src[b * N_PASS + pass] == SEL_SKIP) return;
    const int n_tiles_topic = (int)((bt.P[b] + RT_TILE - 1) / RT_TILE);
    const int first_in_topic = (seg - bt.seg0[b]) * RT_SEG;
    const int n = min(RT_SEG, n_tiles_topic - first_in_topic);
    uint32_t* h = bt.tile_hist + (size_t)(bt.tile0[b] + first_in_topic) * 256 + tid;
    uint32_t run = 0;
    for (int t = 0; t < n; t += 8) {
        uint32_t v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (t + q < n) ? h[(size_t)(t + q) * 256] : 0u;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (t + q < n) h[(size_t)(t + q) * 256] = run;
            run += v[q];
        }
    }
    bt.seg_total[(size_t)seg * 256 + tid] = run;
}

// 4c: one CTA per topic; seg_total becomes the global start of (segment, digit) in the topic ----
// (digit bases = exclusive scan over the digit values of the topic's totals, computed here from the
// segment totals of this very pass)
__global__ void __launch_bounds__(256)
k_big_topic_scan(BigTopics bt, int pass) {
    __shared__ uint32_t wsum[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (bt.pass_any[pass] == 0) return;
    if (bt.sel_src[b * N_PASS + pass] == SEL_SKIP) return;
    const int n_tiles_topic = (int)((bt.P[b] + RT_TILE - 1) / RT_TILE);
    const int n = (n_tiles_topic + RT_SEG - 1) / RT_SEG;
    uint32_t* s = bt.seg_total + (size_t)bt.seg0[b] * 256 + tid;
    uint32_t total = 0;
    for (int t = 0; t < n; t += 8) {
        uint32_t v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (t + q < n) ? s[(size_t)(t + q) * 256] : 0u;
#pragma unroll
        for (int q = 0; q < 8; ++q) total += v[q];
    }
    uint32_t x = total;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    uint32_t add = 0;
    for (int w = 0; w < warp; ++w) add += wsum[w];
    uint32_t run = x - total + add;              // first position of digit value `tid` in the sorted topic
    for (int t = 0; t < n; t += 8) {
        uint32_t v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (t + q < n) ? s[(size_t)(t + q) * 256] : 0u;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (t + q < n) s[(size_t)(t + q) * 256] = run;
            run += v[q];
        }
    }
}

// 4d ------------------------------------------------------------------------------------------
// CTAs of SC_THREADS threads walk the tiles with a grid stride. One elected thread fetches a tile
// with two bulk asynchronous copies (TMA; keys and payloads) into a ring of RT_STAGES shared-memory
// stages, completion on an mbarrier per stage: HBM loads stay in flight while the current tile is
// ranked and scattered and no thread spends issue slots on the copy. The stage being processed
// doubles as the staging buffer of the tile-local scatter and is refilled as soon as the sorted
// tile has been read back out of it.
//
// Ranking (stable: elements with equal digit keep memory order). A warp owns a contiguous run of
// the tile, item k of lane l being element warp*32*IPT + k*32 + l, and keeps running digit counts
// in its own row of shared-memory counters.
//  * Fast form (EXACT = false): rank = atomicAdd(&count[warp][digit], 1), ONE instruction per
//    element (ATOMS.ADD, 1.7 SM-cycles per warp instruction measured on B200). Atomicity makes the
//    ranks of a digit a permutation of 0..n-1 whatever happens; that lanes of one instruction that
//    hit the same counter are served in LANE order — which is what stability needs — is how the
//    B200 behaves (0 violations in 2e8 probes, scripts/ubench) but is not architecturally promised.
//    So the sort is CERTIFIED afterwards: k_big_verify checks that the final sequence is strictly
//    increasing in (key, tie-break, position), which together with the permutation property proves
//    the result exact; if it ever is not, lagassign_execute() repeats the sort with the exact form.
//  * Exact form (EXACT = true): the lanes with the same digit are found with eight ballots (one per
//    digit bit: 24 SM-cycles per warp item; match.any costs 64), the lowest of them adds their
//    number to the counter with one atomic whose return value, broadcast by a shuffle, is the
//    group's base. Atomics of one warp on one address complete in program order, so items need no
//    barrier between them.
#ifndef LAGASSIGN_SC_THREADS
#define LAGASSIGN_SC_THREADS 256
#endif
constexpr int SC_THREADS = LAGASSIGN_SC_THREADS;
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr int SC_IPT = RT_TILE / SC_THREADS;
static_assert(SC_THREADS >= 256 && SC_THREADS * SC_IPT == RT_TILE, "a tile is SC_THREADS x SC_IPT elements; threads 0..255 also own one digit value each");
static_assert(RT_TILE < 65536 && SC_WARPS % 2 == 0, "tile-local offsets are packed two per word");

struct PassSmem {
    alignas(16) uint64_t key[RT_STAGES][RT_TILE + 2];   // + 2: a raw source may start 8 bytes after a 16-byte boundary
    alignas(16) uint32_t pay[RT_STAGES][RT_TILE + 4];
    uint32_t whist[SC_WARPS / 2][256];   // per-warp running digit counts, then tile-local offsets per (warp, digit):
                                         // 16 bits each (<= RT_TILE), warps 2j and 2j+1 in the low / high half of word j
    int32_t dst_base[256];           // index in the topic of the tile's first element of each digit, minus its tile-local offset
    uint32_t wsum[8];
    alignas(8) uint64_t full_bar[RT_STAGES];
};

#ifdef LAGASSIGN_RT_TIMING
#define RT_TICK(slot) do { if (tid == 0) { long long _t = clock64(); t_acc[slot] += _t - t_last; t_last = _t; } } while (0)
#else
#define RT_TICK(slot) do { } while (0)
#endif

__device__ __forceinline__ uint32_t peers_of_digit(uint32_t dig) {
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const bool set = (dig >> bit) & 1u;
        const uint32_t m = __ballot_sync(0xffffffffu, set);
        peers &= set ? m : ~m;
    }
    return peers;
}

template <bool HAS_PART, bool EXACT>
__global__ void __launch_bounds__(SC_THREADS, LAGASSIGN_RT_MINBLOCKS)
k_big_scatter(BigTopics bt, int pass, const int64_t* __restrict__ lag, const int32_t* __restrict__ part) {
#ifdef LAGASSIGN_HOST_EMULATION
    unsigned char* smem_raw = emu::cta().dyn_smem;
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    PassSmem& sm = *reinterpret_cast<PassSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (bt.pass_any[pass] == 0) return;   // every topic's digit is constant: nothing to move
    const int n_tiles = bt.n_tiles;
#ifdef LAGASSIGN_RT_TIMING
    long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_last = clock64();
    int t_tiles = 0;
#endif

    // A tile's geometry is one 32-byte record; the record also remembers which buffer the
    // topic reads in this pass (-1: pass skipped for the topic, or no such tile).
    struct Tile { TileDesc d; int src, dst, nxt_pass; };
    auto load_tile = [&](int g) {
        Tile t;
        t.src = SEL_SKIP;
        t.dst = SEL_SKIP;
        t.nxt_pass = -1;
        t.d.b = -1;
        if (g < n_tiles) {
            const int4 a = __ldg(bt.tile_desc + 2 * (size_t)g), c = __ldg(bt.tile_desc + 2 * (size_t)g + 1);
            t.d.b = a.x; t.d.tile_in_topic = a.y; t.d.n_valid = a.z; t.d.seg = a.w;
            t.d.buf_off = ((int64_t)(uint32_t)c.y << 32) | (uint32_t)c.x;
            t.d.raw_off = ((int64_t)(uint32_t)c.w << 32) | (uint32_t)c.z;
            t.src = bt.sel_src[t.d.b * N_PASS + pass];
            t.dst = bt.sel_dst[t.d.b * N_PASS + pass];
            t.nxt_pass = bt.nxt_pass[t.d.b * N_PASS + pass];   // the topic's next executed pass (-1: none)
        }
        return t;
    };
    // A raw source (the caller's lag array) is only 8-byte aligned: the copy then starts one element early.
    auto raw_shift = [&](const Tile& t) -> int {
        return (int)((reinterpret_cast<uintptr_t>(lag + t.d.raw_off) >> 3) & 1u);
    };
    // one thread: fetch a tile into ring stage `stage`
    auto issue_loads = [&](int stage, const Tile& t) {
        if (t.src == SEL_SKIP) return;
        uint64_t* bar = &sm.full_bar[stage];
        if (t.src == SEL_RAW) {
            const int sh = raw_shift(t);
            const uint32_t nb = (uint32_t)(((t.d.n_valid + sh) * 8 + 15) & ~15);
            mbar_expect_tx(bar, nb);
            bulk_load(&sm.key[stage][0], lag + t.d.raw_off - sh, nb, bar);
        } else {
            const uint64_t* skey = ((t.src == SEL_A) ? bt.keyA : bt.keyB) + t.d.buf_off;   // buf_off is a multiple of 16 elements
            const uint32_t* spay = ((t.src == SEL_A) ? bt.payA : bt.payB) + t.d.buf_off;
            const uint32_t nbk = (uint32_t)((t.d.n_valid * 8 + 15) & ~15), nbp = (uint32_t)((t.d.n_valid * 4 + 15) & ~15);
            mbar_expect_tx(bar, nbk + nbp);
            bulk_load(&sm.key[stage][0], skey, nbk, bar);
            bulk_load(&sm.pay[stage][0], spay, nbp, bar);
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < RT_STAGES; ++j) mbar_init(&sm.full_bar[j], 1);
        fence_mbar_init();
    }
#pragma unroll
    for (int w = 0; w < SC_WARPS / 2; ++w)
        if (tid < 256) sm.whist[w][tid] = 0;
    __syncthreads();
    const int G = (int)gridDim.x;
    uint32_t* const my_cnt = sm.whist[warp >> 1];      // this warp's counters: the (warp & 1) half of these words
    const int my_sh = (warp & 1) * 16;
    const int g0 = (int)blockIdx.x;
    // q[j] = the tile of iteration it + j, already being fetched into ring stage (it + j) % RT_STAGES; `far` is the
    // tile after those: it is fetched into the stage of the current tile as soon as every thread has read the
    // tile-local scatter back out of it, BEFORE the global stores are issued — so a warp that is held up by the
    // store queue does not hold up the refill, and warps flow from their stores straight into the next tile
    // (whose first barrier comes after its ranking).
    Tile q[RT_STAGES];
#pragma unroll
    for (int j = 0; j < RT_STAGES; ++j) q[j] = load_tile(g0 + j * G);
    if (tid == 0) {
#pragma unroll
        for (int j = 0; j < RT_STAGES; ++j) issue_loads(j, q[j]);
    }
    Tile far = load_tile(g0 + RT_STAGES * G);
    uint32_t phase = 0u;                   // bit j: parity of the next completion of stage j's barrier
    int s = 0;                             // ring stage of the tile being processed

    for (int it = 0;; ++it) {
        const int g = g0 + it * G;
        if (g >= n_tiles) break;
        const Tile cur = q[0];
#pragma unroll
        for (int j = 0; j + 1 < RT_STAGES; ++j) q[j] = q[j + 1];
        q[RT_STAGES - 1] = far;                                    // tile it + RT_STAGES: next occupant of this stage
        far = load_tile(g + (RT_STAGES + 1) * G);
        const Tile refill = q[RT_STAGES - 1];
        RT_TICK(0);
        const int s_cur = s;
        s = (s + 1 == RT_STAGES) ? 0 : s + 1;
        if (cur.src == SEL_SKIP) {           // this topic's digit is constant in this pass: nothing was fetched,
            if (tid == 0) issue_loads(s_cur, refill);   // the stage is free (its last reader passed a barrier long ago)
            continue;
        }

        const int src = cur.src;
        const int dst = cur.dst;
        const int nxt_pass = cur.nxt_pass;
        const int64_t t0 = (int64_t)cur.d.tile_in_topic * RT_TILE;
        const int n_valid = cur.d.n_valid;
        const int64_t boff = cur.d.buf_off - t0;
        uint64_t* dkey = (dst == SEL_A) ? bt.keyA + boff : bt.keyB + boff;
        uint32_t* dpay = (dst == SEL_A) ? bt.payA + boff : bt.payB + boff;
        uint8_t* ddig = bt.dig + boff;
        const int32_t* pt = HAS_PART ? part + (cur.d.raw_off - t0) : nullptr;
        const int sh = (src == SEL_RAW) ? raw_shift(cur) : 0;
        // where this tile's elements of digit `tid` start in the topic (4b + 4c); used after the ranking
        uint32_t base_d = 0;
        if (tid < 256) base_d = bt.tile_hist[(size_t)g * 256 + tid] + bt.seg_total[(size_t)cur.d.seg * 256 + tid];
        RT_TICK(1);

        // this thread's elements: item k of lane l in warp w is tile element w*32*IPT + k*32 + l
        mbar_wait(&sm.full_bar[s_cur], (phase >> s_cur) & 1u);
        phase ^= 1u << s_cur;
        uint64_t key[SC_IPT];
        uint32_t pay[SC_IPT];
        uint32_t dig[SC_IPT];
#pragma unroll
        for (int k = 0; k < SC_IPT; ++k) {
            const int i = warp * (32 * SC_IPT) + k * 32 + lane;
            if (i < n_valid) {
                if (src == SEL_RAW) { key[k] = sm.key[s_cur][i + sh] ^ LAG_KEY_MASK; pay[k] = (uint32_t)(t0 + i); }
                else { key[k] = sm.key[s_cur][i]; pay[k] = sm.pay[s_cur][i]; }
                dig[k] = pass_digit(pass, key[k], (HAS_PART && pass < N_TIE_PASS) ? (uint32_t)pt[pay[k]] : 0u);
            } else {
                key[k] = KEY_SENTINEL; pay[k] = PAY_SENTINEL; dig[k] = 255u;  // padding: neither counted nor moved
            }
        }
        RT_TICK(2);

        // rank inside the warp (see above)
        uint32_t rank[SC_IPT];
#pragma unroll
        for (int k = 0; k < SC_IPT; ++k) {
            // Padding (a topic's last tile) takes no part: valid elements then get the tile positions 0..n_valid-1
            // by atomicity alone, whatever the order in which same-digit lanes are served.
            const bool valid = warp * (32 * SC_IPT) + k * 32 + lane < n_valid;
            if constexpr (!EXACT) {
                rank[k] = 0;
                if (valid) rank[k] = (atomicAdd(&my_cnt[dig[k]], 1u << my_sh) >> my_sh) & 0xFFFFu;
            } else {
                const uint32_t peers = peers_of_digit(dig[k]) & __ballot_sync(0xffffffffu, valid);
                const uint32_t below = peers & ((1u << lane) - 1u);
                uint32_t pre_cnt = 0;
                if (valid && below == 0) pre_cnt = (atomicAdd(&my_cnt[dig[k]], (uint32_t)__popc(peers) << my_sh) >> my_sh) & 0xFFFFu;
                pre_cnt = __shfl_sync(0xffffffffu, pre_cnt, (__ffs(peers) - 1) & 31);
                rank[k] = pre_cnt + __popc(below);
            }
        }
        __syncthreads();   // all counts in; every thread holds its elements in registers (the stage may be overwritten)
        RT_TICK(3);

        // thread d owns digit value d: offsets over warps and over digits inside the tile
        {
            uint32_t cnt = 0, x = 0;
            if (tid < 256) {
#pragma unroll
                for (int w = 0; w < SC_WARPS / 2; ++w) {
                    const uint32_t v = sm.whist[w][tid];          // counts of warps 2w (low half) and 2w+1 (high half)
                    const uint32_t lo = v & 0xFFFFu, hi = v >> 16;
                    sm.whist[w][tid] = cnt | ((cnt + lo) << 16);   // their exclusive offsets inside the digit
                    cnt += lo + hi;
                }
                x = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= d) x += y;
                }
                if (lane == 31) sm.wsum[warp] = x;
            }
            __syncthreads();
            if (tid < 256) {
                uint32_t add = 0;
                for (int w = 0; w < warp; ++w) add += sm.wsum[w];
                const uint32_t excl = x - cnt + add;
#pragma unroll
                for (int w = 0; w < SC_WARPS / 2; ++w) sm.whist[w][tid] += excl * 0x10001u;   // both halves stay below RT_TILE
                sm.dst_base[tid] = (int32_t)(base_d - excl);
            }
        }
        __syncthreads();
        RT_TICK(4);

        // scatter inside the tile (the consumed ring stage is the staging buffer) ...
#pragma unroll
        for (int k = 0; k < SC_IPT; ++k) {
            if (warp * (32 * SC_IPT) + k * 32 + lane < n_valid) {
                const uint32_t pos = ((my_cnt[dig[k]] >> my_sh) & 0xFFFFu) + rank[k];
                sm.key[s_cur][pos] = key[k];
                sm.pay[s_cur][pos] = pay[k];
            }
        }
        __syncthreads();
        RT_TICK(5);
        // (the counters are free again: zero them for the next tile; the barrier below orders it)
#pragma unroll
        for (int w = 0; w < SC_WARPS / 2; ++w)
            if (tid < 256) sm.whist[w][tid] = 0;
        // ... then out in digit runs (coalesced), with the digit byte of the topic's next pass beside the key
        {
            uint64_t kk[SC_IPT];
            uint32_t pp[SC_IPT];
            int32_t oo[SC_IPT];
            uint32_t pv[SC_IPT];
#pragma unroll
            for (int k = 0; k < SC_IPT; ++k) {       // all shared-memory reads first, then the stores
                const int i = k * SC_THREADS + tid;
                kk[k] = sm.key[s_cur][i];
                pp[k] = sm.pay[s_cur][i];
            }
            fence_proxy_async();   // this thread's reads/writes of the stage before the bulk copy into it
            __syncthreads();       // everyone holds its part of the sorted tile in registers, counters are zeroed
            if (tid == 0) issue_loads(s_cur, refill);
            RT_TICK(6);
#pragma unroll
            for (int k = 0; k < SC_IPT; ++k) {
                const int i = k * SC_THREADS + tid;
                pv[k] = 0;
                if (HAS_PART && (pass < N_TIE_PASS || (nxt_pass >= 0 && nxt_pass < N_TIE_PASS)) && i < n_valid) pv[k] = (uint32_t)pt[pp[k]];
                const uint32_t d = (i < n_valid) ? pass_digit(pass, kk[k], pv[k]) : 255u;
                oo[k] = sm.dst_base[d] + i;
            }
#pragma unroll
            for (int k = 0; k < SC_IPT; ++k) {
                const int i = k * SC_THREADS + tid;
                if (i < n_valid) {
                    dkey[oo[k]] = kk[k];      // (streaming stores, st.global.cs, measured: no difference)
                    dpay[oo[k]] = pp[k];
                    if (nxt_pass >= 0) ddig[oo[k]] = (uint8_t)pass_digit(nxt_pass, kk[k], pv[k]);
                }
            }
        }
        RT_TICK(7);
#ifdef LAGASSIGN_RT_TIMING
        ++t_tiles;
#endif
    }
#ifdef LAGASSIGN_RT_TIMING
    if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 200) && t_tiles > 0)
        printf("pass %d cta %d tiles %d | cycles per tile: top %lld meta %lld wait+lds %lld rank+bar %lld scan+bar %lld local-scatter+bar %lld readback+bar %lld stores %lld\n",
               pass, (int)blockIdx.x, t_tiles, t_acc[0] / t_tiles, t_acc[1] / t_tiles, t_acc[2] / t_tiles, t_acc[3] / t_tiles,
               t_acc[4] / t_tiles, t_acc[5] / t_tiles, t_acc[6] / t_tiles, t_acc[7] / t_tiles);
#endif
}

// ---- 4e. certificate of the whole sort (see the ranking note above) --------------------------
// The sorted sequence of a topic must be strictly increasing in (key, partition id when the ids
// had to be ordered, position). Ranks are a permutation by atomicity, so this proves the result.
__global__ void __launch_bounds__(256)
k_big_verify(BigTopics bt, const int32_t* __restrict__ part) {
    int bad = 0;
    for (int g = blockIdx.x; g < bt.n_tiles; g += gridDim.x) {
        const int b = bt.tile_bt[g];
        const int sel = bt.sel_final[b];
        if (sel == SEL_RAW) continue;          // no pass ran: the input order is the result
        const int64_t P = bt.P[b], boff = bt.boff[b];
        const uint64_t* fkey = (sel == SEL_A) ? bt.keyA + boff : bt.keyB + boff;
        const uint32_t* fpay = (sel == SEL_A) ? bt.payA + boff : bt.payB + boff;
        const int32_t* pt = (part != nullptr && bt.need_tie[b] != 0) ? part + bt.off[b] : nullptr;
        const int64_t t0 = (int64_t)(g - bt.tile0[b]) * RT_TILE;
        for (int i = threadIdx.x; i < RT_TILE; i += blockDim.x) {
            const int64_t e = t0 + i;
            if (e + 1 >= P) break;
            const uint64_t k0 = fkey[e], k1 = fkey[e + 1];
            const uint32_t p0 = fpay[e], p1 = fpay[e + 1];
            bool ok = k0 < k1;
            if (k0 == k1) {
                if (pt) {
                    const int32_t i0 = pt[p0], i1 = pt[p1];
                    ok = i0 < i1 || (i0 == i1 && p0 < p1);
                } else {
                    ok = p0 < p1;
                }
            }
            bad |= !ok;
        }
    }
    if (bad) atomicOr(bt.verify_fail, 1u);
}

// ---- 5. finalize: out_partition (and out_consumer for single-round topics) ---------------
// VERIFY: also the certificate 4e, on the same read of the sorted sequence (the separate kernel costs
// another 12 B/element; it remains for the sort-only mode of the grouping step, which has no finalize).
template <bool VERIFY>
__global__ void __launch_bounds__(256)
k_big_finalize(BigTopics bt, const int64_t* __restrict__ lag, const int32_t* __restrict__ part,
               const int32_t* __restrict__ topic_cset, const int64_t* __restrict__ cset_off,
               const int32_t* __restrict__ cset_rank, const int64_t* __restrict__ tc_off,
               int32_t* __restrict__ out_part, int32_t* __restrict__ out_cons,
               int64_t* __restrict__ out_total, int32_t* __restrict__ out_count) {
    int bad = 0;
    // What a tile needs to know about its topic sits behind three levels of dependent loads (tile -> topic ->
    // consumer set): ~2 us of latency against < 1 us of copying per tile. The next tile's are fetched while the
    // current tile is processed (BASELINE config 5: this kernel 0.22 -> see profiles/).
    struct Meta { int b, t, sel, tile0; int64_t P, off, boff, coff, C; int8_t all_equal; };
    auto load_meta = [&](int g) {
        Meta m;
        m.b = -1; m.t = 0; m.sel = SEL_RAW; m.tile0 = 0; m.P = 0; m.off = 0; m.boff = 0; m.coff = 0; m.C = 0; m.all_equal = 0;
        if (g < bt.n_tiles) {
            m.b = bt.tile_bt[g];
            m.t = bt.topic[m.b];
            m.P = bt.P[m.b]; m.off = bt.off[m.b]; m.boff = bt.boff[m.b];
            m.sel = bt.sel_final[m.b];
            m.all_equal = bt.all_equal[m.b];
            m.tile0 = bt.tile0[m.b];
            const int s = topic_cset[m.t];
            m.coff = cset_off[s];
            m.C = cset_off[s + 1] - m.coff;
        }
        return m;
    };
    Meta nxt = load_meta((int)blockIdx.x);
    for (int g = blockIdx.x; g < bt.n_tiles; g += gridDim.x) {
        const Meta cur = nxt;
        nxt = load_meta(g + (int)gridDim.x);
        const int b = cur.b;
        const int t = cur.t;
        const int64_t P = cur.P, off = cur.off, boff = cur.boff;
        const int sel = cur.sel;
        const uint64_t* fkey = (sel == SEL_A) ? bt.keyA + boff : bt.keyB + boff;
        const uint32_t* fpay = (sel == SEL_A) ? bt.payA + boff : bt.payB + boff;
        const int64_t coff = cur.coff, C = cur.C;
        const bool single_round = P <= C;
        // All lags equal (an idle topic: zero everywhere): every round starts with equal
        // totals, so its picks are the members in id order and the sorted partition k goes to
        // member k mod C. No sequential fill is needed.
        const bool cyclic = !single_round && cur.all_equal != 0;
        const int64_t t0 = (int64_t)(g - cur.tile0) * RT_TILE;
        const bool sorted = sel != SEL_RAW;
        const bool tie = part != nullptr && bt.need_tie[b] != 0;
        const bool need_key = sorted && (VERIFY || (single_round && out_total != nullptr));
        const int n_valid = (int)min((int64_t)RT_TILE, P - t0);
        const unsigned lane = lane_id();
        // Topics that needed no pass (idle topics: every lag equal; BASELINE config 5) only GENERATE their result:
        // position -> itself, member -> position mod C. Kept apart from the general code below, whose 64-bit index
        // arithmetic and per-element case distinctions cost 64 instructions per element there (ncu) — more than the
        // 8 bytes per element this path has to write.
        if (!sorted && part == nullptr && out_total == nullptr && (single_round || cyclic)) {
            int32_t* __restrict__ op = out_part + off + t0;
            int32_t* __restrict__ oc = out_cons + off + t0;
            const int32_t* __restrict__ cs = cset_rank + coff;
            const uint32_t e0 = (uint32_t)t0;                                   // positions are below 2^31 (RT_MAX_P)
            const uint32_t wrap = cyclic ? (uint32_t)C : 0xFFFFFFFFu;           // single round: member index = position
            const uint32_t step = cyclic ? (uint32_t)(256 % C) : 256u;
            uint32_t m = cyclic ? (uint32_t)((t0 + (int64_t)threadIdx.x) % C) : e0 + threadIdx.x;
#pragma unroll 4
            for (int i = (int)threadIdx.x; i < n_valid; i += 256) {
                op[i] = (int32_t)(e0 + (uint32_t)i);
                oc[i] = cs[m];
                m += step;
                if (m >= wrap) m -= wrap;
            }
            continue;
        }
        // Two mappings of a tile onto the CTA, chosen by the alignment of the caller's slice (uniform over the CTA):
        // 16-byte aligned -> four consecutive elements per thread, 16-byte loads and stores; otherwise -> elements 256
        // apart, coalesced 4-byte accesses (16-byte stores at an odd offset would split every sector four ways).
        const bool vec_out = ((reinterpret_cast<uintptr_t>(out_part + off + t0) | reinterpret_cast<uintptr_t>(out_cons + off + t0)) & 15u) == 0;
        if (vec_out) {
            // A thread owns groups of FOUR consecutive elements (group gq = q*256 + tid, q = 0..3): 16-byte loads of the
            // sorted payloads / keys and 16-byte stores of the two result arrays when the caller's slice is 16-byte aligned
            // (topic offset a multiple of 4). ONE round trip to memory per tile: everything that does not depend on loaded
            // data is requested up front (own elements, the consumers of the slots, and — lane 31 only — the first element
            // of the next group for the certificate).
            constexpr int GQ = RT_TILE / (4 * 256);
            static_assert(GQ * 4 * 256 == RT_TILE, "a tile is 256 threads x GQ groups x 4 elements");
            // cyclic topics: position e goes to member e mod C; one modulo per thread and tile, then additions
            const uint32_t cstep = cyclic ? (uint32_t)(1024 % C) : 0u;
            uint32_t cmod = cyclic ? (uint32_t)((t0 + 4 * (int64_t)threadIdx.x) % C) : 0u;
            uint32_t pay[GQ][4], p1x[GQ];
            uint64_t key[GQ][4], k1x[GQ];
            int32_t pid[GQ][4], cons[GQ][4];
    #pragma unroll
            for (int q = 0; q < GQ; ++q) {
                const int i = 4 * (q * 256 + (int)threadIdx.x);            // first element of the group inside the tile
                const int64_t e = t0 + i;
    #pragma unroll
                for (int x = 0; x < 4; ++x) { pay[q][x] = (uint32_t)(e + x); key[q][x] = 0; cons[q][x] = 0; }
                p1x[q] = (uint32_t)(e + 4); k1x[q] = 0;
                if (i < n_valid) {
                    if (sorted) {                                          // (the sort buffers are padded to 16 elements per topic)
                        const uint4 pv = *reinterpret_cast<const uint4*>(fpay + e);
                        pay[q][0] = pv.x; pay[q][1] = pv.y; pay[q][2] = pv.z; pay[q][3] = pv.w;
                        if (need_key) {
                            const ulonglong2 ka = *reinterpret_cast<const ulonglong2*>(fkey + e), kb = *reinterpret_cast<const ulonglong2*>(fkey + e + 2);
                            key[q][0] = ka.x; key[q][1] = ka.y; key[q][2] = kb.x; key[q][3] = kb.y;
                        }
                        if (VERIFY && lane == 31 && i + 4 < n_valid) { k1x[q] = fkey[e + 4]; p1x[q] = fpay[e + 4]; }
                    }
                    if (single_round) {
    #pragma unroll
                        for (int x = 0; x < 4; ++x) if (i + x < n_valid) cons[q][x] = cset_rank[coff + e + x];
                    } else if (cyclic) {
                        uint32_t m = cmod;
    #pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            cons[q][x] = cset_rank[coff + m];
                            m = (m + 1 == (uint32_t)C) ? 0u : m + 1;
                        }
                    }
                }
                if (cyclic) {                       // (for every q: keeps cmod = (t0 + 4 * (q*256 + tid)) mod C)
                    cmod += cstep;
                    if ((int64_t)cmod >= C) cmod -= (uint32_t)C;
                }
            }
    #pragma unroll
            for (int q = 0; q < GQ; ++q) {
                const int i = 4 * (q * 256 + (int)threadIdx.x);
    #pragma unroll
                for (int x = 0; x < 4; ++x) {
                    pid[q][x] = (int32_t)pay[q][x];
                    if (part != nullptr && i + x < n_valid) pid[q][x] = part[off + pay[q][x]];
                }
            }
            if (VERIFY && sorted) {
                // strictly increasing in (key, id when the ids had to be ordered, position); the successor of a group's
                // last element is the next lane's first (lane 31 fetched it; the first group of the next tile is checked
                // by that tile's predecessor check below)
    #pragma unroll
                for (int q = 0; q < GQ; ++q) {
                    const int i = 4 * (q * 256 + (int)threadIdx.x);
                    uint64_t k1 = __shfl_down_sync(0xffffffffu, key[q][0], 1);
                    uint32_t p1 = __shfl_down_sync(0xffffffffu, pay[q][0], 1);
                    int32_t i1 = __shfl_down_sync(0xffffffffu, pid[q][0], 1);
                    if (lane == 31) {
                        k1 = k1x[q];
                        p1 = p1x[q];
                        i1 = (part != nullptr && i + 4 < n_valid) ? part[off + p1] : (int32_t)p1;
                    }
    #pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const uint64_t kn = x < 3 ? key[q][x + 1] : k1;
                        const uint32_t pn = x < 3 ? pay[q][x + 1] : p1;
                        const int32_t in = x < 3 ? pid[q][x + 1] : i1;
                        if (i + x + 1 < n_valid) {
                            bool ok = key[q][x] < kn;
                            if (key[q][x] == kn) ok = tie ? (pid[q][x] < in || (pid[q][x] == in && pay[q][x] < pn)) : (pay[q][x] < pn);
                            bad |= !ok;
                        }
                    }
                }
                // the pair that straddles two tiles of a topic: this tile's first element against its predecessor
                if (threadIdx.x == 0 && t0 > 0) {
                    const uint64_t kp = fkey[t0 - 1];
                    const uint32_t pp = fpay[t0 - 1];
                    bool ok = kp < key[0][0];
                    if (kp == key[0][0]) {
                        if (tie) {
                            const int32_t ip = part[off + pp];
                            ok = ip < pid[0][0] || (ip == pid[0][0] && pp < pay[0][0]);
                        } else {
                            ok = pp < pay[0][0];
                        }
                    }
                    bad |= !ok;
                }
            }
    #pragma unroll
            for (int q = 0; q < GQ; ++q) {
                const int i = 4 * (q * 256 + (int)threadIdx.x);
                const int64_t e = t0 + i;
                if (i >= n_valid) continue;
                const bool has_cons = single_round || cyclic;
                if (i + 4 <= n_valid) {
                    *reinterpret_cast<int4*>(out_part + off + e) = make_int4(pid[q][0], pid[q][1], pid[q][2], pid[q][3]);
                    if (has_cons) *reinterpret_cast<int4*>(out_cons + off + e) = make_int4(cons[q][0], cons[q][1], cons[q][2], cons[q][3]);
                } else {
    #pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        if (i + x < n_valid) {
                            out_part[off + e + x] = pid[q][x];
                            if (has_cons) out_cons[off + e + x] = cons[q][x];
                        }
                    }
                }
                if (out_total != nullptr) {
    #pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int64_t ex = e + x;
                        if (i + x >= n_valid) continue;
                        if (single_round) {
                            out_total[tc_off[t] + ex] = sorted ? (int64_t)(key[q][x] ^ LAG_KEY_MASK) : lag[off + ex];
                            out_count[tc_off[t] + ex] = 1;
                        } else if (cyclic && ex < C) {
                            const int64_t cnt = P / C + (ex < P % C ? 1 : 0);
                            out_total[tc_off[t] + ex] = (int64_t)((uint64_t)cnt * (uint64_t)lag[off]);   // wraps like repeated addition
                            out_count[tc_off[t] + ex] = (int32_t)cnt;
                        }
                    }
                }
            }
        } else {
            constexpr int FB = 8;                       // elements per thread and batch: all loads of a batch before its stores
            // cyclic topics: position e goes to member e mod C. One 64-bit modulo per thread and tile; the elements of a
            // thread are 256 apart, so the rest is additions (a 64-bit modulo per element made this kernel compute-bound:
            // BASELINE config 5, 76 M elements, 0.26 ms)
            const uint32_t cstep = cyclic ? (uint32_t)(256 % C) : 0u;
            uint32_t cmod = cyclic ? (uint32_t)((t0 + (int64_t)threadIdx.x) % C) : 0u;
            for (int i0 = 0; i0 < RT_TILE; i0 += FB * 256) {
                if (t0 + i0 >= P) break;                // uniform over the CTA
                // ONE round trip to memory per batch: everything that does not depend on loaded data is requested up front
                // (own element, the consumer of the slot, and — lane 31 only — the successor element for the certificate)
                uint32_t pay[FB], p1x[FB];
                uint64_t key[FB], k1x[FB];
                int32_t pid[FB], cons[FB];
    #pragma unroll
                for (int q = 0; q < FB; ++q) {
                    const int64_t e = t0 + i0 + q * 256 + (int)threadIdx.x;
                    pay[q] = (uint32_t)e; p1x[q] = (uint32_t)(e + 1);
                    key[q] = 0; k1x[q] = 0;
                    cons[q] = 0;
                    if (e < P) {
                        if (sorted) {
                            pay[q] = fpay[e];
                            if (need_key) key[q] = fkey[e];
                            if (VERIFY && lane == 31 && e + 1 < P) { k1x[q] = fkey[e + 1]; p1x[q] = fpay[e + 1]; }
                        }
                        if (single_round) cons[q] = cset_rank[coff + e];
                        else if (cyclic) cons[q] = cset_rank[coff + cmod];
                    }
                    if (cyclic) {                       // (for every q, also past the end: keeps cmod = (t0 + i0 + q*256 + tid) mod C)
                        cmod += cstep;
                        if ((int64_t)cmod >= C) cmod -= (uint32_t)C;
                    }
                }
    #pragma unroll
                for (int q = 0; q < FB; ++q) {
                    const int64_t e = t0 + i0 + q * 256 + (int)threadIdx.x;
                    pid[q] = (int32_t)pay[q];
                    if (e < P && part != nullptr) pid[q] = part[off + pay[q]];
                }
                if (VERIFY && sorted) {
                    // strictly increasing in (key, id when the ids had to be ordered, position): the successor of an
                    // element is the next lane's
    #pragma unroll
                    for (int q = 0; q < FB; ++q) {
                        const int64_t e = t0 + i0 + q * 256 + (int)threadIdx.x;
                        uint64_t k1 = __shfl_down_sync(0xffffffffu, key[q], 1);
                        uint32_t p1 = __shfl_down_sync(0xffffffffu, pay[q], 1);
                        int32_t i1 = __shfl_down_sync(0xffffffffu, pid[q], 1);
                        if (lane == 31) {
                            k1 = k1x[q];
                            p1 = p1x[q];
                            i1 = (part != nullptr && e + 1 < P) ? part[off + p1] : (int32_t)p1;
                        }
                        if (e + 1 < P) {
                            bool ok = key[q] < k1;
                            if (key[q] == k1) ok = tie ? (pid[q] < i1 || (pid[q] == i1 && pay[q] < p1)) : (pay[q] < p1);
                            bad |= !ok;
                        }
                    }
                }
    #pragma unroll
                for (int q = 0; q < FB; ++q) {
                    const int64_t e = t0 + i0 + q * 256 + (int)threadIdx.x;
                    if (e >= P) continue;
                    out_part[off + e] = pid[q];
                    if (single_round) {
                        out_cons[off + e] = cons[q];
                        if (out_total) {
                            const int64_t lg = sorted ? (int64_t)(key[q] ^ LAG_KEY_MASK) : lag[off + e];
                            out_total[tc_off[t] + e] = lg;
                            out_count[tc_off[t] + e] = 1;
                        }
                    } else if (cyclic) {
                        out_cons[off + e] = cons[q];
                        if (out_total && e < C) {
                            const int64_t cnt = P / C + (e < P % C ? 1 : 0);
                            out_total[tc_off[t] + e] = (int64_t)((uint64_t)cnt * (uint64_t)lag[off]);   // wraps like repeated addition
                            out_count[tc_off[t] + e] = (int32_t)cnt;
                        }
                    }
                }
            }
        }
    }
    if (VERIFY && bad) atomicOr(bt.verify_fail, 1u);
}

// ---- 6. round fill of a large topic: one CTA per topic -----------------------------------

__global__ void __launch_bounds__(BIGFILL_THREADS)
k_big_fill(BigTopics bt, const int32_t* __restrict__ fill_list, int n_fill,
           const int64_t* __restrict__ lag,
           const int32_t* __restrict__ topic_cset, const int64_t* __restrict__ cset_off,
           const int32_t* __restrict__ cset_rank, const int64_t* __restrict__ tc_off,
           int32_t* __restrict__ out_cons, int64_t* __restrict__ out_total, int32_t* __restrict__ out_count) {
#ifdef LAGASSIGN_HOST_EMULATION
    unsigned char* smem_raw = emu::cta().dyn_smem;
#else
    extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
    uint64_t* ckey = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* cidx = reinterpret_cast<uint32_t*>(smem_raw + sizeof(uint64_t) * FILL_CTA_MAX_C);
    for (int li = blockIdx.x; li < n_fill; li += gridDim.x) {
        const int b = fill_list[li];
        if (bt.all_equal[b]) continue;   // assigned in closed form by k_big_finalize
        const int t = bt.topic[b];
        const int64_t P = bt.P[b], off = bt.off[b], boff = bt.boff[b];
        const int sel = bt.sel_final[b];
        const uint64_t* fkey = (sel == SEL_A) ? bt.keyA + boff : bt.keyB + boff;
        const int64_t* lg = lag + off;
        const int s = topic_cset[t];
        const int64_t coff = cset_off[s];
        const int C = (int)(cset_off[s + 1] - coff);
        const int32_t* __restrict__ cset = cset_rank + coff;
        int64_t* tot = out_total ? out_total + tc_off[t] : nullptr;
        int32_t* cnt = out_count ? out_count + tc_off[t] : nullptr;
        auto lag_at = [&](int64_t k) -> int64_t {
            return (sel == SEL_RAW) ? __ldg(lg + k) : (int64_t)(__ldg(fkey + k) ^ LAG_KEY_MASK);
        };
        auto put = [&](int64_t k, int32_t rank) { out_cons[off + k] = rank; };
        uint32_t* xbuf = reinterpret_cast<uint32_t*>(smem_raw);
        if (C <= FILL_WARP_MAX_C) {
            if (threadIdx.x < 32) warp_round_fill_dispatch(P, C, cset, lag_at, put, tot, cnt);
        } else if (C <= 512) {
            cta_round_fill_regs<2, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
        } else if (C <= 1024) {
            cta_round_fill_regs<4, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
        } else if (C <= 2048) {
            cta_round_fill_regs<8, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
        } else if (C <= BIGFILL_REG_MAX_C) {
            cta_round_fill_regs<16, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
        } else if (C <= FILL_CTA_MAX_C) {
            cta_round_fill<BIGFILL_MAXJ>(P, C, max(64, 1 << (32 - __clz(C - 1))), cset, ckey, cidx,
                                         lag_at, put, tot, cnt);
        } else {
            // more consumers than fit on chip (and still more partitions): the same rounds in a global workspace
            const int64_t ws = bt.fill_ws_off[b];
            cta_round_fill_any(P, (int64_t)C, (int64_t)1 << (32 - __clz(C - 1)), cset, bt.fill_ws_key + ws, bt.fill_ws_idx + ws,
                               lag_at, put, tot, cnt);
        }
        __syncthreads();
    }
}

// ---- 7. grouping of the result by member -----------------------
// The reference appends to one list per member (Main:264), so a member's list is ordered by
// topic (iteration order), then by assignment order: exactly the flat result order, filtered
// by member. Grouping is therefore a STABLE partition of the flat result by member rank, done
// with the same radix machinery: the whole flat result is one "topic" whose lag is
// M - rank (descending lag == ascending rank; unassigned elements get M + 1 and sort first).
__global__ void __launch_bounds__(256)
k_group_keys(int64_t n, const int32_t* __restrict__ out_cons, int64_t n_members, int64_t* __restrict__ glag) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int32_t r = out_cons[e];
        glag[e] = r < 0 ? n_members + 1 : n_members - r;
    }
}

// sorted position i -> (topic index, partition) of the element, in the member-grouped arrays;
// member_start[m] = first grouped position of member m (members without partitions stay unset).
// The topic of flat element e is the last t with topic_off[t] <= e: a two-level search, first in a shared-memory
// sample of every `stride`-th offset, then among the `stride` offsets in between (17 dependent global loads per
// element for 100 000 topics made this kernel 0.55 ms at BASELINE config 4; now 11 shared-memory steps + 6 loads).
constexpr int GROUP_COARSE = 2048;
__global__ void __launch_bounds__(256)
k_group_finalize(BigTopics bt, const int64_t* __restrict__ glag, int64_t n_members, int64_t n_unassigned,
                 const int32_t* __restrict__ out_part, const int64_t* __restrict__ topic_off, int64_t n_topics,
                 int32_t* __restrict__ grp_topic, int32_t* __restrict__ grp_part, int64_t* __restrict__ member_start) {
    __shared__ int64_t coarse[GROUP_COARSE];
    const int64_t stride = (n_topics + GROUP_COARSE - 1) / GROUP_COARSE > 0 ? (n_topics + GROUP_COARSE - 1) / GROUP_COARSE : 1;
    const int n_coarse = (int)((n_topics + stride - 1) / stride);                  // coarse[c] = topic_off[c * stride]
    for (int c = threadIdx.x; c < n_coarse; c += blockDim.x) coarse[c] = topic_off[(int64_t)c * stride];
    __syncthreads();
    const int64_t n = bt.P[0];
    const int sel = bt.sel_final[0];
    const uint32_t* fpay = (sel == SEL_A) ? bt.payA : bt.payB;
    const uint64_t* fkey = (sel == SEL_A) ? bt.keyA : bt.keyB;
    // the member rank of sorted position i comes from the SORTED key (a sequential read), not from glag[e] (a random one)
    auto rank_at = [&](int64_t i) -> int64_t {
        return n_members - ((sel == SEL_RAW) ? glag[i] : (int64_t)(fkey[i] ^ LAG_KEY_MASK));
    };
    for (int64_t i = n_unassigned + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = (sel == SEL_RAW) ? i : (int64_t)fpay[i];
        const int64_t rank = rank_at(i);
        const int64_t o = i - n_unassigned;
        int clo = 0, chi = n_coarse;                                               // last c with coarse[c] <= e
        while (chi - clo > 1) {
            const int mid = (clo + chi) >> 1;
            if (coarse[mid] <= e) clo = mid; else chi = mid;
        }
        int64_t lo = (int64_t)clo * stride, hi = min(n_topics, lo + stride);       // last t in [lo, hi) with topic_off[t] <= e
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (topic_off[mid] <= e) lo = mid; else hi = mid;
        }
        grp_topic[o] = (int32_t)lo;
        grp_part[o] = out_part[e];
        const bool first = (i == n_unassigned) || rank_at(i - 1) != rank;
        if (first) member_start[rank] = o;
    }
}

// member_off[m] = start of member m, members without partitions take the next member's start
__global__ void __launch_bounds__(1024)
k_group_offsets(int64_t n_members, int64_t n_assigned, const int64_t* __restrict__ member_start,
                int64_t* __restrict__ member_off) {
    __shared__ int64_t carry;
    __shared__ int64_t wmin[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry = n_assigned; member_off[n_members] = n_assigned; }
    __syncthreads();
    // suffix minimum, chunks of blockDim.x from the end (starts increase with m where set; unset = -1)
    for (int64_t hi = n_members; hi > 0; hi -= blockDim.x) {
        const int64_t m = hi - 1 - tid;
        int64_t v = INT64_MAX;
        if (m >= 0) { const int64_t sv = member_start[m]; if (sv >= 0) v = sv; }
        // inclusive suffix-min over the chunk in thread order (thread 0 = highest m)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v = o < v ? o : v;
        }
        if (lane == 31) wmin[warp] = v;
        __syncthreads();
        int64_t pre = carry;
        for (int w = 0; w < warp; ++w) pre = wmin[w] < pre ? wmin[w] : pre;
        v = pre < v ? pre : v;
        if (m >= 0) member_off[m] = v;
        __syncthreads();
        if (tid == blockDim.x - 1) carry = v;
        __syncthreads();
    }
}

}  // namespace lagassign
