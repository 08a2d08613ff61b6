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

// ---- 2. digit histograms of every pass, per topic -------------------------------------
template <bool HAS_PART>
__global__ void __launch_bounds__(RT_THREADS)
k_big_hist(BigTopics bt, const int64_t* __restrict__ lag, const int32_t* __restrict__ part) {
    __shared__ uint32_t h[N_PASS][256];
    for (int i = threadIdx.x; i < N_PASS * 256; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    // contiguous range of tiles per CTA so the shared histogram is flushed rarely
    const int per = (bt.n_tiles + gridDim.x - 1) / gridDim.x;
    const int g0 = blockIdx.x * per, g1 = min(bt.n_tiles, g0 + per);
    int cur_b = -1;
    auto flush = [&](int b) {
        __syncthreads();
        if (b >= 0) {
            uint32_t* gh = bt.ghist + (size_t)b * N_PASS * 256;
            for (int i = threadIdx.x; i < N_PASS * 256; i += blockDim.x) {
                uint32_t v = (&h[0][0])[i];
                if (v) { atomicAdd(gh + i, v); (&h[0][0])[i] = 0; }
            }
        }
        __syncthreads();
    };
    for (int g = g0; g < g1; ++g) {
        const int b = bt.tile_bt[g];
        if (b != cur_b) { flush(cur_b); cur_b = b; }
        const int64_t P = bt.P[b];
        const int64_t t0 = (int64_t)(g - bt.tile0[b]) * RT_TILE;
        const int64_t* lg = lag + bt.off[b];
        const bool tie = HAS_PART && bt.need_tie[b] != 0;
#ifdef LAGASSIGN_HIST_TILE_SKIP
        // Experimental (compiled out by default; verified under tests/emu, not yet measured on a B200): a digit
        // on which all keys of the tile agree (the high bytes of small lags, every byte of an idle topic) is
        // counted with ONE atomic for the tile instead of one same-address atomic per element.
        __shared__ uint32_t red[RT_WARPS][4];
        uint64_t key[RT_IPT];
        uint64_t o = 0, a = ~0ull;
#pragma unroll
        for (int k = 0; k < RT_IPT; ++k) {
            const int64_t e = t0 + k * RT_THREADS + threadIdx.x;
            key[k] = (e < P) ? ((uint64_t)ld_stream(lg + e) ^ LAG_KEY_MASK) : 0ull;
            if (e < P) { o |= key[k]; a &= key[k]; }
        }
        {
            const uint32_t o_lo = __reduce_or_sync(0xffffffffu, (uint32_t)o), o_hi = __reduce_or_sync(0xffffffffu, (uint32_t)(o >> 32));
            const uint32_t a_lo = __reduce_and_sync(0xffffffffu, (uint32_t)a), a_hi = __reduce_and_sync(0xffffffffu, (uint32_t)(a >> 32));
            if ((threadIdx.x & 31) == 0) {
                uint32_t* row = red[threadIdx.x >> 5];
                row[0] = o_lo; row[1] = o_hi; row[2] = a_lo; row[3] = a_hi;
            }
        }
        __syncthreads();
        uint64_t all_or = 0, all_and = ~0ull;
#pragma unroll
        for (int w = 0; w < RT_WARPS; ++w) {
            all_or |= ((uint64_t)red[w][1] << 32) | red[w][0];
            all_and &= ((uint64_t)red[w][3] << 32) | red[w][2];
        }
        const uint64_t diff = all_or ^ all_and;            // bits in which the tile's keys differ
        const uint32_t n_valid = (uint32_t)min((int64_t)RT_TILE, P - t0);
#pragma unroll
        for (int d = 0; d < N_LAG_PASS; ++d) {
            if (((diff >> (8 * d)) & 0xFFull) == 0) {      // uniform over the CTA
                if (threadIdx.x == 0) atomicAdd(&h[N_TIE_PASS + d][digit_of(all_and, d)], n_valid);
            } else {
#pragma unroll
                for (int k = 0; k < RT_IPT; ++k) {
                    const int64_t e = t0 + k * RT_THREADS + threadIdx.x;
                    if (e < P) atomicAdd(&h[N_TIE_PASS + d][digit_of(key[k], d)], 1u);
                }
            }
        }
        if (tie) {
#pragma unroll 4
            for (int k = 0; k < RT_IPT; ++k) {
                const int64_t e = t0 + k * RT_THREADS + threadIdx.x;
                if (e < P) {
                    const uint32_t pk = (uint32_t)part[bt.off[b] + e] ^ 0x80000000u;
#pragma unroll
                    for (int d = 0; d < N_TIE_PASS; ++d) atomicAdd(&h[d][(pk >> (8 * d)) & 0xFFu], 1u);
                }
            }
        }
        __syncthreads();      // red[] is rewritten by the next tile
#else
#pragma unroll 4
        for (int k = 0; k < RT_IPT; ++k) {
            const int64_t e = t0 + k * RT_THREADS + threadIdx.x;
            const bool live = e < P;
            if (live) {
                const uint64_t key = (uint64_t)ld_stream(lg + e) ^ LAG_KEY_MASK;
#pragma unroll
                for (int d = 0; d < N_LAG_PASS; ++d) atomicAdd(&h[N_TIE_PASS + d][digit_of(key, d)], 1u);
                if (tie) {
                    const uint32_t pk = (uint32_t)part[bt.off[b] + e] ^ 0x80000000u;
#pragma unroll
                    for (int d = 0; d < N_TIE_PASS; ++d) atomicAdd(&h[d][(pk >> (8 * d)) & 0xFFu], 1u);
                }
            }
        }
#endif
    }
    flush(cur_b);
}

// ---- 3. per topic: prefix sums, skip decisions, buffer ping-pong ------------------------
__global__ void __launch_bounds__(256)
k_big_scan(BigTopics bt, int first_pass) {
    __shared__ uint32_t wsum[8];
    __shared__ int s_cur, s_run, s_lagconst;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t P = bt.P[b];
    if (tid == 0) { s_cur = SEL_RAW; s_run = 0; s_lagconst = 0; }
    __syncthreads();
    for (int p = 0; p < N_PASS; ++p) {
        const size_t base = ((size_t)b * N_PASS + p) * 256;
        uint32_t c = bt.ghist[base + tid];
        bool off = p < first_pass || (p < N_TIE_PASS && bt.need_tie[b] == 0);
        int constant = __syncthreads_or((int64_t)c == P);
        // exclusive scan over the 256 digit values
        uint32_t x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        uint32_t add = 0;
        for (int w = 0; w < warp; ++w) add += wsum[w];
        bt.gbase[base + tid] = x - c + add;
        if (tid == 0) {
            if (p >= N_TIE_PASS && constant) s_lagconst += 1;
            if (off || constant) {
                bt.sel_src[b * N_PASS + p] = SEL_SKIP;
                bt.sel_dst[b * N_PASS + p] = SEL_SKIP;
            } else {
                int dst = (s_cur == SEL_A) ? SEL_B : SEL_A;
                bt.sel_src[b * N_PASS + p] = (int8_t)s_cur;
                bt.sel_dst[b * N_PASS + p] = (int8_t)dst;
                s_cur = dst;
                s_run += 1;
                atomicOr(&bt.pass_any[p], 1u);
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        bt.sel_final[b] = (int8_t)s_cur;
        bt.all_equal[b] = (int8_t)(s_lagconst == N_LAG_PASS);
        atomicMax(bt.passes_run, s_run);
    }
}

// ---- 4. one digit pass = four kernels, no inter-CTA waiting ------------------------------
//   4a k_big_tile_hist   digit counts of every tile                    (reads 8 B/element)
//   4b k_big_seg_scan    exclusive scan over the tiles of a segment (RT_SEG tiles), per digit
//   4c k_big_topic_scan  exclusive scan over a topic's segments, seeded with the digit base
//   4d k_big_scatter     rank inside the tile, scatter                 (reads 12, writes 12 B/element)
// An earlier version of this pass was a single "onesweep" kernel whose tiles obtained
// their offsets by decoupled look-back (one- and two-level, windowed). On B200 it stayed
// near 25 % of HBM peak regardless of tile shape, prefetch depth or look-back window:
// with ~450 tiles in flight, per-tile cycle counters showed a third of the time spent
// waiting for predecessor tiles (the chain forces completion in tile order, so every
// tile runs at the pace of the slowest CTA ahead of it). Paying 8 B/element for a
// separate counting kernel removes every dependency between CTAs.
#ifndef LAGASSIGN_RT_STAGES
#define LAGASSIGN_RT_STAGES 2
#endif
constexpr int RT_STAGES = LAGASSIGN_RT_STAGES;
static_assert(RT_THREADS == 256, "the pass kernels map one thread to one digit value");

#ifdef LAGASSIGN_HOST_EMULATION   // tests/emu: the copy happens at issue, groups are no-ops
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) { std::memcpy(smem_dst, gsrc, 8); }
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) { std::memcpy(smem_dst, gsrc, 4); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// 4a ------------------------------------------------------------------------------------------
template <bool HAS_PART>
__global__ void __launch_bounds__(RT_THREADS)
k_big_tile_hist(BigTopics bt, int pass, const int64_t* __restrict__ lag, const int32_t* __restrict__ part) {
    __shared__ uint32_t whist[RT_WARPS][256];
    const int tid = threadIdx.x, warp = tid >> 5;
    if (bt.pass_any[pass] == 0) return;
    for (int g = blockIdx.x; g < bt.n_tiles; g += gridDim.x) {
        const int b = bt.tile_bt[g];
        const int src = bt.sel_src[b * N_PASS + pass];
        if (src == SEL_SKIP) continue;
        const int64_t t0 = (int64_t)(g - bt.tile0[b]) * RT_TILE;
        const int n_valid = (int)min((int64_t)RT_TILE, bt.P[b] - t0);
        const int64_t boff = bt.boff[b];
        const uint64_t* skey = ((src == SEL_A) ? bt.keyA : bt.keyB) + boff + t0;
        const uint32_t* spay = ((src == SEL_A) ? bt.payA : bt.payB) + boff + t0;
        const int64_t* lg = lag + bt.off[b] + t0;
        const int32_t* pt = HAS_PART ? part + bt.off[b] : nullptr;
#pragma unroll
        for (int w = 0; w < RT_WARPS; ++w) whist[w][tid] = 0;
        __syncthreads();
        uint32_t dig[RT_IPT];
#pragma unroll
        for (int k = 0; k < RT_IPT; ++k) {
            const int i = k * RT_THREADS + tid;
            dig[k] = 0xFFFFFFFFu;
            if (i < n_valid) {
                if (HAS_PART && pass < N_TIE_PASS) {
                    const uint32_t pp = (src == SEL_RAW) ? (uint32_t)(t0 + i) : spay[i];
                    dig[k] = (((uint32_t)pt[pp] ^ 0x80000000u) >> (8 * pass)) & 0xFFu;
                } else {
                    const uint64_t key = (src == SEL_RAW) ? ((uint64_t)lg[i] ^ LAG_KEY_MASK) : skey[i];
                    dig[k] = digit_of(key, pass - N_TIE_PASS);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < RT_IPT; ++k)
            if (dig[k] != 0xFFFFFFFFu) atomicAdd(&whist[warp][dig[k]], 1u);
        __syncthreads();
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < RT_WARPS; ++w) c += whist[w][tid];
        bt.tile_hist[(size_t)g * 256 + tid] = c;
        __syncthreads();
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
__global__ void __launch_bounds__(256)
k_big_topic_scan(BigTopics bt, int pass) {
    const int b = blockIdx.x, tid = threadIdx.x;
    if (bt.pass_any[pass] == 0) return;
    if (bt.sel_src[b * N_PASS + pass] == SEL_SKIP) return;
    const int n_tiles_topic = (int)((bt.P[b] + RT_TILE - 1) / RT_TILE);
    const int n = (n_tiles_topic + RT_SEG - 1) / RT_SEG;
    uint32_t* s = bt.seg_total + (size_t)bt.seg0[b] * 256 + tid;
    uint32_t run = bt.gbase[((size_t)b * N_PASS + pass) * 256 + tid];
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
// CTAs of 256 threads (one per digit value) walk the tiles with a grid stride. Tile data is
// fetched with cp.async (LDGSTS) into a ring of RT_STAGES shared-memory stages, RT_STAGES-1
// tiles ahead, so HBM loads stay in flight while the current tile is ranked and scattered; a
// thread only reads back the elements it fetched itself, so cp.async.wait_group is the only
// synchronisation the ring needs. The stage being processed doubles as the staging buffer of
// the tile-local scatter.
struct PassSmem {
    uint64_t key[RT_STAGES][RT_TILE];
    uint32_t pay[RT_STAGES][RT_TILE];
    uint32_t whist[RT_WARPS][256];   // per-warp digit counts, then tile-local offsets per (warp, digit)
    int64_t dst_base[256];           // index in the topic of the tile's first element of each digit, minus its tile-local offset
    uint32_t wsum[RT_WARPS];
};

#ifdef LAGASSIGN_RT_TIMING
#define RT_TICK(slot) do { if (tid == 0) { long long _t = clock64(); t_acc[slot] += _t - t_last; t_last = _t; } } while (0)
#else
#define RT_TICK(slot) do { } while (0)
#endif

template <bool HAS_PART>
__global__ void __launch_bounds__(RT_THREADS, LAGASSIGN_RT_MINBLOCKS)
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
    struct Tile { TileDesc d; int src; };
    auto load_tile = [&](int g) {
        Tile t;
        t.src = SEL_SKIP;
        t.d.b = -1;
        if (g < n_tiles) {
            const int4 a = __ldg(bt.tile_desc + 2 * (size_t)g), c = __ldg(bt.tile_desc + 2 * (size_t)g + 1);
            t.d.b = a.x; t.d.tile_in_topic = a.y; t.d.n_valid = a.z; t.d.seg = a.w;
            t.d.buf_off = ((int64_t)(uint32_t)c.y << 32) | (uint32_t)c.x;
            t.d.raw_off = ((int64_t)(uint32_t)c.w << 32) | (uint32_t)c.z;
            t.src = bt.sel_src[t.d.b * N_PASS + pass];
        }
        return t;
    };
    // fetch this thread's elements of a tile into ring stage `stage` (one commit group per call)
    auto issue_loads = [&](int stage, const Tile& t) {
        if (t.src != SEL_SKIP) {
            const uint64_t* skey = ((t.src == SEL_A) ? bt.keyA : bt.keyB) + t.d.buf_off;
            const uint32_t* spay = ((t.src == SEL_A) ? bt.payA : bt.payB) + t.d.buf_off;
            const int64_t* lg = lag + t.d.raw_off;
#pragma unroll
            for (int k = 0; k < RT_IPT; ++k) {
                const int i = warp * (32 * RT_IPT) + k * 32 + lane;
                if (i < t.d.n_valid) {
                    if (t.src == SEL_RAW) {
                        cp_async_8(&sm.key[stage][i], lg + i);
                    } else {
                        cp_async_8(&sm.key[stage][i], skey + i);
                        cp_async_4(&sm.pay[stage][i], spay + i);
                    }
                }
            }
        }
        cp_async_commit();
    };

    static_assert(RT_STAGES == 2, "the descriptor pipeline below is written for a two-stage ring");
    const int G = (int)gridDim.x;
    const int g0 = (int)blockIdx.x;
    Tile cur = load_tile(g0);              // tile being processed
    issue_loads(0, cur);
    Tile pre = load_tile(g0 + G);          // next tile: its loads are issued at the top of the iteration

    for (int it = 0;; ++it) {
        const int s = it & 1;
        const int g = g0 + it * G;
        __syncthreads();   // previous tile fully stored: its stage may be refilled
        const Tile nxt = pre;
        issue_loads(s ^ 1, nxt);
        pre = load_tile(g + 2 * G);        // consumed by the next iteration
        RT_TICK(0);
        if (g >= n_tiles) break;
        if (cur.src == SEL_SKIP) { cur = nxt; continue; }   // this topic's digit is constant in this pass

        const int b = cur.d.b;
        const int src = cur.src;
        const int dst = bt.sel_dst[b * N_PASS + pass];
        const int tile_in_topic = cur.d.tile_in_topic;
        const int64_t t0 = (int64_t)tile_in_topic * RT_TILE;
        const int n_valid = cur.d.n_valid;
        const int64_t boff = cur.d.buf_off - t0;
        uint64_t* dkey = (dst == SEL_A) ? bt.keyA + boff : bt.keyB + boff;
        uint32_t* dpay = (dst == SEL_A) ? bt.payA + boff : bt.payB + boff;
        const int32_t* pt = HAS_PART ? part + (cur.d.raw_off - t0) : nullptr;
        // where this tile's elements of digit `tid` start in the topic (4b + 4c)
        const uint32_t base_d = bt.tile_hist[(size_t)g * 256 + tid] + bt.seg_total[(size_t)cur.d.seg * 256 + tid];
        cur = nxt;

#pragma unroll
        for (int w = 0; w < RT_WARPS; ++w) sm.whist[w][tid] = 0;
        RT_TICK(1);

        // this thread's elements: item k of lane l in warp w is tile element w*32*IPT + k*32 + l
        cp_async_wait<RT_STAGES - 1>();
        uint64_t key[RT_IPT];
        uint32_t pay[RT_IPT];
        uint32_t dig[RT_IPT];
#pragma unroll
        for (int k = 0; k < RT_IPT; ++k) {
            const int i = warp * (32 * RT_IPT) + k * 32 + lane;
            if (i < n_valid) {
                if (src == SEL_RAW) { key[k] = sm.key[s][i] ^ LAG_KEY_MASK; pay[k] = (uint32_t)(t0 + i); }
                else { key[k] = sm.key[s][i]; pay[k] = sm.pay[s][i]; }
                if (HAS_PART && pass < N_TIE_PASS)
                    dig[k] = (((uint32_t)pt[pay[k]] ^ 0x80000000u) >> (8 * pass)) & 0xFFu;
                else
                    dig[k] = digit_of(key[k], pass - N_TIE_PASS);
            } else {
                key[k] = KEY_SENTINEL; pay[k] = PAY_SENTINEL; dig[k] = 255u;  // sorts behind every valid element
            }
        }
        __syncthreads();   // counters zeroed; every thread holds its elements in registers
        RT_TICK(2);

        // rank inside the warp: elements with equal digit keep memory order
        uint32_t rank[RT_IPT];
#pragma unroll
        for (int k = 0; k < RT_IPT; ++k) {
            const unsigned peers = __match_any_sync(0xffffffffu, dig[k]);
            const uint32_t pre = sm.whist[warp][dig[k]];
            __syncwarp();
            if ((peers & ((1u << lane) - 1u)) == 0) sm.whist[warp][dig[k]] = pre + __popc(peers);
            __syncwarp();
            rank[k] = pre + __popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
        RT_TICK(3);

        // thread d owns digit value d: offsets over warps and over digits inside the tile
        {
            uint32_t cnt = 0;
#pragma unroll
            for (int w = 0; w < RT_WARPS; ++w) {
                const uint32_t v = sm.whist[w][tid];
                sm.whist[w][tid] = cnt;
                cnt += v;
            }
            uint32_t x = cnt;   // padding elements are counted in digit 255, which is last
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            if (lane == 31) sm.wsum[warp] = x;
            __syncthreads();
            uint32_t add = 0;
            for (int w = 0; w < warp; ++w) add += sm.wsum[w];
            const uint32_t excl = x - cnt + add;
#pragma unroll
            for (int w = 0; w < RT_WARPS; ++w) sm.whist[w][tid] += excl;
            sm.dst_base[tid] = (int64_t)base_d - (int64_t)excl;
        }
        __syncthreads();
        RT_TICK(4);

        // scatter inside the tile (the consumed ring stage is the staging buffer) ...
#pragma unroll
        for (int k = 0; k < RT_IPT; ++k) {
            const uint32_t pos = sm.whist[warp][dig[k]] + rank[k];
            sm.key[s][pos] = key[k];
            sm.pay[s][pos] = pay[k];
        }
        __syncthreads();
        // ... then out in digit runs (coalesced)
        {
            uint64_t kk[RT_IPT];
            uint32_t pp[RT_IPT];
            int64_t oo[RT_IPT];
#pragma unroll
            for (int k = 0; k < RT_IPT; ++k) {       // all shared-memory reads first, then the stores
                const int i = k * RT_THREADS + tid;
                kk[k] = sm.key[s][i];
                pp[k] = sm.pay[s][i];
            }
#pragma unroll
            for (int k = 0; k < RT_IPT; ++k) {
                const int i = k * RT_THREADS + tid;
                uint32_t d;
                if (HAS_PART && pass < N_TIE_PASS) d = (i < n_valid) ? ((((uint32_t)pt[pp[k]] ^ 0x80000000u) >> (8 * pass)) & 0xFFu) : 255u;
                else d = digit_of(kk[k], pass - N_TIE_PASS);
                oo[k] = sm.dst_base[d] + i;
            }
#pragma unroll
            for (int k = 0; k < RT_IPT; ++k) {
                const int i = k * RT_THREADS + tid;
                if (i < n_valid) { dkey[oo[k]] = kk[k]; dpay[oo[k]] = pp[k]; }
            }
        }
        RT_TICK(5);
#ifdef LAGASSIGN_RT_TIMING
        ++t_tiles;
#endif
    }
    cp_async_wait<0>();
#ifdef LAGASSIGN_RT_TIMING
    if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 200) && t_tiles > 0)
        printf("pass %d cta %d tiles %d | cycles per tile: top+issue %lld meta %lld cpwait+lds+bar %lld rank+bar %lld scan+bar %lld scatter+store %lld\n",
               pass, (int)blockIdx.x, t_tiles, t_acc[0] / t_tiles, t_acc[1] / t_tiles, t_acc[2] / t_tiles, t_acc[3] / t_tiles,
               t_acc[4] / t_tiles, t_acc[5] / t_tiles);
#endif
}

// ---- 5. finalize: out_partition (and out_consumer for single-round topics) ---------------
__global__ void __launch_bounds__(256)
k_big_finalize(BigTopics bt, const int64_t* __restrict__ lag, const int32_t* __restrict__ part,
               const int32_t* __restrict__ topic_cset, const int64_t* __restrict__ cset_off,
               const int32_t* __restrict__ cset_rank, const int64_t* __restrict__ tc_off,
               int32_t* __restrict__ out_part, int32_t* __restrict__ out_cons,
               int64_t* __restrict__ out_total, int32_t* __restrict__ out_count) {
    for (int g = blockIdx.x; g < bt.n_tiles; g += gridDim.x) {
        const int b = bt.tile_bt[g];
        const int t = bt.topic[b];
        const int64_t P = bt.P[b], off = bt.off[b], boff = bt.boff[b];
        const int sel = bt.sel_final[b];
        const uint64_t* fkey = (sel == SEL_A) ? bt.keyA + boff : bt.keyB + boff;
        const uint32_t* fpay = (sel == SEL_A) ? bt.payA + boff : bt.payB + boff;
        const int s = topic_cset[t];
        const int64_t coff = cset_off[s], C = cset_off[s + 1] - coff;
        const bool single_round = P <= C;
        // All lags equal (an idle topic: zero everywhere): every round starts with equal
        // totals, so its picks are the members in id order and the sorted partition k goes to
        // member k mod C. No sequential fill is needed.
        const bool cyclic = !single_round && bt.all_equal[b] != 0;
        const int64_t t0 = (int64_t)(g - bt.tile0[b]) * RT_TILE;
        for (int i = threadIdx.x; i < RT_TILE; i += blockDim.x) {
            const int64_t e = t0 + i;
            if (e >= P) break;
            const uint32_t pay = (sel == SEL_RAW) ? (uint32_t)e : fpay[e];
            out_part[off + e] = part ? part[off + pay] : (int32_t)pay;
            if (single_round) {
                out_cons[off + e] = cset_rank[coff + e];
                if (out_total) {
                    int64_t lg = (sel == SEL_RAW) ? lag[off + e] : (int64_t)(fkey[e] ^ LAG_KEY_MASK);
                    out_total[tc_off[t] + e] = lg;
                    out_count[tc_off[t] + e] = 1;
                }
            } else if (cyclic) {
                out_cons[off + e] = cset_rank[coff + e % C];
                if (out_total && e < C) {
                    const int64_t cnt = P / C + (e < P % C ? 1 : 0);
                    out_total[tc_off[t] + e] = (int64_t)((uint64_t)cnt * (uint64_t)lag[off]);   // wraps like repeated addition
                    out_count[tc_off[t] + e] = (int32_t)cnt;
                }
            }
        }
    }
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
#if LAGASSIGN_BIGFILL_THREADS == 256
        } else if (C <= 512) {
            cta_round_fill_regs<2, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
        } else if (C <= 1024) {
            cta_round_fill_regs<4, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
        } else if (C <= 2048) {
            cta_round_fill_regs<8, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
        } else if (C <= BIGFILL_REG_MAX_C) {
            cta_round_fill_regs<16, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
#else   // experimental thread counts: the smallest register tile that holds the consumers
        } else if (C <= BIGFILL_THREADS) {
            cta_round_fill_regs<1, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
        } else if (C <= 2 * BIGFILL_THREADS) {
            cta_round_fill_regs<2, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
        } else if (C <= 4 * BIGFILL_THREADS && 4 * BIGFILL_THREADS <= BIGFILL_REG_MAX_C) {
            cta_round_fill_regs<4, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
        } else if (C <= 8 * BIGFILL_THREADS && 8 * BIGFILL_THREADS <= BIGFILL_REG_MAX_C) {
            cta_round_fill_regs<8, BIGFILL_THREADS>(P, C, cset, xbuf, lag_at, put, tot, cnt);
#endif
        } else {
            cta_round_fill<BIGFILL_MAXJ>(P, C, max(64, 1 << (32 - __clz(C - 1))), cset, ckey, cidx,
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
// member_start[m] = first grouped position of member m (members without partitions stay unset)
__global__ void __launch_bounds__(256)
k_group_finalize(BigTopics bt, const int64_t* __restrict__ glag, int64_t n_members, int64_t n_unassigned,
                 const int32_t* __restrict__ out_part, const int64_t* __restrict__ topic_off, int64_t n_topics,
                 int32_t* __restrict__ grp_topic, int32_t* __restrict__ grp_part, int64_t* __restrict__ member_start) {
    const int64_t n = bt.P[0];
    const int sel = bt.sel_final[0];
    const uint32_t* fpay = (sel == SEL_A) ? bt.payA : bt.payB;
    for (int64_t i = n_unassigned + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = (sel == SEL_RAW) ? i : (int64_t)fpay[i];
        const int64_t rank = n_members - glag[e];
        const int64_t o = i - n_unassigned;
        // topic of flat element e: last t with topic_off[t] <= e
        int64_t lo = 0, hi = n_topics;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (topic_off[mid] <= e) lo = mid; else hi = mid;
        }
        grp_topic[o] = (int32_t)lo;
        grp_part[o] = out_part[e];
        bool first = (i == n_unassigned);
        if (!first) {
            const int64_t ep = (sel == SEL_RAW) ? i - 1 : (int64_t)fpay[i - 1];
            first = (n_members - glag[ep]) != rank;
        }
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
