// radix.cuh — device-wide path for topics too large for one CTA (P > 4096):
// a segmented, stable, least-significant-digit radix sort (8-bit digits, 64-bit
// key + 32-bit payload) in the "onesweep" style — one histogram kernel for all
// digit positions, then one pass per digit whose tiles obtain their global
// offsets by decoupled look-back over the previous tiles OF THE SAME TOPIC —
// followed by the finalize / round-fill kernels.
//
// Segmentation: a tile never straddles topics, histograms and look-back chains
// are per topic, and scatter destinations stay inside the topic's own range, so
// any number of large topics share one sequence of launches with no topic digit.
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
#include <cstdio>

namespace lagassign {

// Look-back status words are polled by other CTAs while they are being written:
// the accesses must be real memory operations every time (never hoisted, merged
// or speculated by the compiler), at gpu scope. Flag and value share one 32-bit
// word, so relaxed ordering is sufficient.
__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

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
#pragma unroll 4
        for (int k = 0; k < RT_IPT; ++k) {
            int64_t e = t0 + k * RT_THREADS + threadIdx.x;
            if (e < P) {
                uint64_t key = (uint64_t)ld_stream(lg + e) ^ LAG_KEY_MASK;
#pragma unroll
                for (int d = 0; d < N_LAG_PASS; ++d) atomicAdd(&h[N_TIE_PASS + d][digit_of(key, d)], 1u);
                if (tie) {
                    uint32_t pk = (uint32_t)part[bt.off[b] + e] ^ 0x80000000u;
#pragma unroll
                    for (int d = 0; d < N_TIE_PASS; ++d) atomicAdd(&h[d][(pk >> (8 * d)) & 0xFFu], 1u);
                }
            }
        }
    }
    flush(cur_b);
}

// ---- 3. per topic: prefix sums, skip decisions, buffer ping-pong ------------------------
__global__ void __launch_bounds__(256)
k_big_scan(BigTopics bt, int first_pass) {
    __shared__ uint32_t wsum[8];
    __shared__ int s_cur, s_run;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t P = bt.P[b];
    if (tid == 0) { s_cur = SEL_RAW; s_run = 0; }
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
        atomicMax(bt.passes_run, s_run);
    }
}

// ---- 4. one digit pass ------------------------------------------------------------------
// Persistent CTAs of 256 threads (one thread per digit value). A CTA draws its rank r
// from a counter when it starts and then owns tiles r, r+G, r+2G, ... (G = grid
// size = exactly one resident wave, launched cooperatively): the tiles in flight at
// any moment are consecutive, so the look-back of a tile only ever waits for tiles
// that are being processed right now, even though every CTA has fetched further
// tiles ahead. (Handing out tiles one by one from a ticket counter would let a CTA
// sit on a prefetched ticket whose successors then stall.) Tile data is fetched with
// cp.async (LDGSTS) into a ring of RT_STAGES shared-memory stages, RT_STAGES-1
// tiles ahead, so HBM loads stay in flight while the current tile is ranked and
// scattered; a thread only ever reads back the elements it fetched itself, so
// cp.async.wait_group is the only synchronisation the ring needs. The stage being
// processed doubles as the staging buffer of the tile-local scatter.
#ifndef LAGASSIGN_RT_STAGES
#define LAGASSIGN_RT_STAGES 2
#endif
constexpr int RT_STAGES = LAGASSIGN_RT_STAGES;
#ifndef LAGASSIGN_RT_LOOKBACK
#define LAGASSIGN_RT_LOOKBACK 8
#endif
constexpr int RT_LOOKBACK = LAGASSIGN_RT_LOOKBACK;
static_assert(RT_THREADS == 256, "k_big_pass maps one thread to one digit value");

struct PassSmem {
    uint64_t key[RT_STAGES][RT_TILE];
    uint32_t pay[RT_STAGES][RT_TILE];
    uint32_t whist[RT_WARPS][256];   // per-warp digit counts, then tile-local offsets per (warp, digit)
    int64_t dst_base[256];           // global index of the tile's first element of each digit, minus its tile-local offset
    uint32_t wsum[RT_WARPS];
    int rank;                        // this CTA's arrival rank
};

__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifdef LAGASSIGN_RT_TIMING
#define RT_TICK(slot) do { if (tid == 0) { long long _t = clock64(); t_acc[slot] += _t - t_last; t_last = _t; } } while (0)
#else
#define RT_TICK(slot) do { } while (0)
#endif

template <bool HAS_PART>
__global__ void __launch_bounds__(RT_THREADS, LAGASSIGN_RT_MINBLOCKS)
k_big_pass(BigTopics bt, int pass, const int64_t* __restrict__ lag, const int32_t* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PassSmem& sm = *reinterpret_cast<PassSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (bt.pass_any[pass] == 0) return;   // every topic's digit is constant: nothing to move
    uint32_t* status = bt.status + (size_t)pass * bt.n_tiles * 256;
    uint32_t* cstatus = bt.cstatus + (size_t)pass * bt.n_chunks * 256;
    const int n_tiles = bt.n_tiles;

    // fetch this thread's elements of tile g into ring stage `stage` (one commit group per call)
    auto issue_loads = [&](int stage, int g) {
        if (g < n_tiles) {
            const int b = bt.tile_bt[g];
            const int src = bt.sel_src[b * N_PASS + pass];
            if (src != SEL_SKIP) {
                const int64_t t0 = (int64_t)(g - bt.tile0[b]) * RT_TILE;
                const int n_valid = (int)min((int64_t)RT_TILE, bt.P[b] - t0);
                const int64_t boff = bt.boff[b];
                const uint64_t* skey = ((src == SEL_A) ? bt.keyA : bt.keyB) + boff + t0;
                const uint32_t* spay = ((src == SEL_A) ? bt.payA : bt.payB) + boff + t0;
                const int64_t* lg = lag + bt.off[b] + t0;
#pragma unroll
                for (int k = 0; k < RT_IPT; ++k) {
                    const int i = warp * (32 * RT_IPT) + k * 32 + lane;
                    if (i < n_valid) {
                        if (src == SEL_RAW) {
                            cp_async_8(&sm.key[stage][i], lg + i);
                        } else {
                            cp_async_8(&sm.key[stage][i], skey + i);
                            cp_async_4(&sm.pay[stage][i], spay + i);
                        }
                    }
                }
            }
        }
        cp_async_commit();
    };

#ifdef LAGASSIGN_RT_TIMING
    long long t_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_last = clock64();
    int t_tiles = 0;
#endif
    if (tid == 0) sm.rank = (int)atomicAdd(&bt.ticket[pass], 1u);
    __syncthreads();
    const int G = (int)gridDim.x;
    const int g0 = sm.rank;
    // prologue: the first RT_STAGES-1 tiles of this CTA start loading now
#pragma unroll
    for (int q = 0; q < RT_STAGES - 1; ++q) issue_loads(q, g0 + q * G);

    for (int it = 0;; ++it) {
        const int s = it % RT_STAGES;
        const int sp = (it + RT_STAGES - 1) % RT_STAGES;   // stage freed by the previous iteration
        __syncthreads();   // previous tile fully stored
        RT_TICK(0);
        const int g = g0 + it * G;
        issue_loads(sp, g + (RT_STAGES - 1) * G);
        RT_TICK(1);
        if (g >= n_tiles) break;

        const int b = bt.tile_bt[g];
        const int src = bt.sel_src[b * N_PASS + pass];
        if (src == SEL_SKIP) continue;   // this topic's digit is constant in this pass
        const int dst = bt.sel_dst[b * N_PASS + pass];
        const int64_t P = bt.P[b];
        const int tile_in_topic = g - bt.tile0[b];
        const int64_t t0 = (int64_t)tile_in_topic * RT_TILE;
        const int n_valid = (int)min((int64_t)RT_TILE, P - t0);
        const int64_t boff = bt.boff[b];
        uint64_t* dkey = (dst == SEL_A) ? bt.keyA + boff : bt.keyB + boff;
        uint32_t* dpay = (dst == SEL_A) ? bt.payA + boff : bt.payB + boff;
        const int32_t* pt = HAS_PART ? part + bt.off[b] : nullptr;
        const uint32_t gb = bt.gbase[((size_t)b * N_PASS + pass) * 256 + tid];   // needed after the look-back

#pragma unroll
        for (int w = 0; w < RT_WARPS; ++w) sm.whist[w][tid] = 0;

        RT_TICK(2);
        // this thread's elements: item k of lane l in warp w is tile element w*32*IPT + k*32 + l
        cp_async_wait<RT_STAGES - 1>();
        RT_TICK(3);
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
        RT_TICK(4);

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
        RT_TICK(5);

        // thread d owns digit value d: offsets over warps, tile count, look-back
        {
            uint32_t cnt = 0;
#pragma unroll
            for (int w = 0; w < RT_WARPS; ++w) {
                const uint32_t v = sm.whist[w][tid];
                sm.whist[w][tid] = cnt;
                cnt += v;
            }
            uint32_t valid_cnt = cnt;
            if (tid == 255) valid_cnt -= (uint32_t)(RT_TILE - n_valid);   // padding is not data
            uint32_t* st = status + (size_t)g * 256 + tid;
            st_status(st, ST_AGG | valid_cnt);   // this tile's own count: later tiles can proceed
            // exclusive scan of cnt over digits (tile-local positions; padding is last)
            uint32_t x = cnt;
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

            RT_TICK(6);
            // Two-level decoupled look-back. Tiles of a topic are grouped in chunks of
            // RT_CHUNK. A tile sums the published counts of the earlier tiles of its own
            // chunk (independent loads); the chunk's last tile also publishes the chunk
            // aggregate, walks back over earlier CHUNK words (RT_LOOKBACK at a time) to the
            // nearest inclusive prefix and publishes the chunk's inclusive prefix, which is
            // all the other tiles of the next chunk wait for. A one-level walk costs about
            // G/(4*window) dependent L2 round trips per tile with G tiles in flight, and
            // that chain, not HBM, bounded the pass.
            const int p_in = tile_in_topic % RT_CHUNK;
            const int chunk = tile_in_topic / RT_CHUNK;
            const bool leader = (p_in == RT_CHUNK - 1) || (t0 + RT_TILE >= P);
            uint32_t* cst_base = cstatus + (size_t)bt.chunk0[b] * 256 + tid;
            uint32_t in_chunk = 0;
            {
                const uint32_t* first_in_chunk = status + (size_t)(g - p_in) * 256 + tid;
#pragma unroll
                for (int q0 = 0; q0 < RT_CHUNK - 1; q0 += 8) {
                    if (q0 < p_in) {
                        uint32_t v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            v[q] = (q0 + q < p_in) ? ld_status(first_in_chunk + (size_t)(q0 + q) * 256) : ST_AGG;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            while ((v[q] >> 30) == 0) v[q] = ld_status(first_in_chunk + (size_t)(q0 + q) * 256);
                            in_chunk += v[q] & ST_VAL;
                        }
                    }
                }
            }
            uint32_t prefix_chunks = 0;
            if (leader) {
                const uint32_t chunk_total = in_chunk + valid_cnt;
                uint32_t* cst = cst_base + (size_t)chunk * 256;
                if (chunk == 0) {
                    st_status(cst, ST_INC | chunk_total);
                } else {
                    st_status(cst, ST_AGG | chunk_total);
                    int look = chunk - 1;
                    bool done = false;
                    while (!done) {
                        uint32_t v[RT_LOOKBACK];
#pragma unroll
                        for (int q = 0; q < RT_LOOKBACK; ++q)
                            v[q] = (look - q >= 0) ? ld_status(cst_base + (size_t)(look - q) * 256) : ST_INC;
                        int used = RT_LOOKBACK;
#pragma unroll
                        for (int q = 0; q < RT_LOOKBACK; ++q) {
                            if (!done && used == RT_LOOKBACK) {
                                const uint32_t state = v[q] >> 30;
                                if (state == 0) {
                                    used = q;                // not published yet: poll again from here
                                } else {
                                    prefix_chunks += v[q] & ST_VAL;
                                    if (state == 2u) done = true;
                                }
                            }
                        }
                        look -= used;
                    }
                    st_status(cst, ST_INC | (prefix_chunks + chunk_total));
                }
            } else if (chunk > 0) {
                const uint32_t* cprev = cst_base + (size_t)(chunk - 1) * 256;
                uint32_t v;
                do { v = ld_status(cprev); } while ((v >> 30) != 2u);
                prefix_chunks = v & ST_VAL;
            }
            const uint32_t prefix = prefix_chunks + in_chunk;
            sm.dst_base[tid] = (int64_t)gb + (int64_t)prefix - (int64_t)excl;
        }
        RT_TICK(7);
        __syncthreads();

        // scatter inside the tile (the consumed ring stage is the staging buffer) ...
#pragma unroll
        for (int k = 0; k < RT_IPT; ++k) {
            const uint32_t pos = sm.whist[warp][dig[k]] + rank[k];
            sm.key[s][pos] = key[k];
            sm.pay[s][pos] = pay[k];
        }
        __syncthreads();
        // ... then out in digit runs (coalesced)
        for (int i = tid; i < n_valid; i += RT_THREADS) {
            const uint64_t kk = sm.key[s][i];
            const uint32_t pp = sm.pay[s][i];
            uint32_t d;
            if (HAS_PART && pass < N_TIE_PASS) d = (((uint32_t)pt[pp] ^ 0x80000000u) >> (8 * pass)) & 0xFFu;
            else d = digit_of(kk, pass - N_TIE_PASS);
            const int64_t o = sm.dst_base[d] + i;
            dkey[o] = kk;
            dpay[o] = pp;
        }
        RT_TICK(8);
#ifdef LAGASSIGN_RT_TIMING
        ++t_tiles;
#endif
    }
    cp_async_wait<0>();
#ifdef LAGASSIGN_RT_TIMING
    if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 200) && t_tiles > 0)
        printf("pass %d cta %d tiles %d | per tile: topbar %lld issue %lld meta %lld cpwait %lld lds+bar %lld rank+bar %lld scan %lld lookback %lld scatter+store %lld\n",
               pass, (int)blockIdx.x, t_tiles, t_acc[0] / t_tiles, t_acc[1] / t_tiles, t_acc[2] / t_tiles, t_acc[3] / t_tiles,
               t_acc[4] / t_tiles, t_acc[5] / t_tiles, t_acc[6] / t_tiles, t_acc[7] / t_tiles, t_acc[8] / t_tiles);
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
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* ckey = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* cidx = reinterpret_cast<uint32_t*>(smem_raw + sizeof(uint64_t) * FILL_CTA_MAX_C);
    for (int li = blockIdx.x; li < n_fill; li += gridDim.x) {
        const int b = fill_list[li];
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
        } else {
            cta_round_fill<BIGFILL_MAXJ>(P, C, max(64, 1 << (32 - __clz(C - 1))), cset, ckey, cidx,
                                         lag_at, put, tot, cnt);
        }
        __syncthreads();
    }
}

}  // namespace lagassign
