// fill.cuh — the greedy bin-fill of LagBasedPartitionAssignor.java:237-277 in its
// round form.
//
// The reference picks, for every partition in sorted order, the consumer that is
// minimal under (count asc, total asc, member id asc) (Main:240-263) and then adds
// the partition's lag to its total and 1 to its count (Main:265-266). Because
// the count criterion dominates, a consumer that just received a partition cannot
// be chosen again before every other consumer has caught up: partitions are
// consumed in ROUNDS of C, and inside a round the untouched consumers' totals do
// not change, so the round's picks are simply the consumers in ascending
// (total at round start, member id) order. One round = one C-wide sort + one
// C-wide wrapping add; ceil(P/C) rounds are sequentially dependent.
//
// Consumers are (key = total ^ 0x8000..., pay = index in the consumer set); the
// set is listed in member-id order so "smaller id" == "smaller pay".
#pragma once
#include "common.cuh"
#include "sort_cta.cuh"
#include "sort_warp.cuh"

namespace lagassign {

template <int NT> struct Log2NT;
template <> struct Log2NT<256> { static constexpr int value = 8; };
template <> struct Log2NT<512> { static constexpr int value = 9; };
template <> struct Log2NT<1024> { static constexpr int value = 10; };
template <int E> struct Log2E;
template <> struct Log2E<1> { static constexpr int value = 0; };
template <> struct Log2E<2> { static constexpr int value = 1; };
template <> struct Log2E<4> { static constexpr int value = 2; };
template <> struct Log2E<8> { static constexpr int value = 3; };
template <> struct Log2E<16> { static constexpr int value = 4; };

// Sort the C live consumers of a warp by (tkey, cid); padding (KEY_SENTINEL, PAY_SENTINEL) stays behind them.
template <int EPTC>
__device__ __forceinline__ void warp_sort_consumers_packed(uint64_t (&tkey)[EPTC], uint32_t (&cid)[EPTC], int C) {
    constexpr int B = 5 + Log2E<EPTC>::value;          // log2 of the slot count
    constexpr uint32_t IDMASK = (1u << B) - 1u;
    const unsigned lane = lane_id();
    uint64_t lmin = ~0ull, lmax = 0;
#pragma unroll
    for (int r = 0; r < EPTC; ++r) {
        if (cid[r] != PAY_SENTINEL) {
            lmin = tkey[r] < lmin ? tkey[r] : lmin;
            lmax = tkey[r] > lmax ? tkey[r] : lmax;
        }
    }
    const uint32_t mnh = __reduce_min_sync(0xffffffffu, (uint32_t)(lmin >> 32));
    const uint32_t mnl = __reduce_min_sync(0xffffffffu, ((uint32_t)(lmin >> 32) == mnh) ? (uint32_t)lmin : 0xFFFFFFFFu);
    const uint32_t mxh = __reduce_max_sync(0xffffffffu, (uint32_t)(lmax >> 32));
    const uint32_t mxl = __reduce_max_sync(0xffffffffu, ((uint32_t)(lmax >> 32) == mxh) ? (uint32_t)lmax : 0u);
    const uint64_t gmin = ((uint64_t)mnh << 32) | mnl, gmax = ((uint64_t)mxh << 32) | mxl;
    const uint64_t range = gmax - gmin;
    if ((range >> (32 - B)) == 0) {
        uint32_t w[1][EPTC];
#pragma unroll
        for (int r = 0; r < EPTC; ++r)
            w[0][r] = (cid[r] != PAY_SENTINEL) ? (((uint32_t)(tkey[r] - gmin) << B) | cid[r]) : 0xFFFFFFFFu;
        warp_bitonic_sort_words<1, EPTC>(w);
#pragma unroll
        for (int r = 0; r < EPTC; ++r) {
            const bool live = (int)lane * EPTC + r < C;        // padding sorts last
            tkey[r] = live ? gmin + (uint64_t)(w[0][r] >> B) : KEY_SENTINEL;
            cid[r] = live ? (w[0][r] & IDMASK) : PAY_SENTINEL;
        }
    } else if ((range >> (64 - B)) == 0) {
        uint32_t w[2][EPTC];
#pragma unroll
        for (int r = 0; r < EPTC; ++r) {
            const uint64_t v = (cid[r] != PAY_SENTINEL) ? (((tkey[r] - gmin) << B) | cid[r]) : ~0ull;
            w[0][r] = (uint32_t)v;
            w[1][r] = (uint32_t)(v >> 32);
        }
        warp_bitonic_sort_words<2, EPTC>(w);
#pragma unroll
        for (int r = 0; r < EPTC; ++r) {
            const bool live = (int)lane * EPTC + r < C;
            const uint64_t v = ((uint64_t)w[1][r] << 32) | w[0][r];
            tkey[r] = live ? gmin + (v >> B) : KEY_SENTINEL;
            cid[r] = live ? ((uint32_t)v & IDMASK) : PAY_SENTINEL;
        }
    } else {
        warp_bitonic_sort<EPTC>(tkey, cid);
    }
}

// One warp, C <= 32*EPTC consumers in registers (position j = lane*EPTC + r).
//   lag_at(k)        -> int64 lag of the k-th partition in sorted order
//   put(k, rank)     -> record that sorted partition k goes to global member `rank`
//   cset             -> global ranks of the set's members (already offset)
//   out_total/out_count (may be null): per-consumer results, indexed by set position
template <int EPTC, class LagAt, class PutCons>
__device__ __forceinline__ void warp_round_fill(int64_t P, int C, const int32_t* __restrict__ cset,
                                                LagAt lag_at, PutCons put,
                                                int64_t* out_total, int32_t* out_count) {
    const unsigned lane = lane_id();
    uint64_t tkey[EPTC];
    uint32_t cid[EPTC];
    int64_t lag_cur[EPTC];
#pragma unroll
    for (int r = 0; r < EPTC; ++r) {
        int j = (int)lane * EPTC + r;
        bool live = j < C;
        tkey[r] = live ? TOTAL_KEY_MASK : KEY_SENTINEL;   // total 0  (Main:218)
        cid[r] = live ? (uint32_t)j : PAY_SENTINEL;
        lag_cur[r] = (live && j < P) ? lag_at((int64_t)j) : 0;
    }
    for (int64_t base = 0; base < P; base += C) {
        // issue next round's loads before the dependent sort
        int64_t lag_nxt[EPTC];
#pragma unroll
        for (int r = 0; r < EPTC; ++r) {
            int j = (int)lane * EPTC + r;
            int64_t k = base + C + j;
            lag_nxt[r] = (j < C && k < P) ? lag_at(k) : 0;
        }
        // Round 0 starts in id order with all totals 0. After rounds that added the same lag to
        // everyone (typically zero: an idle topic, or the zero-lag tail of a busy one) the
        // order survives: test before sorting.
        if (base > 0) {
            bool ok = true;
#pragma unroll
            for (int r = 0; r + 1 < EPTC; ++r) ok = ok && !elem_less(tkey[r + 1], cid[r + 1], tkey[r], cid[r]);
            const uint64_t nk = __shfl_down_sync(0xffffffffu, tkey[0], 1);
            const uint32_t nc = __shfl_down_sync(0xffffffffu, cid[0], 1);
            if (lane < 31) ok = ok && !elem_less(nk, nc, tkey[EPTC - 1], cid[EPTC - 1]);
            // Totals of one round lie close together, so sort (total - smallest total) << B | index as ONE
            // 32-bit or 64-bit word when it fits instead of three words (same idea as cta_round_fill_regs;
            // config 7 on B200: 0.085 -> 0.073 ms).
            if (!__all_sync(0xffffffffu, ok)) warp_sort_consumers_packed<EPTC>(tkey, cid, C);
        }
#pragma unroll
        for (int r = 0; r < EPTC; ++r) {
            int j = (int)lane * EPTC + r;
            if (j < C && base + j < P) {
                put(base + j, cset[cid[r]]);                                       // Main:264
                tkey[r] = ((tkey[r] ^ TOTAL_KEY_MASK) + (uint64_t)lag_cur[r]) ^ TOTAL_KEY_MASK;  // Main:265, wraps
            }
            lag_cur[r] = lag_nxt[r];
        }
    }
    if (out_total != nullptr) {
        const int64_t full = P / C;
        const int rem = (int)(P % C);
#pragma unroll
        for (int r = 0; r < EPTC; ++r) {
            int j = (int)lane * EPTC + r;
            if (j < C) {
                out_total[cid[r]] = (int64_t)(tkey[r] ^ TOTAL_KEY_MASK);
                out_count[cid[r]] = (int32_t)(full + (j < rem ? 1 : 0));           // Main:266
            }
        }
    }
}

template <class LagAt, class PutCons>
__device__ __forceinline__ void warp_round_fill_dispatch(int64_t P, int C, const int32_t* __restrict__ cset,
                                                         LagAt lag_at, PutCons put,
                                                         int64_t* out_total, int32_t* out_count) {
    if (C <= 32) warp_round_fill<1>(P, C, cset, lag_at, put, out_total, out_count);
    else if (C <= 64) warp_round_fill<2>(P, C, cset, lag_at, put, out_total, out_count);
    else if (C <= 128) warp_round_fill<4>(P, C, cset, lag_at, put, out_total, out_count);
    else warp_round_fill<8>(P, C, cset, lag_at, put, out_total, out_count);
}

// One CTA, 256 < C <= FILL_CTA_MAX_C consumers in shared memory (ckey/cidx hold
// nc = pow2 >= C slots). MAXJ*blockDim.x must cover C. All threads call.
template <int MAXJ, class LagAt, class PutCons>
__device__ __forceinline__ void cta_round_fill(int64_t P, int C, int nc, const int32_t* __restrict__ cset,
                                               uint64_t* ckey, uint32_t* cidx,
                                               LagAt lag_at, PutCons put,
                                               int64_t* out_total, int32_t* out_count) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = tid; j < nc; j += nt) {
        bool live = j < C;
        ckey[j] = live ? TOTAL_KEY_MASK : KEY_SENTINEL;
        cidx[j] = live ? (uint32_t)j : PAY_SENTINEL;
    }
    int64_t lag_cur[MAXJ];
#pragma unroll
    for (int m = 0; m < MAXJ; ++m) {
        int j = tid + m * nt;
        lag_cur[m] = (j < C && j < P) ? lag_at((int64_t)j) : 0;
    }
    __syncthreads();
    for (int64_t base = 0; base < P; base += C) {
        int64_t lag_nxt[MAXJ];
#pragma unroll
        for (int m = 0; m < MAXJ; ++m) {
            int j = tid + m * nt;
            int64_t k = base + C + j;
            lag_nxt[m] = (j < C && k < P) ? lag_at(k) : 0;
        }
        if (base > 0) cta_bitonic_sort(ckey, cidx, nc);
#pragma unroll
        for (int m = 0; m < MAXJ; ++m) {
            int j = tid + m * nt;
            if (j < C && base + j < P) {
                put(base + j, cset[cidx[j]]);
                ckey[j] = ((ckey[j] ^ TOTAL_KEY_MASK) + (uint64_t)lag_cur[m]) ^ TOTAL_KEY_MASK;
            }
            lag_cur[m] = lag_nxt[m];
        }
        // the next sort starts with a __syncthreads()
    }
    __syncthreads();
    if (out_total != nullptr) {
        const int64_t full = P / C;
        const int rem = (int)(P % C);
        for (int j = tid; j < C; j += nt) {
            out_total[cidx[j]] = (int64_t)(ckey[j] ^ TOTAL_KEY_MASK);
            out_count[cidx[j]] = (int32_t)(full + (j < rem ? 1 : 0));
        }
    }
}

// One CTA, ANY number of consumers: ckey/cidx (nc = pow2 >= C slots) live wherever the caller puts them — the
// device-wide path hands in a global-memory workspace when a topic has more consumers than fit in shared memory
// (C > FILL_CTA_MAX_C). Same rounds as above with run-time loops; correct for every shape, not tuned: such
// topics (tens of thousands of consumers AND even more partitions) are rare.
template <class LagAt, class PutCons>
__device__ __forceinline__ void cta_round_fill_any(int64_t P, int64_t C, int64_t nc, const int32_t* __restrict__ cset,
                                                   uint64_t* ckey, uint32_t* cidx, LagAt lag_at, PutCons put,
                                                   int64_t* out_total, int32_t* out_count) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int64_t j = tid; j < nc; j += nt) {
        const bool live = j < C;
        ckey[j] = live ? TOTAL_KEY_MASK : KEY_SENTINEL;
        cidx[j] = live ? (uint32_t)j : PAY_SENTINEL;
    }
    __syncthreads();
    for (int64_t base = 0; base < P; base += C) {
        if (base > 0) cta_bitonic_sort(ckey, cidx, (int)nc);       // starts and ends with a __syncthreads()
        for (int64_t j = tid; j < C && base + j < P; j += nt) {
            put(base + j, cset[cidx[j]]);                                                            // Main:264
            ckey[j] = ((ckey[j] ^ TOTAL_KEY_MASK) + (uint64_t)lag_at(base + j)) ^ TOTAL_KEY_MASK;    // Main:265, wraps
        }
    }
    __syncthreads();
    if (out_total != nullptr) {
        const int64_t full = P / C, rem = P % C;
        for (int64_t j = tid; j < C; j += nt) {
            out_total[cidx[j]] = (int64_t)(ckey[j] ^ TOTAL_KEY_MASK);
            out_count[cidx[j]] = (int32_t)(full + (j < rem ? 1 : 0));                                // Main:266
        }
    }
}

// One CTA of NT threads, consumers in REGISTERS (position j = tid*EPT + r, n =
// NT*EPT >= C): per round one cta_bitonic_sort_regs() and a wrapping add, nothing
// else touches shared memory. Used for 256 < C <= 4096.
template <int EPT, int NT, class LagAt, class PutCons>
__device__ __forceinline__ void cta_round_fill_regs(int64_t P, int C, const int32_t* __restrict__ cset,
                                                    uint32_t* xbuf, LagAt lag_at, PutCons put,
                                                    int64_t* out_total, int32_t* out_count) {
    const int tid = threadIdx.x;
    uint32_t w[3][EPT];          // w[2]:w[1] = total ^ 0x8000..., w[0] = consumer index
    int64_t lag_cur[EPT];
#pragma unroll
    for (int r = 0; r < EPT; ++r) {
        const int j = tid * EPT + r;
        const bool live = j < C;
        w[2][r] = live ? 0x80000000u : 0xFFFFFFFFu;    // total 0 (Main:218) / padding
        w[1][r] = live ? 0u : 0xFFFFFFFFu;
        w[0][r] = live ? (uint32_t)j : PAY_SENTINEL;
        lag_cur[r] = (live && j < P) ? lag_at((int64_t)j) : 0;
    }
    for (int64_t base = 0; base < P; base += C) {
        int64_t lag_nxt[EPT];
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int j = tid * EPT + r;
            const int64_t k = base + C + j;
            lag_nxt[r] = (j < C && k < P) ? lag_at(k) : 0;
        }
        if (base > 0) {
            // After rounds that added the same lag to every consumer (typically zero: an idle
            // topic, or the zero-lag tail of a busy one) the (total, id) order survives: test
            // (two barriers, ~5 % of a sort) and skip the sort when it holds.
            bool skip = false;
            constexpr int NW = NT / 32;
            uint32_t* mm = xbuf + 3 * NT;     // [NW][4]: per-warp min / max of the totals
            uint64_t gmin = ~0ull, gmax = 0;
            {
                bool ok = true;
#pragma unroll
                for (int r = 0; r + 1 < EPT; ++r)
                    ok = ok && !lt3(w[2][r + 1], w[1][r + 1], w[0][r + 1], w[2][r], w[1][r], w[0][r]);
                // smallest and largest total of this warp's live consumers (64-bit, hi then lo)
                uint64_t lmin = ~0ull, lmax = 0;
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    if (tid * EPT + r < C) {
                        const uint64_t tk = ((uint64_t)w[2][r] << 32) | w[1][r];
                        lmin = tk < lmin ? tk : lmin;
                        lmax = tk > lmax ? tk : lmax;
                    }
                }
                const uint32_t mnh = __reduce_min_sync(0xffffffffu, (uint32_t)(lmin >> 32));
                const uint32_t mnl = __reduce_min_sync(0xffffffffu, ((uint32_t)(lmin >> 32) == mnh) ? (uint32_t)lmin : 0xFFFFFFFFu);
                const uint32_t mxh = __reduce_max_sync(0xffffffffu, (uint32_t)(lmax >> 32));
                const uint32_t mxl = __reduce_max_sync(0xffffffffu, ((uint32_t)(lmax >> 32) == mxh) ? (uint32_t)lmax : 0u);
                // first element of the next thread (through shared memory: it may be in another warp)
                __syncthreads();
                xbuf[tid] = w[0][0]; xbuf[NT + tid] = w[1][0]; xbuf[2 * NT + tid] = w[2][0];
                if ((tid & 31) == 0) { uint32_t* row = mm + (tid >> 5) * 4; row[0] = mnh; row[1] = mnl; row[2] = mxh; row[3] = mxl; }
                __syncthreads();
                if (tid + 1 < NT)
                    ok = ok && !lt3(xbuf[2 * NT + tid + 1], xbuf[NT + tid + 1], xbuf[tid + 1], w[2][EPT - 1], w[1][EPT - 1], w[0][EPT - 1]);
                // (read the per-warp rows before the barrier below: the sort reuses this memory)
#pragma unroll
                for (int q = 0; q < NW; ++q) {
                    const uint64_t a = ((uint64_t)mm[q * 4] << 32) | mm[q * 4 + 1];
                    const uint64_t b2 = ((uint64_t)mm[q * 4 + 2] << 32) | mm[q * 4 + 3];
                    gmin = a < gmin ? a : gmin;
                    gmax = b2 > gmax ? b2 : gmax;
                }
                skip = __syncthreads_and(ok) != 0;
            }
            if (!skip) {
                // Totals of one round lie close together (that is what the greedy fill is for):
                // if (total - smallest total) fits above the index bits of one 64-bit word, sort
                // that word (6 instructions per compare-exchange instead of 10); else three words.
                constexpr int B = Log2NT<NT>::value + Log2E<EPT>::value;          // log2(NT*EPT)
                // Small lags keep a round's totals within 2^(32-B) of each other; then ONE 32-bit word per
                // consumer orders them (min/max compare-exchange: 2 instructions instead of 6; one topic of
                // 2^20 partitions over 1024 consumers, lags < 1000, on B200: 4.1 -> 2.3 ms).
                if (((gmax - gmin) >> (32 - B)) == 0) {
                    uint32_t w1[1][EPT];
#pragma unroll
                    for (int r = 0; r < EPT; ++r) {
                        const uint64_t tk = ((uint64_t)w[2][r] << 32) | w[1][r];
                        w1[0][r] = (w[0][r] != PAY_SENTINEL) ? (((uint32_t)(tk - gmin) << B) | w[0][r]) : 0xFFFFFFFFu;
                    }
                    cta_bitonic_sort_regs<1, EPT, NT>(w1, xbuf);
#pragma unroll
                    for (int r = 0; r < EPT; ++r) {
                        const bool live = tid * EPT + r < C;       // padding sorts last
                        const uint64_t tk = gmin + (uint64_t)(w1[0][r] >> B);
                        w[0][r] = live ? (w1[0][r] & ((1u << B) - 1u)) : PAY_SENTINEL;
                        w[1][r] = live ? (uint32_t)tk : 0xFFFFFFFFu;
                        w[2][r] = live ? (uint32_t)(tk >> 32) : 0xFFFFFFFFu;
                    }
                } else
                if (((gmax - gmin) >> (64 - B)) == 0) {
                    uint32_t w2[2][EPT];
#pragma unroll
                    for (int r = 0; r < EPT; ++r) {
                        const bool live = w[0][r] != PAY_SENTINEL;
                        const uint64_t tk = ((uint64_t)w[2][r] << 32) | w[1][r];
                        const uint64_t v = live ? (((tk - gmin) << B) | w[0][r]) : ~0ull;
                        w2[0][r] = (uint32_t)v; w2[1][r] = (uint32_t)(v >> 32);
                    }
                    cta_bitonic_sort_regs<2, EPT, NT>(w2, xbuf);
#pragma unroll
                    for (int r = 0; r < EPT; ++r) {
                        const uint64_t v = ((uint64_t)w2[1][r] << 32) | w2[0][r];
                        const bool live = tid * EPT + r < C;       // padding sorts last
                        const uint64_t tk = (v >> B) + gmin;
                        w[0][r] = live ? (uint32_t)v & ((1u << B) - 1u) : PAY_SENTINEL;
                        w[1][r] = live ? (uint32_t)tk : 0xFFFFFFFFu;
                        w[2][r] = live ? (uint32_t)(tk >> 32) : 0xFFFFFFFFu;
                    }
                } else {
                    cta_bitonic_sort_regs<3, EPT, NT>(w, xbuf);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int j = tid * EPT + r;
            if (j < C && base + j < P) {
                put(base + j, cset[w[0][r]]);                                        // Main:264
                const uint64_t tk = (((uint64_t)w[2][r] << 32) | w[1][r]);
                const uint64_t nk = ((tk ^ TOTAL_KEY_MASK) + (uint64_t)lag_cur[r]) ^ TOTAL_KEY_MASK;  // Main:265
                w[1][r] = (uint32_t)nk;
                w[2][r] = (uint32_t)(nk >> 32);
            }
            lag_cur[r] = lag_nxt[r];
        }
    }
    if (out_total != nullptr) {
        const int64_t full = P / C;
        const int rem = (int)(P % C);
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int j = tid * EPT + r;
            if (j < C) {
                const uint64_t tk = (((uint64_t)w[2][r] << 32) | w[1][r]);
                out_total[w[0][r]] = (int64_t)(tk ^ TOTAL_KEY_MASK);
                out_count[w[0][r]] = (int32_t)(full + (j < rem ? 1 : 0));
            }
        }
    }
}

}  // namespace lagassign
