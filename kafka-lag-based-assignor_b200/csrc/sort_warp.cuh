// sort_warp.cuh — bitonic sort of 32*EPT (key, pay) elements held in the registers
// of one warp. Logical index i = lane*EPT + r ("blocked"): the low log2(EPT) index
// bits are register bits (compare-exchange without communication), the five high
// bits are lane bits (shuffle). The low bits are the ones a bitonic network visits
// most often, so blocked costs 15 shuffle stages at EPT=8 where a striped layout
// would cost 30. Input placement is arbitrary (the payload remembers where an
// element came from), so loads can stay fully coalesced.
#pragma once
#include "common.cuh"

namespace lagassign {

// In-register compare-exchange: after the call (ka,pa) <= (kb,pb) if asc, else >=.
__device__ __forceinline__ void cmpx(uint64_t& ka, uint32_t& pa, uint64_t& kb, uint32_t& pb, bool asc) {
    bool swap = asc ? elem_less(kb, pb, ka, pa) : elem_less(ka, pa, kb, pb);
    if (swap) {
        uint64_t tk = ka; ka = kb; kb = tk;
        uint32_t tp = pa; pa = pb; pb = tp;
    }
}

template <int EPT>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t (&key)[EPT], uint32_t (&pay)[EPT]) {
    constexpr int N = 32 * EPT;
    const unsigned lane = lane_id();
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= EPT) {
                // partner lives in another lane, same register slot
                const unsigned lj = (unsigned)(j / EPT);
                // direction of the size-k block this element is in (k > j >= EPT: a lane bit)
                const bool asc = (k == N) ? true : ((lane & (unsigned)(k / EPT)) == 0);
                const bool keep_min = (((lane & lj) == 0) == asc);
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    uint64_t ok = __shfl_xor_sync(0xffffffffu, key[r], lj);
                    uint32_t op = __shfl_xor_sync(0xffffffffu, pay[r], lj);
                    bool take = keep_min ? elem_less(ok, op, key[r], pay[r])
                                         : elem_less(key[r], pay[r], ok, op);
                    if (take) { key[r] = ok; pay[r] = op; }
                }
            } else {
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    if ((r & j) == 0) {
                        bool asc;
                        if (k == N) asc = true;
                        else if (k < EPT) asc = ((r & k) == 0);
                        else asc = ((lane & (unsigned)(k / EPT)) == 0);
                        cmpx(key[r], pay[r], key[r | j], pay[r | j], asc);
                    }
                }
            }
        }
    }
}

}  // namespace lagassign
