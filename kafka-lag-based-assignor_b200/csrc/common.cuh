// common.cuh — shared device helpers for the lag-assignor kernels (sm_100a).
//
// Element model used by every sort in this library: a 64-bit unsigned KEY plus a
// 32-bit unsigned PAYLOAD that doubles as the tie-break, ordered by (key, pay)
// ascending. Both sorts of the reference map onto it:
//
//  * partitions  (LagBasedPartitionAssignor.java:228-235: lag DESC by signed
//    Long.compare, ties partition ASC): key = lag ^ 0x7FFF...F turns "signed
//    descending" into "unsigned ascending"; pay = position in a list that is
//    already ascending in partition id.
//  * consumers of a round (:240-263 with the count criterion factored out, see
//    fill.cuh): key = total ^ 0x8000...0 ("signed ascending" as unsigned);
//    pay = index in the topic's consumer list, which is member-id order (:259).
#pragma once
#include <cstdint>
#ifdef LAGASSIGN_HOST_EMULATION
#include "cuda_emu.hpp"   // tests/emu/: the same device templates compiled with g++ (test infrastructure only)
#else
#include <cuda_runtime.h>
#endif

namespace lagassign {

constexpr uint64_t LAG_KEY_MASK   = 0x7FFFFFFFFFFFFFFFull;  // lag  <-> descending key
constexpr uint64_t TOTAL_KEY_MASK = 0x8000000000000000ull;  // total <-> ascending key
constexpr uint64_t KEY_SENTINEL   = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t PAY_SENTINEL   = 0xFFFFFFFFu;

constexpr int WARP_MAX_P = 256;    // topics up to this size: one warp, registers
constexpr int CTA_MAX_P  = 4096;   // up to this size: one CTA, shared memory
constexpr int FILL_WARP_MAX_C = 256;    // consumer rounds sorted by one warp
constexpr int FILL_CTA_MAX_C  = 8192;   // consumer rounds sorted by one CTA in smem

__device__ __forceinline__ bool elem_less(uint64_t ka, uint32_t pa, uint64_t kb, uint32_t pb) {
    return ka < kb || (ka == kb && pa < pb);
}

// number of significant bits of x (0 for x == 0)
__device__ __forceinline__ int bits_of(uint64_t x) {
    const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
    return hi ? 64 - __clz(hi) : (lo ? 32 - __clz(lo) : 0);
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// Streaming loads/stores: inputs are read once and outputs written once.
__device__ __forceinline__ int64_t ld_stream(const int64_t* p) { return __ldcs(p); }
__device__ __forceinline__ int32_t ld_stream(const int32_t* p) { return __ldcs(p); }

}  // namespace lagassign
