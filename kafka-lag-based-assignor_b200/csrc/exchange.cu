// exchange.cu — compact form of a rank's results for the multi-GPU gather (include/lagassign.h, "result exchange").
//
// With topics sharded over ranks (one process per GPU) every rank finally holds the whole assignment map as two
// int32 arrays. What has to cross NVLink is much less than those 8 bytes per partition:
//   * partition ids fit 1 or 2 bytes whenever every topic has at most 256 / 65536 partitions;
//   * the consumer of a topic with P <= C is not information at all: the single partial round on all-zero totals
//     gives sorted partition k to the k-th member of the topic's set (Main:240-263), which every
//     rank can write from the replicated consumer-set table.
// k_exchange_pack narrows a rank's own results into its slot of the gather buffer; after the all-gather
// k_exchange_unpack widens every slot into the final arrays, one warp per topic, deriving the consumers where it can.
#include "../../include/lagassign.h"

#include <algorithm>
#include <cstdint>
#include <cuda_runtime.h>

namespace {

template <int W> struct Narrow;
template <> struct Narrow<1> { using type = uint8_t; };
template <> struct Narrow<2> { using type = uint16_t; };
template <> struct Narrow<4> { using type = uint32_t; };

// sixteen 1-byte ids per thread and step: four 16-byte loads, one 16-byte store (a byte store per id keeps this
// kernel at a third of the copy bandwidth); [0, n16) with n16 a multiple of 16, both pointers 16-byte aligned
__device__ __forceinline__ void pack_ids_16(int64_t n16, const int32_t* __restrict__ part, uint8_t* __restrict__ dp) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g * 16 < n16; g += (int64_t)gridDim.x * blockDim.x) {
        const int4* src = reinterpret_cast<const int4*>(part + g * 16);
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int4 v = src[q];
            w[q] = ((uint32_t)v.x & 0xFFu) | (((uint32_t)v.y & 0xFFu) << 8) | (((uint32_t)v.z & 0xFFu) << 16) | ((uint32_t)v.w << 24);
        }
        *reinterpret_cast<uint4*>(dp + g * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
__device__ __forceinline__ bool both_aligned16(const void* a, const void* b) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0;
}

template <int PW, int CW>
__global__ void __launch_bounds__(256)
k_exchange_pack(int64_t n, const int32_t* __restrict__ part, const int32_t* __restrict__ cons, uint8_t* __restrict__ dst) {
    using PT = typename Narrow<PW>::type;
    PT* dp = reinterpret_cast<PT*>(dst);
    int64_t first = 0;                         // ids below `first` are packed sixteen at a time
    if constexpr (PW == 1) {
        if (both_aligned16(part, dst)) {
            first = n & ~(int64_t)15;
            pack_ids_16(first, part, dst);
        }
    }
    const int64_t e0 = (CW == 0) ? first : 0;
    for (int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        if (e >= first) dp[e] = (PT)(uint32_t)part[e];
        if constexpr (CW != 0) {
            using CT = typename Narrow<CW>::type;
            CT* dc = reinterpret_cast<CT*>(dst + ((n * PW + 15) & ~(int64_t)15));
            dc[e] = (CT)(uint32_t)cons[e];        // -1 (no subscriber) becomes all ones
        }
    }
}

struct UnpackArgs {
    int64_t max_topics;          // largest range of a rank (grid dimension x covers it, y = rank)
    const int64_t* topic_off;    // [T+1] of the WHOLE batch
    const int32_t* topic_cset;   // [T]
    const int64_t* cset_off;     // [S+1]
    const int32_t* cset_rank;
    const int64_t* range_lo;     // [world] rank r's slot holds the topics [range_lo[r], range_hi[r])
    const int64_t* range_hi;     // [world]
    int world;
    int skip_rank;               // this rank's range is already in place (-1: none)
    int64_t slot_bytes;          // distance between the ranks' slots in `src`
    const uint8_t* src;          // gathered slots
    int32_t* out_part;           // [N]
    int32_t* out_cons;           // [N]
};

template <int PW, int CW>
__global__ void __launch_bounds__(256)
k_exchange_unpack(UnpackArgs a) {
    using PT = typename Narrow<PW>::type;
    const int lane = threadIdx.x & 31;
    const int r = (int)blockIdx.y;                                           // rank whose slot this CTA row widens
    if (r == a.skip_rank) return;
    const int64_t t_first = a.range_lo[r], t_end = a.range_hi[r];
    const int64_t start = a.topic_off[t_first];                              // first partition in that slot
    const int64_t n_r = a.topic_off[t_end] - start;
    const uint8_t* slot = a.src + (int64_t)r * a.slot_bytes;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = t_first + warp0; t < t_end; t += n_warps) {
        const int64_t off = a.topic_off[t], P = a.topic_off[t + 1] - off;
        const PT* sp = reinterpret_cast<const PT*>(slot) + (off - start);
        const int s = a.topic_cset[t];
        const int64_t coff = a.cset_off[s], C = a.cset_off[s + 1] - coff;
        if constexpr (PW == 1 && CW == 0) {
            // the common shape (ids below 256, P <= C, partition counts that are multiples of 4): four entries per
            // lane and step, 4-byte loads of the ids, 16-byte loads of the members and 16-byte stores
            if (((P | off | (off - start) | coff) & 3) == 0 && C > 0) {
                for (int64_t k = (int64_t)lane * 4; k < P; k += 128) {
                    const uchar4 v = *reinterpret_cast<const uchar4*>(sp + k);
                    *reinterpret_cast<int4*>(a.out_part + off + k) = make_int4(v.x, v.y, v.z, v.w);
                    *reinterpret_cast<int4*>(a.out_cons + off + k) = *reinterpret_cast<const int4*>(a.cset_rank + coff + k);
                }
                continue;
            }
        }
        if constexpr (CW == 0) {
            // every topic of the batch has P <= C (or no subscriber): consumers from the table
            for (int64_t k = lane; k < P; k += 32) {
                a.out_part[off + k] = (int32_t)(uint32_t)sp[k];
                a.out_cons[off + k] = C > 0 ? a.cset_rank[coff + k] : -1;
            }
        } else {
            using CT = typename Narrow<CW>::type;
            const CT* sc = reinterpret_cast<const CT*>(slot + ((n_r * PW + 15) & ~(int64_t)15)) + (off - start);
            for (int64_t k = lane; k < P; k += 32) {
                a.out_part[off + k] = (int32_t)(uint32_t)sp[k];
                const CT c = sc[k];
                a.out_cons[off + k] = (c == (CT)~(CT)0) ? -1 : (int32_t)(uint32_t)c;
            }
        }
    }
}

// narrow form with separate destinations (lagassign_assign_batch_narrow): entry e of `part` / `cons` -> dp[e] / dc[e]
template <int PW, int CW>
__global__ void __launch_bounds__(256)
k_narrow_results(int64_t n, const int32_t* __restrict__ part, const int32_t* __restrict__ cons,
                 uint8_t* __restrict__ dpart, uint8_t* __restrict__ dcons) {
    using PT = typename Narrow<PW>::type;
    PT* dp = reinterpret_cast<PT*>(dpart);
    int64_t first = 0;                         // ids below `first` are packed sixteen at a time
    if constexpr (PW == 1) {
        if (both_aligned16(part, dpart)) {
            first = n & ~(int64_t)15;
            pack_ids_16(first, part, dpart);
        }
    }
    const int64_t e0 = (CW == 0) ? first : 0;
    for (int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        if (e >= first) dp[e] = (PT)(uint32_t)part[e];
        if constexpr (CW != 0) {
            using CT = typename Narrow<CW>::type;
            reinterpret_cast<CT*>(dcons)[e] = (CT)(uint32_t)cons[e];        // -1 (no subscriber) becomes all ones
        }
    }
}

thread_local const char* g_exch_err = "";

int exch_fail(const char* m) { g_exch_err = m; return LAGASSIGN_ERR_ARG; }

}  // namespace

namespace lagassign {
// every (pw, cw) of lagassign_narrow_widths except (4, 4), which is the int32 arrays themselves; false otherwise
bool launch_narrow_results(cudaStream_t st, int64_t n, const int32_t* part, const int32_t* cons, int pw, int cw,
                           uint8_t* dpart, uint8_t* dcons) {
    if (n <= 0) return true;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
#define NARROW_CASE(PW, CW) if (pw == PW && cw == CW) { k_narrow_results<PW, CW><<<grid, 256, 0, st>>>(n, part, cons, dpart, dcons); return true; }
    NARROW_CASE(1, 0) NARROW_CASE(2, 0) NARROW_CASE(4, 0) NARROW_CASE(1, 2) NARROW_CASE(2, 2) NARROW_CASE(4, 2)
    NARROW_CASE(1, 4) NARROW_CASE(2, 4)
#undef NARROW_CASE
    return false;
}
}  // namespace lagassign

extern "C" {

int64_t lagassign_exchange_slot_bytes(int64_t n, int part_width, int cons_width) {
    if (n < 0) return -1;
    return ((n * part_width + 15) & ~(int64_t)15) + ((n * cons_width + 15) & ~(int64_t)15);
}

int lagassign_exchange_pack(void* cuda_stream, int64_t n, const int32_t* d_part, const int32_t* d_cons,
                            int part_width, int cons_width, void* d_slot) {
    if (n < 0 || !d_slot || (n > 0 && (!d_part || (cons_width != 0 && !d_cons)))) return exch_fail("NULL argument");
    if (n == 0) return LAGASSIGN_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    uint8_t* dst = static_cast<uint8_t*>(d_slot);
#define PACK_CASE(PW, CW) if (part_width == PW && cons_width == CW) { k_exchange_pack<PW, CW><<<grid, 256, 0, st>>>(n, d_part, d_cons, dst); } else
    PACK_CASE(1, 0) PACK_CASE(2, 0) PACK_CASE(4, 0) PACK_CASE(1, 2) PACK_CASE(2, 2) PACK_CASE(4, 2)
    PACK_CASE(1, 4) PACK_CASE(2, 4) PACK_CASE(4, 4) { return exch_fail("unsupported widths"); }
#undef PACK_CASE
    return cudaGetLastError() == cudaSuccess ? LAGASSIGN_OK : LAGASSIGN_ERR_CUDA;
}

int lagassign_exchange_unpack_ranges(void* cuda_stream, const int64_t* d_topic_off, const int32_t* d_topic_cset,
                                     const int64_t* d_cset_off, const int32_t* d_cset_rank, const int64_t* d_range_lo,
                                     const int64_t* d_range_hi, int world, int skip_rank, int64_t max_topics_per_rank,
                                     const void* d_slots, int64_t slot_bytes, int part_width, int cons_width,
                                     int32_t* d_out_partition, int32_t* d_out_consumer_rank) {
    if (world < 1 || world > 65535 || !d_topic_off || !d_cset_off || !d_range_lo || !d_range_hi || !d_slots ||
        !d_out_partition || !d_out_consumer_rank || (max_topics_per_rank > 0 && !d_topic_cset))
        return exch_fail("NULL argument");
    if (max_topics_per_rank <= 0) return LAGASSIGN_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    UnpackArgs a{max_topics_per_rank, d_topic_off, d_topic_cset, d_cset_off, d_cset_rank, d_range_lo, d_range_hi, world, skip_rank,
                 slot_bytes, static_cast<const uint8_t*>(d_slots), d_out_partition, d_out_consumer_rank};
    const dim3 grid((unsigned)std::min<int64_t>((max_topics_per_rank + 7) / 8, std::max(1, 148 * 16 / world)), (unsigned)world);
#define UNPACK_CASE(PW, CW) if (part_width == PW && cons_width == CW) { k_exchange_unpack<PW, CW><<<grid, 256, 0, st>>>(a); } else
    UNPACK_CASE(1, 0) UNPACK_CASE(2, 0) UNPACK_CASE(4, 0) UNPACK_CASE(1, 2) UNPACK_CASE(2, 2) UNPACK_CASE(4, 2)
    UNPACK_CASE(1, 4) UNPACK_CASE(2, 4) UNPACK_CASE(4, 4) { return exch_fail("unsupported widths"); }
#undef UNPACK_CASE
    return cudaGetLastError() == cudaSuccess ? LAGASSIGN_OK : LAGASSIGN_ERR_CUDA;
}

// contiguous shards: rank r holds the topics [d_shard_first_topic[r], d_shard_first_topic[r + 1])
int lagassign_exchange_unpack(void* cuda_stream, int64_t n_topics, const int64_t* d_topic_off, const int32_t* d_topic_cset,
                              const int64_t* d_cset_off, const int32_t* d_cset_rank, const int64_t* d_shard_first_topic,
                              int world, const void* d_slots, int64_t slot_bytes, int part_width, int cons_width,
                              int32_t* d_out_partition, int32_t* d_out_consumer_rank) {
    if (n_topics < 0 || !d_shard_first_topic) return exch_fail("NULL argument");
    if (n_topics == 0) return LAGASSIGN_OK;
    return lagassign_exchange_unpack_ranges(cuda_stream, d_topic_off, d_topic_cset, d_cset_off, d_cset_rank, d_shard_first_topic,
                                            d_shard_first_topic + 1, world, -1, n_topics, d_slots, slot_bytes, part_width,
                                            cons_width, d_out_partition, d_out_consumer_rank);
}

}  // extern "C"
