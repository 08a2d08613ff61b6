// launch_big.cu — translation unit of the device-wide path (radix.cuh).
#include <algorithm>

#include "radix.cuh"

namespace lagassign {

static int g_scatter_ctas_per_sm = 1;   // resident CTAs per SM of k_big_scatter (the minimum over its instantiations)

template <class K>
static cudaError_t prepare_scatter(K kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem));
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, SC_THREADS, sizeof(PassSmem));
    if (e != cudaSuccess) return e;
    g_scatter_ctas_per_sm = std::max(1, std::min(g_scatter_ctas_per_sm, n));
    return cudaSuccess;
}

cudaError_t configure_big_kernels() {
    g_scatter_ctas_per_sm = 64;
    cudaError_t e = prepare_scatter(k_big_scatter<false, false>);
    if (e != cudaSuccess) return e;
    e = prepare_scatter(k_big_scatter<true, false>);
    if (e != cudaSuccess) return e;
    e = prepare_scatter(k_big_scatter<false, true>);
    if (e != cudaSuccess) return e;
    e = prepare_scatter(k_big_scatter<true, true>);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_big_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIGFILL_SMEM);
}

// Are the lanes of one shared-memory atomic instruction that hit the same address served in lane order?
// (What the fast ranking of k_big_scatter assumes and k_big_verify certifies; probed once per context.)
__global__ void __launch_bounds__(256)
k_probe_atomic_lane_order(uint32_t* violations) {
    __shared__ uint32_t c[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t nbad = 0;
    uint32_t x = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + 12345u;
    for (int it = 0; it < 512; ++it) {
        c[warp][lane] = 0; c[warp][lane + 32] = 0;
        __syncwarp();
        x = x * 1664525u + 1013904223u;
        const uint32_t d = (x >> 20) & ((it & 1) ? 63u : 3u);     // many and few distinct addresses
        const uint32_t r = atomicAdd(&c[warp][d], 1u);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        nbad += (r != (uint32_t)__popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
    }
    if (nbad) atomicAdd(violations, nbad);
}

cudaError_t probe_atomic_lane_order(int* in_order, cudaStream_t st) {
    uint32_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(uint32_t));
    if (e != cudaSuccess) return e;
    uint32_t h = 1;
    e = cudaMemsetAsync(d, 0, sizeof(uint32_t), st);
    if (e == cudaSuccess) { k_probe_atomic_lane_order<<<148, 256, 0, st>>>(d); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    *in_order = (h == 0) ? 1 : 0;
    return e;
}

void launch_big_order_check(const BigTopics& bt, const int32_t* part, int sms, cudaStream_t st) {
    k_big_order_check<<<std::min(bt.n_tiles, sms * 8), 256, 0, st>>>(bt, part);
}

void launch_big_range(const BigTopics& bt, const int64_t* lag, const LagOffsets& offs, const int32_t* part, int sms, cudaStream_t st) {
    const int grid = std::min(bt.n_tiles, sms * 8);
    int64_t* lg = const_cast<int64_t*>(lag);   // written only with raw offsets (the context's own buffer)
    if (part) k_big_range<true><<<grid, RT_THREADS, 0, st>>>(bt, lg, offs, part);
    else k_big_range<false><<<grid, RT_THREADS, 0, st>>>(bt, lg, offs, nullptr);
}

void launch_big_scan(const BigTopics& bt, int first_pass, cudaStream_t st) {
    k_big_scan<<<bt.n_big, 256, 0, st>>>(bt, first_pass);
}

void launch_big_tile_hist(const BigTopics& bt, int pass, const int64_t* lag, const int32_t* part, int sms, cudaStream_t st) {
    const int grid = std::min(bt.n_tiles, sms * 8);
    if (part) k_big_tile_hist<true><<<grid, RT_THREADS, 0, st>>>(bt, pass, lag, part);
    else k_big_tile_hist<false><<<grid, RT_THREADS, 0, st>>>(bt, pass, lag, nullptr);
}

void launch_big_pass_scans(const BigTopics& bt, int pass, cudaStream_t st) {
    k_big_seg_scan<<<bt.n_segs, 256, 0, st>>>(bt, pass);
    k_big_topic_scan<<<bt.n_big, 256, 0, st>>>(bt, pass);
}

void launch_big_scatter(const BigTopics& bt, int pass, const int64_t* lag, const int32_t* part, bool exact, int sms, cudaStream_t st) {
    const int grid = std::min(bt.n_tiles, sms * g_scatter_ctas_per_sm);
    if (part) {
        if (exact) k_big_scatter<true, true><<<grid, SC_THREADS, sizeof(PassSmem), st>>>(bt, pass, lag, part);
        else k_big_scatter<true, false><<<grid, SC_THREADS, sizeof(PassSmem), st>>>(bt, pass, lag, part);
    } else {
        if (exact) k_big_scatter<false, true><<<grid, SC_THREADS, sizeof(PassSmem), st>>>(bt, pass, lag, nullptr);
        else k_big_scatter<false, false><<<grid, SC_THREADS, sizeof(PassSmem), st>>>(bt, pass, lag, nullptr);
    }
}

void launch_big_verify(const BigTopics& bt, const int32_t* part, int sms, cudaStream_t st) {
    k_big_verify<<<std::min(bt.n_tiles, sms * 8), 256, 0, st>>>(bt, part);
}

void launch_big_finalize(const BigTopics& bt, const DevBatch& b, bool verify, int sms, cudaStream_t st) {
    const int grid = std::min(bt.n_tiles, sms * 8);
    if (verify) k_big_finalize<true><<<grid, 256, 0, st>>>(bt, b.lag, b.part, b.topic_cset, b.cset_off, b.cset_rank,
                                                          b.tc_off, b.out_part, b.out_cons, b.out_total, b.out_count);
    else k_big_finalize<false><<<grid, 256, 0, st>>>(bt, b.lag, b.part, b.topic_cset, b.cset_off, b.cset_rank,
                                                     b.tc_off, b.out_part, b.out_cons, b.out_total, b.out_count);
}

void launch_big_fill(const BigTopics& bt, const int32_t* fill_list, int n_fill, const DevBatch& b, int sms, cudaStream_t st) {
    k_big_fill<<<std::min(n_fill, sms), BIGFILL_THREADS, BIGFILL_SMEM, st>>>(
        bt, fill_list, n_fill, b.lag, b.topic_cset, b.cset_off, b.cset_rank, b.tc_off, b.out_cons, b.out_total, b.out_count);
}

void launch_group_keys(int64_t n, const int32_t* out_cons, int64_t n_members, int64_t* glag, int sms, cudaStream_t st) {
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16);
    k_group_keys<<<grid, 256, 0, st>>>(n, out_cons, n_members, glag);
}

void launch_group_finalize(const BigTopics& bt, const int64_t* glag, int64_t n_members, int64_t n_unassigned,
                           const int32_t* out_part, const int64_t* topic_off, int64_t n_topics, int32_t* grp_topic,
                           int32_t* grp_part, int64_t* member_start, int64_t* member_off, int64_t n_assigned, int sms,
                           cudaStream_t st) {
    cudaMemsetAsync(member_start, 0xFF, (size_t)(n_members + 1) * sizeof(int64_t), st);   // -1 = no partitions
    if (n_assigned > 0) {
        const int grid = (int)std::min<int64_t>((n_assigned + 255) / 256, (int64_t)sms * 16);
        k_group_finalize<<<grid, 256, 0, st>>>(bt, glag, n_members, n_unassigned, out_part, topic_off, n_topics,
                                               grp_topic, grp_part, member_start);
    }
    k_group_offsets<<<1, 1024, 0, st>>>(n_members, n_assigned, member_start, member_off);
}

}  // namespace lagassign
