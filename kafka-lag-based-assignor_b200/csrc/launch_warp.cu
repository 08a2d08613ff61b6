// launch_warp.cu — translation unit of k_warp_topics (one warp per topic, P <= 256).
#include "kernels_warp.cuh"

namespace lagassign {

void launch_warp_topics(const DevBatch& b, const int32_t* list, int n, cudaStream_t st) {
    const int grid = (n + SMALL_WARPS - 1) / SMALL_WARPS;
    const bool plain = b.part == nullptr && b.out_total == nullptr;
    if (plain && b.offs.end == nullptr) k_warp_topics<true, LAG_PLAIN><<<grid, SMALL_WARPS * 32, 0, st>>>(b, list, n);
    else if (plain) k_warp_topics<true, LAG_OFFS><<<grid, SMALL_WARPS * 32, 0, st>>>(b, list, n);
    else k_warp_topics<false, LAG_EITHER><<<grid, SMALL_WARPS * 32, 0, st>>>(b, list, n);
}

}  // namespace lagassign
