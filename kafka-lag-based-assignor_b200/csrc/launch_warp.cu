// launch_warp.cu — translation unit of k_warp_topics (one warp per topic, P <= 256).
#include "kernels_warp.cuh"

namespace lagassign {

void launch_warp_topics(const DevBatch& b, const int32_t* list, int n, cudaStream_t st) {
    const int grid = (n + SMALL_WARPS - 1) / SMALL_WARPS;
#ifdef LAGASSIGN_WARP_GROUPS     // experimental: four topics per warp on the common path (kernels_warp.cuh)
    if (b.part == nullptr && b.out_total == nullptr) {
        const int warps = (n + GRP_TOPICS - 1) / GRP_TOPICS;
        k_warp_topics_grouped<true><<<(warps + SMALL_WARPS - 1) / SMALL_WARPS, SMALL_WARPS * 32, 0, st>>>(b, list, n);
        return;
    }
#endif
    if (b.part == nullptr && b.out_total == nullptr) k_warp_topics<true><<<grid, SMALL_WARPS * 32, 0, st>>>(b, list, n);
    else k_warp_topics<false><<<grid, SMALL_WARPS * 32, 0, st>>>(b, list, n);
}

}  // namespace lagassign
