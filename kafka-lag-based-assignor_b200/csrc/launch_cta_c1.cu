// launch_cta_c1.cu — k_cta_topics<4>: topics of up to 1024 partitions, one CTA each.
#include <algorithm>

#include "kernels_cta.cuh"

namespace lagassign {

static int g_per_sm = 1;

cudaError_t configure_cta_class1() {
    cudaError_t e = cudaFuncSetAttribute(k_cta_topics<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CtaRegSmem<4>::total);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_per_sm, k_cta_topics<4>, CTA_THREADS, CtaRegSmem<4>::total);
}

void launch_cta_class1(const DevBatch& b, const int32_t* list, int n, int sms, cudaStream_t st) {
    const int grid = std::min(n, sms * std::max(1, g_per_sm));
    k_cta_topics<4><<<grid, CTA_THREADS, CtaRegSmem<4>::total, st>>>(b, list, n);
}

}  // namespace lagassign
