// launch_cta_c3.cu — k_cta_topics<16>: topics of up to 4096 partitions, one CTA each.
#include <algorithm>

#include "kernels_cta.cuh"

namespace lagassign {

static int g_per_sm = 1;

cudaError_t configure_cta_class3() {
    cudaError_t e = cudaFuncSetAttribute(k_cta_topics<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CtaRegSmem<16>::total);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_per_sm, k_cta_topics<16>, CTA_THREADS, CtaRegSmem<16>::total);
}

void launch_cta_class3(const DevBatch& b, const int32_t* list, int n, int sms, cudaStream_t st) {
    const int grid = std::min(n, sms * std::max(1, g_per_sm));
    k_cta_topics<16><<<grid, CTA_THREADS, CtaRegSmem<16>::total, st>>>(b, list, n);
}

}  // namespace lagassign
