// launch_cta.cu — translation unit of k_cta_topics (one CTA per topic, P <= 4096)
// and k_unassigned_topics.
#include <algorithm>

#include "kernels_cta.cuh"

namespace lagassign {

cudaError_t configure_cta_kernels() {
    return cudaFuncSetAttribute(k_cta_topics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CtaSmemLayout::total);
}

void launch_unassigned_topics(const DevBatch& b, const int32_t* list, int n, cudaStream_t st) {
    dim3 grid(8, (unsigned)std::min(n, 65535));
    k_unassigned_topics<<<grid, 256, 0, st>>>(b, list, n);
}

void launch_cta_topics(const DevBatch& b, const int32_t* list, int n, int sms, cudaStream_t st) {
    const int grid = std::min(n, sms * 2);
    k_cta_topics<<<grid, CTA_THREADS, CtaSmemLayout::total, st>>>(b, list, n);
}

}  // namespace lagassign
