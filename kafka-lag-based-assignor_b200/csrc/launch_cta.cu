// launch_cta.cu — translation unit of k_cta_topics (one CTA per topic, P <= 4096)
// and k_unassigned_topics.
#include <algorithm>

#include "kernels_cta.cuh"

namespace lagassign {

// Topics with an empty consumer set are never visited by the reference (Main:177
// iterates consumersPerTopic; Main:211-213): partitions stay in input order,
// unassigned.
__global__ void __launch_bounds__(256)
k_unassigned_topics(DevBatch b, const int32_t* __restrict__ list, int n_list) {
    for (int li = blockIdx.y; li < n_list; li += gridDim.y) {
        const int t = list[li];
        const int64_t off = b.topic_off[t];
        const int64_t P = b.topic_off[t + 1] - off;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
            b.out_part[off + i] = b.part ? b.part[off + i] : (int32_t)i;
            b.out_cons[off + i] = -1;
        }
    }
}


// one translation unit per size class (launch_cta_c0..3.cu): the register-resident sorts are slow to compile
cudaError_t configure_cta_class0(); cudaError_t configure_cta_class1();
cudaError_t configure_cta_class2(); cudaError_t configure_cta_class3();
void launch_cta_class0(const DevBatch& b, const int32_t* list, int n, int sms, cudaStream_t st);
void launch_cta_class1(const DevBatch& b, const int32_t* list, int n, int sms, cudaStream_t st);
void launch_cta_class2(const DevBatch& b, const int32_t* list, int n, int sms, cudaStream_t st);
void launch_cta_class3(const DevBatch& b, const int32_t* list, int n, int sms, cudaStream_t st);

cudaError_t configure_cta_kernels() {
    cudaError_t e = configure_cta_class0();
    if (e == cudaSuccess) e = configure_cta_class1();
    if (e == cudaSuccess) e = configure_cta_class2();
    if (e == cudaSuccess) e = configure_cta_class3();
    return e;
}

void launch_unassigned_topics(const DevBatch& b, const int32_t* list, int n, cudaStream_t st) {
    dim3 grid(8, (unsigned)std::min(n, 65535));
    k_unassigned_topics<<<grid, 256, 0, st>>>(b, list, n);
}

void launch_cta_topics(const DevBatch& b, const int32_t* list, int n, int cls, int sms, cudaStream_t st) {
    switch (cls) {
        case 0: launch_cta_class0(b, list, n, sms, st); break;
        case 1: launch_cta_class1(b, list, n, sms, st); break;
        case 2: launch_cta_class2(b, list, n, sms, st); break;
        default: launch_cta_class3(b, list, n, sms, st); break;
    }
}

}  // namespace lagassign
