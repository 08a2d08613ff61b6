// membw.cu — streaming bandwidth probes on B200 behind the HBM rooflines quoted in DESIGN.md (not product code):
// write-only, read-only and copy kernels over 640 MB (far beyond the 126 MB L2), 4-byte and 16-byte accesses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/membw scripts/ubench/membw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <class T> __global__ void __launch_bounds__(256) k_write(T* p, size_t n, T v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
template <class T> __global__ void __launch_bounds__(256) k_write2(T* p, T* q, size_t n, T v) {     // two output streams (finalize writes two arrays)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { p[i] = v; q[i] = v; }
}
__global__ void __launch_bounds__(256) k_read(const int4* p, size_t n, int* sink) {
    int acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { int4 v = p[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678) *sink = acc;
}
__global__ void __launch_bounds__(256) k_copy(const int4* p, int4* q, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) q[i] = p[i];
}
int main() {
    const size_t bytes = 640ull << 20;
    void *a, *b; int* sink;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 4);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, double moved, auto launch) {
        float best = 1e9f;
        for (int it = 0; it < 6; ++it) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
        }
        printf("%-44s %7.3f ms  %7.1f GB/s  %s\n", name, best, moved / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
    };
    const int grid = 148 * 8;
    time("write-only 4 B stores (640 MB)", (double)bytes, [&] { k_write<int><<<grid, 256>>>((int*)a, bytes / 4, 7); });
    time("write-only 16 B stores (640 MB)", (double)bytes, [&] { k_write<int4><<<grid, 256>>>((int4*)a, bytes / 16, make_int4(1, 2, 3, 4)); });
    time("write-only 4 B stores, two arrays (2 x 320 MB)", (double)bytes, [&] { k_write2<int><<<grid, 256>>>((int*)a, (int*)b, bytes / 8, 7); });
    time("cudaMemsetAsync (640 MB)", (double)bytes, [&] { cudaMemsetAsync(a, 3, bytes); });
    time("read-only 16 B loads (640 MB)", (double)bytes, [&] { k_read<<<grid, 256>>>((const int4*)a, bytes / 16, sink); });
    time("copy 16 B (640 MB read + 640 MB written)", 2.0 * bytes, [&] { k_copy<<<grid, 256>>>((const int4*)a, (int4*)b, bytes / 16); });
    return 0;
}
