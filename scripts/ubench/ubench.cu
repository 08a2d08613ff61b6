// ubench.cu — instruction-throughput probes on B200 for the radix ranking design (not product code).
// Each probe: 148 CTAs x 1024 threads, every warp issues ITER x UNROLL independent ops; reports SM cycles per
// warp-instruction (all 32 warps of the SM issuing), i.e. the reciprocal throughput per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITER = 2048;
template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t* out, long long* cyc, uint32_t seed) {
    __shared__ uint32_t sm[8192];
    __shared__ unsigned long long sm64[2048];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 8192; i += 1024) sm[i] = i * 2654435761u;
    for (int i = tid; i < 2048; i += 1024) sm64[i] = i;
    __syncthreads();
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (tid * 2654435761u + j * 40503u + seed) ;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t v = x[j];
            if (OP == 0) { acc += __match_any_sync(0xffffffffu, v & 0xFFu); }
            else if (OP == 1) { acc += __ballot_sync(0xffffffffu, v & 1u); }
            else if (OP == 2) { acc += atomicAdd(&sm[warp * 256 + (v & 0xFFu)], 1u); }
            else if (OP == 3) { acc += __shfl_xor_sync(0xffffffffu, v, 1); }
            else if (OP == 4) { acc += sm[warp * 256 + (v & 0xFFu)]; }
            else if (OP == 5) { sm64[(warp * 64 + (v & 0x3Fu))] = v; }
            else if (OP == 6) { acc += __reduce_add_sync(0xffffffffu, v); }
            else if (OP == 7) { acc += __popc(v) ; }
            else if (OP == 8) { atomicAdd(&sm[warp * 256 + (v & 0xFFu)], 1u); }           // RED (no return)
            else if (OP == 9) { acc += __match_any_sync(0xffffffffu, v & 0xFu); }
            else if (OP == 10) { sm[warp * 256 + (v & 0xFFu)] = v; }
            else if (OP == 11) { acc += (uint32_t)sm64[(warp * 64 + (v & 0x3Fu))]; }
            else if (OP == 12) {   // 8-ballot match (the candidate replacement of match.any)
                uint32_t peers = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const bool bit = (v >> b) & 1u;
                    const uint32_t m = __ballot_sync(0xffffffffu, bit);
                    peers &= bit ? m : ~m;
                }
                acc += peers;
            }
            x[j] = v * 1664525u + 1013904223u + acc * (OP == 7);
        }
    }
    long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 1024 + tid] = acc + sm[tid] + (uint32_t)sm64[tid];
}
// Are same-address shared atomics of one warp instruction served in lane order? (empirical)
__global__ void k_order(uint32_t* bad) {
    __shared__ uint32_t c[32];
    const int lane = threadIdx.x & 31;
    uint32_t nbad = 0;
    for (int it = 0; it < 100000; ++it) {
        if (lane < 32) c[lane] = 0;
        __syncwarp();
        const uint32_t d = ((it * 2654435761u) >> (lane & 7)) & 3u;    // few distinct addresses, many collisions
        const uint32_t r = atomicAdd(&c[d], 1u);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const uint32_t want = __popc(peers & ((1u << lane) - 1u));
        nbad += (r != want);
        __syncwarp();
    }
    atomicAdd(bad, nbad);
}
int main() {
    uint32_t* out; long long* cyc; uint32_t* bad;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    const char* names[] = {"match.any 8-bit", "ballot", "atoms.add ret (warp-private 256 ctrs, random)", "shfl.xor", "lds.32 random(256)",
                           "sts.64 random(64)", "redux.add", "popc(alu ref)", "red.add shared random", "match.any 4-bit", "sts.32 random(256)",
                           "lds.64 random(64)", "8 x (ballot + select) match"};
    auto run = [&](auto kern, int op) {
        kern<<<148, 1024>>>(out, cyc, 1u); cudaDeviceSynchronize();
        kern<<<148, 1024>>>(out, cyc, 2u); cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < 148; ++i) s += (double)h[i];
        s /= 148;
        printf("%-52s %8.2f SM-cycles per warp-instruction (32 warps/SM issuing)  err=%s\n", names[op], s / (32.0 * ITER * 8), cudaGetErrorString(cudaGetLastError()));
    };
    run(k<0>, 0); run(k<1>, 1); run(k<2>, 2); run(k<3>, 3); run(k<4>, 4); run(k<5>, 5); run(k<6>, 6); run(k<7>, 7); run(k<8>, 8);
    run(k<9>, 9); run(k<10>, 10); run(k<11>, 11); run(k<12>, 12);
    k_order<<<64, 32>>>(bad); cudaDeviceSynchronize();
    uint32_t hb = 0; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("same-address shared atomics NOT in lane order: %u of %d\n", hb, 64 * 32 * 100000);
    return 0;
}
