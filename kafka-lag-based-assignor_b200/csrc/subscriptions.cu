// subscriptions.cu — consumersPerTopic (LagBasedPartitionAssignor.java:410-426) and the de-duplication of consumer
// sets on the device (call sites Main:140-146, :176).
//
// The reference inverts member -> [topic] into topic -> [member] with one HashMap.computeIfAbsent + ArrayList.add per
// subscription entry; include/lagassign.h then wants, per topic, the SET of its members (ascending global rank) and
// identical sets stored once. With strings already replaced by integers on the host (member ranks in
// String.compareTo order, topic indices in the caller's topic order) that is bit-matrix work:
//
//   1. k_sub_bits       bit (t, m) of a T x M bit matrix is set for every subscription entry (atomicOr): the
//                       inversion; a member that lists a topic twice sets the same bit twice (the reference's
//                       consumerTotalLags map de-duplicates it the same way, Main:216-225)
//   2. k_sub_rows       per topic: member count (popcount of its row) and two 64-bit hashes of the row
//   3. k_sub_insert     open-addressing table keyed by (hash, count): every topic finds / claims the slot of its
//                       row's key; the slot keeps the SMALLEST topic index that carries the key
//      k_sub_resolve    rep[t] = that topic, after comparing the two rows word by word; a topic whose row differs from
//                       the slot's representative (a hash collision) is re-inserted under a salted key in another
//                       round, so the result never depends on the hash being collision-free
//   4. k_sub_number     set ids in order of first appearance over the topics (what the host-side flatten produces):
//                       exclusive scans of "is its own representative" and of the representatives' member counts
//   5. k_sub_expand     the rows of the representatives written out as ascending rank lists (cset_off / cset_rank),
//                       topic_cset[t] = set id of rep[t]
//
// Memory: T * ceil(M / 32) words for the matrix (BASELINE config 4: 100 000 x 4096 bits = 51 MB where the entries
// themselves are 4.1e8 x 4 B = 1.6 GB). Exact for every input; no CPU fallback.
#include "../../include/lagassign.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <string>

namespace {

thread_local std::string g_sub_err;

int sub_fail(int code, const char* msg) { g_sub_err = msg; return code; }

#define SUB_TRY(call)                                                                    \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            g_sub_err = std::string(#call) + ": " + cudaGetErrorString(e_);              \
            cleanup();                                                                   \
            return e_ == cudaErrorMemoryAllocation ? LAGASSIGN_ERR_OOM : LAGASSIGN_ERR_CUDA; \
        }                                                                                \
    } while (0)

// ---- 1. inversion ------------------------------------------------------------------------------------------------
// one warp per member, lanes over its entries (coalesced reads of sub_topic; the atomics land in T different rows)
__global__ void __launch_bounds__(256)
k_sub_bits(int64_t n_members, int64_t n_topics, const int64_t* __restrict__ sub_off, const int32_t* __restrict__ sub_topic,
           int words_per_row, uint32_t* __restrict__ bits, uint32_t* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp0; m < n_members; m += n_warps) {
        const int64_t lo = sub_off[m], hi = sub_off[m + 1];
        const uint32_t bit = 1u << (m & 31);
        const int64_t word = m >> 5;
        for (int64_t e = lo + lane; e < hi; e += 32) {
            const int32_t t = sub_topic[e];
            if (t < 0 || t >= n_topics) { atomicOr(bad, 1u); continue; }
            uint32_t* w = bits + (size_t)t * words_per_row + word;
            if ((*w & bit) == 0) atomicOr(w, bit);        // (plain read first: repeated entries cost no atomic)
        }
    }
}

// ---- 2. per topic: member count and row hash ---------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x) {           // splitmix64 finaliser
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

__global__ void __launch_bounds__(256)
k_sub_rows(int64_t n_topics, int words_per_row, const uint32_t* __restrict__ bits, int32_t* __restrict__ count,
           uint64_t* __restrict__ hash_a, uint64_t* __restrict__ hash_b) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp0; t < n_topics; t += n_warps) {
        const uint32_t* row = bits + (size_t)t * words_per_row;
        int c = 0;
        uint64_t a = 0, b = 0;
        for (int w = lane; w < words_per_row; w += 32) {
            const uint32_t v = row[w];
            c += __popc(v);
            // position-dependent contributions combined with + and ^ (order-free across lanes)
            a += mix64(((uint64_t)(uint32_t)w << 32 | v) + 0x9E3779B97F4A7C15ull);
            b ^= mix64(((uint64_t)v << 32 | (uint32_t)w) ^ 0xD6E8FEB86659FD93ull);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, d);
            a += __shfl_xor_sync(0xffffffffu, a, d);
            b ^= __shfl_xor_sync(0xffffffffu, b, d);
        }
        if (lane == 0) { count[t] = c; hash_a[t] = a; hash_b[t] = b; }
    }
}

// ---- 3. equal rows -> one representative ----------------------------------------------------------------------------
// One 64-bit key per row (its hash, its member count, the round's salt) and the smallest topic index carrying it,
// stored as 0x7FFFFFFF - topic so that "empty" is 0 and atomicMax keeps the smallest topic. Equality of rows is
// never taken from the key: k_sub_resolve compares the rows.
struct Slot { unsigned long long key; unsigned int enc_topic; unsigned int pad; };

__device__ __forceinline__ unsigned long long slot_key(uint64_t ha, uint64_t hb, int32_t count, uint32_t salt) {
    return mix64(ha ^ mix64(hb + salt) ^ ((uint64_t)(uint32_t)count << 32)) | 1ull;      // never the "empty" pattern 0
}

__device__ void table_insert(Slot* table, uint64_t mask, unsigned long long key, int32_t t) {
    uint64_t i = (key >> 1) & mask;
    for (;;) {
        Slot* s = table + i;
        const unsigned long long cur = atomicCAS(&s->key, 0ull, key);
        if (cur == 0ull || cur == key) { atomicMax(&s->enc_topic, 0x7FFFFFFFu - (unsigned int)t); return; }
        i = (i + 1) & mask;
    }
}

__device__ int32_t table_find(const Slot* table, uint64_t mask, unsigned long long key) {
    uint64_t i = (key >> 1) & mask;
    for (;;) {
        const Slot* s = table + i;
        if (s->key == key) return (int32_t)(0x7FFFFFFFu - s->enc_topic);
        if (s->key == 0ull) return -1;
        i = (i + 1) & mask;
    }
}

// todo[] = topics still without a representative (all of them in round 0); salt distinguishes the rounds
__global__ void __launch_bounds__(256)
k_sub_insert(int64_t n_todo, const int32_t* __restrict__ todo, const int32_t* __restrict__ count,
             const uint64_t* __restrict__ hash_a, const uint64_t* __restrict__ hash_b, uint32_t salt, Slot* table, uint64_t mask) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_todo; j += (int64_t)gridDim.x * blockDim.x) {
        const int32_t t = todo ? todo[j] : (int32_t)j;
        table_insert(table, mask, slot_key(hash_a[t], hash_b[t], count[t], salt), t);
    }
}

// one warp per pending topic: representative of its key, rows compared word by word
__global__ void __launch_bounds__(256)
k_sub_resolve(int64_t n_todo, const int32_t* __restrict__ todo, const int32_t* __restrict__ count,
              const uint64_t* __restrict__ hash_a, const uint64_t* __restrict__ hash_b, uint32_t salt,
              const Slot* __restrict__ table, uint64_t mask, int words_per_row, const uint32_t* __restrict__ bits,
              int32_t* __restrict__ rep, int32_t* __restrict__ next_todo, unsigned long long* __restrict__ n_next) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp0; j < n_todo; j += n_warps) {
        const int32_t t = todo ? todo[j] : (int32_t)j;
        const int32_t r = table_find(table, mask, slot_key(hash_a[t], hash_b[t], count[t], salt));
        bool same = r >= 0;
        if (same && r != t) {
            const uint32_t* x = bits + (size_t)t * words_per_row;
            const uint32_t* y = bits + (size_t)r * words_per_row;
            bool eq = true;
            for (int w = lane; w < words_per_row; w += 32) eq = eq && (x[w] == y[w]);
            same = __all_sync(0xffffffffu, eq);
        }
        if (lane == 0) {
            if (same) rep[t] = r;
            else next_todo[atomicAdd(n_next, 1ull)] = t;      // a different row under the same 128-bit key: next round
        }
    }
}

// ---- 4. numbering: exclusive scans over the topics (one CTA, chunks with a carry) -----------------------------------
// is_rep[t] = rep[t] == t;  set_id[t] = #representatives before t;  set_start[t] = members of those representatives
__global__ void __launch_bounds__(1024)
k_sub_number(int64_t n_topics, const int32_t* __restrict__ rep, const int32_t* __restrict__ count,
             int32_t* __restrict__ set_id, int64_t* __restrict__ set_start, int64_t* __restrict__ totals) {
    __shared__ int64_t ws_n[32], ws_c[32];
    __shared__ int64_t carry_n, carry_c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { carry_n = 0; carry_c = 0; }
    __syncthreads();
    for (int64_t base = 0; base < n_topics; base += blockDim.x) {
        const int64_t t = base + tid;
        const bool isr = t < n_topics && rep[t] == (int32_t)t;
        int64_t n = isr ? 1 : 0, c = isr ? count[t] : 0;
        int64_t xn = n, xc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t yn = __shfl_up_sync(0xffffffffu, xn, d), yc = __shfl_up_sync(0xffffffffu, xc, d);
            if (lane >= d) { xn += yn; xc += yc; }
        }
        if (lane == 31) { ws_n[warp] = xn; ws_c[warp] = xc; }
        __syncthreads();
        int64_t pn = carry_n, pc = carry_c;
        for (int w = 0; w < warp; ++w) { pn += ws_n[w]; pc += ws_c[w]; }
        if (t < n_topics) { set_id[t] = (int32_t)(pn + xn - n); set_start[t] = pc + xc - c; }
        __syncthreads();
        if (tid == blockDim.x - 1) { carry_n = pn + xn; carry_c = pc + xc; }
        __syncthreads();
    }
    if (tid == 0) { totals[0] = carry_n; totals[1] = carry_c; }
}

// ---- 5. output ---------------------------------------------------------------------------------------------------------
// one warp per topic: topic_cset[t] = set id of its representative; a representative also writes its offsets and ranks
__global__ void __launch_bounds__(256)
k_sub_expand(int64_t n_topics, int words_per_row, const uint32_t* __restrict__ bits, const int32_t* __restrict__ rep,
             const int32_t* __restrict__ count, const int32_t* __restrict__ set_id, const int64_t* __restrict__ set_start,
             int32_t* __restrict__ topic_cset, int64_t* __restrict__ cset_off, int32_t* __restrict__ cset_rank) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp0; t < n_topics; t += n_warps) {
        const int32_t r = rep[t];
        if (lane == 0) topic_cset[t] = set_id[r];
        if (r != (int32_t)t) continue;
        const int32_t s = set_id[t];
        const int64_t start = set_start[t];
        if (lane == 0) { cset_off[s] = start; cset_off[s + 1] = start + count[t]; }   // (the next set rewrites the same value)
        const uint32_t* row = bits + (size_t)t * words_per_row;
        int64_t run = start;
        for (int w0 = 0; w0 < words_per_row; w0 += 32) {
            const int w = w0 + lane;
            const uint32_t v = w < words_per_row ? row[w] : 0u;
            int x = __popc(v);
            const int mine = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            int64_t o = run + x - mine;
            uint32_t rest = v;
            while (rest) {
                const int bit = __ffs(rest) - 1;
                rest &= rest - 1;
                cset_rank[o++] = (int32_t)(w * 32 + bit);
            }
            run += __shfl_sync(0xffffffffu, x, 31);
        }
    }
}

template <class T>
struct DevArr {
    T* p = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
    ~DevArr() { if (p) cudaFree(p); }
};

}  // namespace

extern "C" const char* lagassign_subscriptions_last_error(void) { return g_sub_err.c_str(); }

extern "C" int lagassign_invert_subscriptions(int device, const lagassign_subscriptions* in, lagassign_consumer_sets* out) {
    if (!in || !out) return sub_fail(LAGASSIGN_ERR_ARG, "subscriptions / result is NULL");
    const int64_t M = in->n_members, T = in->n_topics;
    if (M < 0 || T < 0 || T > 0x7FFFFFFF || M > 0x7FFFFFFF) return sub_fail(LAGASSIGN_ERR_ARG, "n_members / n_topics out of range");
    if (M > 0 && !in->sub_off) return sub_fail(LAGASSIGN_ERR_ARG, "sub_off is NULL");
    const int64_t R = M > 0 ? in->sub_off[M] : 0;
    if (M > 0 && in->sub_off[0] != 0) return sub_fail(LAGASSIGN_ERR_ARG, "sub_off must start at 0");
    for (int64_t m = 0; m < M; ++m)
        if (in->sub_off[m + 1] < in->sub_off[m]) return sub_fail(LAGASSIGN_ERR_ARG, "sub_off not monotone");
    if (R > 0 && !in->sub_topic) return sub_fail(LAGASSIGN_ERR_ARG, "sub_topic is NULL");
    if (T > 0 && (!out->topic_cset || !out->cset_off || (R > 0 && !out->cset_rank)))
        return sub_fail(LAGASSIGN_ERR_ARG, "output arrays must not be NULL");
    out->n_csets = 0;
    if (T == 0) { if (out->cset_off) out->cset_off[0] = 0; return LAGASSIGN_OK; }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return sub_fail(LAGASSIGN_ERR_NO_DEVICE, "no CUDA device available; this library has no CPU fallback");
    }
    const int wpr = (int)((M + 31) / 32 > 0 ? (M + 31) / 32 : 1);
    const size_t n_words = (size_t)T * wpr;
    if (n_words > ((size_t)1 << 34)) return sub_fail(LAGASSIGN_ERR_UNSUPPORTED, "topic x member bit matrix larger than 64 GiB");

    cudaStream_t st = nullptr;
    DevArr<int64_t> d_sub_off, d_set_start, d_totals, d_cset_off;
    DevArr<int32_t> d_sub_topic, d_count, d_rep, d_set_id, d_todo_a, d_todo_b, d_topic_cset, d_cset_rank;
    DevArr<uint32_t> d_bits, d_bad;
    DevArr<uint64_t> d_ha, d_hb;
    DevArr<unsigned long long> d_n_next;
    DevArr<Slot> d_table;
    auto cleanup = [&] { if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); st = nullptr; } cudaGetLastError(); };

    SUB_TRY(cudaSetDevice(device));
    SUB_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    uint64_t table_size = 64;
    while (table_size < (uint64_t)T * 2) table_size <<= 1;
    SUB_TRY(d_sub_off.alloc((size_t)M + 1));
    SUB_TRY(d_sub_topic.alloc((size_t)R));
    SUB_TRY(d_bits.alloc(n_words));
    SUB_TRY(d_bad.alloc(1));
    SUB_TRY(d_count.alloc((size_t)T));
    SUB_TRY(d_ha.alloc((size_t)T));
    SUB_TRY(d_hb.alloc((size_t)T));
    SUB_TRY(d_rep.alloc((size_t)T));
    SUB_TRY(d_set_id.alloc((size_t)T));
    SUB_TRY(d_set_start.alloc((size_t)T));
    SUB_TRY(d_totals.alloc(2));
    SUB_TRY(d_todo_a.alloc((size_t)T));
    SUB_TRY(d_todo_b.alloc((size_t)T));
    SUB_TRY(d_n_next.alloc(1));
    SUB_TRY(d_table.alloc((size_t)table_size));
    SUB_TRY(d_topic_cset.alloc((size_t)T));
    SUB_TRY(d_cset_off.alloc((size_t)T + 1));
    SUB_TRY(d_cset_rank.alloc((size_t)std::max<int64_t>(R, 1)));

    if (M > 0) SUB_TRY(cudaMemcpyAsync(d_sub_off.p, in->sub_off, ((size_t)M + 1) * 8, cudaMemcpyHostToDevice, st));
    if (R > 0) SUB_TRY(cudaMemcpyAsync(d_sub_topic.p, in->sub_topic, (size_t)R * 4, cudaMemcpyHostToDevice, st));
    SUB_TRY(cudaMemsetAsync(d_bits.p, 0, n_words * 4, st));
    SUB_TRY(cudaMemsetAsync(d_bad.p, 0, 4, st));
    SUB_TRY(cudaMemsetAsync(d_cset_off.p, 0, ((size_t)T + 1) * 8, st));
    const int grid = sms * 8;
    if (M > 0 && R > 0) k_sub_bits<<<grid, 256, 0, st>>>(M, T, d_sub_off.p, d_sub_topic.p, wpr, d_bits.p, d_bad.p);
    k_sub_rows<<<grid, 256, 0, st>>>(T, wpr, d_bits.p, d_count.p, d_ha.p, d_hb.p);
    // rounds: all topics first; a topic whose row differs from its key's representative goes round again with another salt
    int64_t n_todo = T;
    const int32_t* todo = nullptr;
    int32_t* next = d_todo_a.p;
    for (uint32_t round = 0; n_todo > 0; ++round) {
        if (round > 64) { cleanup(); return sub_fail(LAGASSIGN_ERR_CUDA, "consumer-set de-duplication did not converge"); }
        SUB_TRY(cudaMemsetAsync(d_table.p, 0, (size_t)table_size * sizeof(Slot), st));
        SUB_TRY(cudaMemsetAsync(d_n_next.p, 0, 8, st));
        k_sub_insert<<<grid, 256, 0, st>>>(n_todo, todo, d_count.p, d_ha.p, d_hb.p, round * 0x9E3779B9u, d_table.p, table_size - 1);
        k_sub_resolve<<<grid, 256, 0, st>>>(n_todo, todo, d_count.p, d_ha.p, d_hb.p, round * 0x9E3779B9u, d_table.p,
                                           table_size - 1, wpr, d_bits.p, d_rep.p, next, d_n_next.p);
        unsigned long long n_next = 0;
        SUB_TRY(cudaMemcpyAsync(&n_next, d_n_next.p, 8, cudaMemcpyDeviceToHost, st));
        SUB_TRY(cudaStreamSynchronize(st));
        n_todo = (int64_t)n_next;
        todo = next;
        next = (next == d_todo_a.p) ? d_todo_b.p : d_todo_a.p;
    }
    k_sub_number<<<1, 1024, 0, st>>>(T, d_rep.p, d_count.p, d_set_id.p, d_set_start.p, d_totals.p);
    k_sub_expand<<<grid, 256, 0, st>>>(T, wpr, d_bits.p, d_rep.p, d_count.p, d_set_id.p, d_set_start.p, d_topic_cset.p,
                                      d_cset_off.p, d_cset_rank.p);
    int64_t totals[2] = {0, 0};
    uint32_t bad = 0;
    SUB_TRY(cudaMemcpyAsync(totals, d_totals.p, 16, cudaMemcpyDeviceToHost, st));
    SUB_TRY(cudaMemcpyAsync(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost, st));
    SUB_TRY(cudaStreamSynchronize(st));
    SUB_TRY(cudaGetLastError());
    if (bad) { cleanup(); return sub_fail(LAGASSIGN_ERR_ARG, "sub_topic holds an index outside [0, n_topics)"); }
    out->n_csets = totals[0];
    SUB_TRY(cudaMemcpyAsync(out->topic_cset, d_topic_cset.p, (size_t)T * 4, cudaMemcpyDeviceToHost, st));
    SUB_TRY(cudaMemcpyAsync(out->cset_off, d_cset_off.p, ((size_t)totals[0] + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (totals[1] > 0) SUB_TRY(cudaMemcpyAsync(out->cset_rank, d_cset_rank.p, (size_t)totals[1] * 4, cudaMemcpyDeviceToHost, st));
    SUB_TRY(cudaStreamSynchronize(st));
    cleanup();
    return LAGASSIGN_OK;
}
