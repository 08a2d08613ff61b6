// lagassign_api.cu — the C ABI of include/lagassign.h: context, planning (binning
// topics by size on the host from the offsets array), transfers and kernel
// launches. Replaces the call at LagBasedPartitionAssignor.java:149 (the static
// assign(Map, Map) of :166-188). There is no CPU path in this file: every result
// is produced by the kernels in kernels_small.cuh / radix.cuh.
#include "../../include/lagassign.h"

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "types.cuh"

using namespace lagassign;

namespace {

thread_local std::string g_thread_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// elementwise computePartitionLag (LagBasedPartitionAssignor.java:376-404); only lagassign_download_lags() launches
// it (the assignment kernels evaluate lag_from_offsets themselves while building their keys)
__global__ void __launch_bounds__(256)
k_compute_lag(int64_t n, const int64_t* __restrict__ begin_off, const int64_t* __restrict__ end_off,
              const int64_t* __restrict__ committed, const uint8_t* __restrict__ has_committed,
              int reset_latest, int64_t* __restrict__ lag) {
    const LagOffsets o{begin_off, end_off, committed, has_committed, reset_latest};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        lag[i] = lag_from_offsets(o, i);
}

}  // namespace

enum KernelKind { K_COMPUTE_LAG = 0, K_UNASSIGNED, K_WARP, K_CTA, K_ORDER, K_RANGE, K_SCAN, K_TILE_HIST, K_PASS_SCANS,
                  K_SCATTER, K_VERIFY, K_FINALIZE, K_FILL, K_GROUP, K_KINDS };
static const char* const kKernelNames[K_KINDS] = {"k_compute_lag", "k_unassigned_topics", "k_warp_topics",
                                                  "k_cta_topics", "k_big_order_check", "k_big_range",
                                                  "k_big_scan", "k_big_tile_hist", "k_big_seg_scan+k_big_topic_scan",
                                                  "k_big_scatter", "k_big_verify", "k_big_finalize", "k_big_fill",
                                                  "k_group_keys+finalize+offsets"};
struct ProfRec { int kind; cudaEvent_t a, b; };

struct lagassign_ctx {
    bool profile = false;
    std::vector<ProfRec> prof;        // event pairs of the last execute()
    std::vector<cudaEvent_t> ev_pool; // reusable events
    size_t ev_used = 0;
    float prof_ms[K_KINDS] = {0};
    int prof_n[K_KINDS] = {0};
    bool prof_resolved = true;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    std::string err;
    bool want_totals = false;
    // radix ranking: fast form certified by k_big_verify (false) or exact ballot form (true), see radix.cuh 4d
    bool exact_rank = false;
    int64_t rank_fallbacks = 0;

    // plan of the uploaded batch
    bool uploaded = false, executed = false, has_part = false, has_offsets = false;
    int reset_latest = 0;
    int64_t T = 0, N = 0, S = 0, R = 0, total_tc = 0, n_assigned = 0;
    int n_warp = 0, n_cta[N_CTA_CLASSES] = {0, 0, 0, 0}, n_none = 0, n_big = 0, n_tiles = 0, n_fill = 0;
    int64_t big_elems = 0;

    // inputs in HBM
    DevBuf d_topic_off, d_lag, d_part, d_topic_cset, d_cset_off, d_cset_rank, d_tc_off;
    DevBuf d_begin, d_end, d_committed, d_hascommitted;
    // plan lists
    DevBuf d_list_warp, d_list_cta[N_CTA_CLASSES], d_list_none, d_list_fill;
    DevBuf d_bt_topic, d_bt_off, d_bt_P, d_bt_boff, d_bt_tile0, d_bt_seg0, d_tile_bt, d_seg_bt, d_seg_total, d_tile_desc;
    int n_segs = 0;
    int64_t fill_ws_elems = 0;
    // radix workspaces
    DevBuf d_fill_ws_off, d_fill_ws_key, d_fill_ws_idx;
    DevBuf d_keyA, d_keyB, d_payA, d_payB, d_dig, d_range, d_pass_rec, d_sel, d_need_tie, d_status, d_ticket;
    // outputs
    DevBuf d_out_part, d_out_cons, d_out_total, d_out_count;
    int32_t* bound_part = nullptr;
    int32_t* bound_cons = nullptr;

    // host copies of the bin lists (ascending topic ids) for the chunked assign_batch() pipeline
    std::vector<int32_t> h_list_warp, h_list_cta[N_CTA_CLASSES], h_list_none;
    cudaStream_t stream_in = nullptr, stream_out = nullptr;
    std::vector<cudaEvent_t> pipe_events;
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};   // execute(): size bins on concurrent streams
    int64_t pipe_chunk_elems = 0;   // 0 = default (LAGASSIGN_OPT_PIPE_CHUNK)
    // narrow result form (lagassign_assign_batch_narrow): widths this batch's shape allows (set with the plan),
    // widths imposed by a multi-device parent (0 = none), and the device staging buffers
    int64_t plan_max_p = 0;
    bool plan_multi_round = false;
    int forced_pw = 0, forced_cw = -1;
    DevBuf d_narrow_part, d_narrow_cons;
    // Plan cache: a consumer group is rebalanced again and again over the same topics and
    // subscriptions with fresh lags, so the batch SHAPE (everything but lag / partition ids /
    // offsets) usually repeats. When it does, the plan and the metadata already on the device
    // are reused and only the per-partition data is uploaded.
    bool plan_cache = true;         // LAGASSIGN_OPT_PLAN_CACHE
    struct PlanKey {
        bool valid = false;
        bool has_part = false, has_offsets = false, want_totals = false, want_group = false, sort_only = false;
        std::vector<int64_t> topic_off, cset_off;
        std::vector<int32_t> topic_cset, cset_rank;
        lagassign_stats stats{};
    } plan_key;
    bool trace = false;             // LAGASSIGN_TRACE=1: host-side phase times of assign_batch on stderr

    // grouping of the result by member (LAGASSIGN_OPT_GROUP): an internal child context runs the
    // radix sort of the flat result by member rank (sort_only = the child's mode)
    bool want_group = false, sort_only = false;
    lagassign_ctx* grp = nullptr;
    int64_t n_members = 0;
    DevBuf d_grp_topic, d_grp_part, d_member_start, d_member_off;

    lagassign_stats stats{};

    // lagassign_create_multi(): a parent context owns one child context per device and no device state of its own;
    // topics are split into contiguous ranges, one per child (see the "multi-device" section at the end of this file)
    std::vector<lagassign_ctx*> shards;
    struct ShardRange { int64_t t0 = 0, t1 = 0, e0 = 0, e1 = 0, tc0 = 0; std::vector<int64_t> topic_off; };
    std::vector<ShardRange> shard_ranges;
};

namespace {

// Where a call's results go: the int32 arrays of lagassign_result (pw = cw = 4, copied as they are) or the narrow
// form of lagassign_result_narrow (entries of pw / cw bytes, packed on the device before they cross PCIe).
struct OutSpec {
    int pw = 4, cw = 4;
    bool narrow = false;
    uint8_t* h_part = nullptr;     // host destinations; null = not wanted
    uint8_t* h_cons = nullptr;
    int64_t* h_total = nullptr;
    int32_t* h_count = nullptr;
};

OutSpec spec_of(const lagassign_result* out) {
    OutSpec o;
    o.h_part = reinterpret_cast<uint8_t*>(out->out_partition);
    o.h_cons = reinterpret_cast<uint8_t*>(out->out_consumer_rank);
    o.h_total = out->out_total;
    o.h_count = out->out_count;
    return o;
}

// widths of the narrow form for a planned batch (include/lagassign.h, lagassign_assign_batch_narrow)
void narrow_widths_of(bool has_part, int64_t max_p, bool multi_round, int64_t n_members, int* pw, int* cw) {
    *pw = has_part ? 4 : (max_p <= 256 ? 1 : (max_p <= 65536 ? 2 : 4));
    *cw = !multi_round ? 0 : (n_members < 65535 ? 2 : 4);      // 0xFFFF stands for "no subscriber" (-1)
}

int fail(lagassign_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_thread_error = buf;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(ctx, _e == cudaErrorMemoryAllocation ? LAGASSIGN_ERR_OOM : LAGASSIGN_ERR_CUDA, \
                        "%s failed: %s", #expr, cudaGetErrorString(_e));                      \
    } while (0)

template <class T>
int upload_vec(lagassign_ctx* ctx, DevBuf& buf, const T* host, size_t n) {
    CUDA_TRY(ctx, buf.ensure(std::max<size_t>(n, 1) * sizeof(T)));
    if (n) {
        CUDA_TRY(ctx, cudaMemcpyAsync(buf.p, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        ctx->stats.h2d_bytes += (int64_t)(n * sizeof(T));
    }
    return LAGASSIGN_OK;
}

#define TRY(expr) do { int _rc = (expr); if (_rc != LAGASSIGN_OK) return _rc; } while (0)

cudaEvent_t prof_event(lagassign_ctx* ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[ctx->ev_used++];
}

struct ProfScope {   // records an event pair around one launch when profiling is on
    lagassign_ctx* ctx;
    ProfRec rec;
    ProfScope(lagassign_ctx* c, int kind) : ctx(c) {        // c == nullptr: never records
        if (!ctx || !ctx->profile) return;
        rec.kind = kind; rec.a = prof_event(ctx); rec.b = prof_event(ctx);
        cudaEventRecord(rec.a, ctx->stream);
    }
    ~ProfScope() {
        if (!ctx || !ctx->profile) return;
        cudaEventRecord(rec.b, ctx->stream);
        ctx->prof.push_back(rec);
    }
};

int validate_and_plan(lagassign_ctx* ctx, const lagassign_batch* in,
                      std::vector<int32_t>& l_warp, std::vector<int32_t> (&l_cta)[N_CTA_CLASSES], std::vector<int32_t>& l_none,
                      std::vector<int32_t>& bt_topic, std::vector<int64_t>& bt_off, std::vector<int64_t>& bt_P,
                      std::vector<int64_t>& bt_boff, std::vector<int32_t>& bt_tile0, std::vector<int32_t>& tile_bt,
                      std::vector<int32_t>& l_fill, std::vector<int64_t>& tc_off, std::vector<int32_t>& bt_seg0,
                      std::vector<int32_t>& seg_bt, std::vector<TileDesc>& tile_desc, std::vector<int64_t>& fill_ws_off) {
    if (!in) return fail(ctx, LAGASSIGN_ERR_ARG, "batch is NULL");
    const int64_t T = in->n_topics, S = in->n_csets;
    if (T < 0 || S < 0) return fail(ctx, LAGASSIGN_ERR_ARG, "negative n_topics / n_csets");
    if (T > 0x7FFFFFFF) return fail(ctx, LAGASSIGN_ERR_UNSUPPORTED, "more than 2^31-1 topics");
    if (!in->topic_off || !in->cset_off) return fail(ctx, LAGASSIGN_ERR_ARG, "topic_off / cset_off is NULL");
    if (in->topic_off[0] != 0 || in->cset_off[0] != 0) return fail(ctx, LAGASSIGN_ERR_ARG, "offsets must start at 0");
    for (int64_t s = 0; s < S; ++s) {
        if (in->cset_off[s + 1] < in->cset_off[s]) return fail(ctx, LAGASSIGN_ERR_ARG, "cset_off not monotone at %lld", (long long)s);
    }
    const int64_t R = in->cset_off[S];
    int64_t n_members = 0;
    if (R > 0 && !in->cset_rank) return fail(ctx, LAGASSIGN_ERR_ARG, "cset_rank is NULL");
    for (int64_t s = 0; s < S; ++s) {
        for (int64_t j = in->cset_off[s]; j < in->cset_off[s + 1]; ++j) {
            if (in->cset_rank[j] < 0) return fail(ctx, LAGASSIGN_ERR_ARG, "negative member rank in set %lld", (long long)s);
            n_members = std::max<int64_t>(n_members, (int64_t)in->cset_rank[j] + 1);
            if (j > in->cset_off[s] && in->cset_rank[j] <= in->cset_rank[j - 1])
                return fail(ctx, LAGASSIGN_ERR_ARG, "member ranks of set %lld are not strictly ascending", (long long)s);
        }
    }
    for (int64_t t = 0; t < T; ++t) {
        if (in->topic_off[t + 1] < in->topic_off[t]) return fail(ctx, LAGASSIGN_ERR_ARG, "topic_off not monotone at %lld", (long long)t);
    }
    const int64_t N = in->topic_off[T];
    if (N > 0 && !in->lag) return fail(ctx, LAGASSIGN_ERR_ARG, "lag is NULL");
    if (T > 0 && !in->topic_cset) return fail(ctx, LAGASSIGN_ERR_ARG, "topic_cset is NULL");

    tc_off.assign((size_t)T + 1, 0);
    ctx->n_segs = 0;
    ctx->fill_ws_elems = 0;
    int64_t boff = 0, tiles = 0, assigned = 0, max_p = 0;
    bool multi_round = false;
    for (int64_t t = 0; t < T; ++t) {
        const int32_t s = in->topic_cset[t];
        if (s < 0 || s >= S) return fail(ctx, LAGASSIGN_ERR_ARG, "topic_cset[%lld] out of range", (long long)t);
        const int64_t P = in->topic_off[t + 1] - in->topic_off[t];
        const int64_t C = in->cset_off[s + 1] - in->cset_off[s];
        tc_off[(size_t)t + 1] = tc_off[(size_t)t] + C;
        max_p = std::max(max_p, P);
        multi_round = multi_round || (C > 0 && P > C);
        if (P == 0) continue;
        if (C == 0 && !ctx->sort_only) { l_none.push_back((int32_t)t); continue; }
        assigned += P;
        if (P <= WARP_MAX_P && !ctx->sort_only) l_warp.push_back((int32_t)t);
        else if (P <= CTA_MAX_P && !ctx->sort_only) l_cta[cta_class_of(P)].push_back((int32_t)t);
        else {
            if (P > RT_MAX_P) return fail(ctx, LAGASSIGN_ERR_UNSUPPORTED, "topic %lld has 2^31 or more partitions", (long long)t);
            const int64_t nt = (P + RT_TILE - 1) / RT_TILE;
            if (tiles + nt > 0x7FFFFFFF) return fail(ctx, LAGASSIGN_ERR_UNSUPPORTED, "too many radix tiles");
            const int32_t b = (int32_t)bt_topic.size();
            bt_topic.push_back((int32_t)t);
            bt_off.push_back(in->topic_off[t]);
            bt_P.push_back(P);
            bt_boff.push_back(boff);
            bt_tile0.push_back((int32_t)tiles);
            bt_seg0.push_back((int32_t)ctx->n_segs);
            const int nseg = (int)((nt + RT_SEG - 1) / RT_SEG);
            for (int k = 0; k < nseg; ++k) seg_bt.push_back(b);
            ctx->n_segs += nseg;
            for (int64_t k = 0; k < nt; ++k) {
                tile_bt.push_back(b);
                TileDesc d;
                d.b = b;
                d.tile_in_topic = (int32_t)k;
                d.n_valid = (int32_t)std::min<int64_t>(RT_TILE, P - k * RT_TILE);
                d.seg = (int32_t)(bt_seg0.back() + k / RT_SEG);
                d.buf_off = boff + k * RT_TILE;
                d.raw_off = in->topic_off[t] + k * RT_TILE;
                tile_desc.push_back(d);
            }
            if (P > C) l_fill.push_back(b);
            // more consumers than the on-chip fills hold (and still more partitions): a global workspace of pow2 >= C slots
            int64_t ws = -1;
            if (P > C && C > FILL_CTA_MAX_C) {
                int64_t nc = 1;
                while (nc < C) nc <<= 1;
                ws = ctx->fill_ws_elems;
                ctx->fill_ws_elems += nc;
            }
            fill_ws_off.push_back(ws);
            boff += (P + RT_BUF_ALIGN - 1) / RT_BUF_ALIGN * RT_BUF_ALIGN;
            tiles += nt;
        }
    }
    ctx->T = T; ctx->N = N; ctx->S = S; ctx->R = R;
    ctx->n_members = n_members;
    ctx->plan_max_p = max_p;
    ctx->plan_multi_round = multi_round;
    ctx->total_tc = tc_off[(size_t)T];
    ctx->n_assigned = assigned;
    ctx->big_elems = boff;
    return LAGASSIGN_OK;
}

DevBatch make_dev_batch(lagassign_ctx* ctx) {
    DevBatch b{};
    b.topic_off = ctx->d_topic_off.as<int64_t>();
    b.lag = ctx->d_lag.as<int64_t>();
    if (ctx->has_offsets) {
        b.offs.begin = ctx->d_begin.as<int64_t>();
        b.offs.end = ctx->d_end.as<int64_t>();
        b.offs.committed = ctx->d_committed.as<int64_t>();
        b.offs.has = ctx->d_hascommitted.as<uint8_t>();
        b.offs.reset_latest = ctx->reset_latest;
    }
    b.part = ctx->has_part ? ctx->d_part.as<int32_t>() : nullptr;
    b.topic_cset = ctx->d_topic_cset.as<int32_t>();
    b.cset_off = ctx->d_cset_off.as<int64_t>();
    b.cset_rank = ctx->d_cset_rank.as<int32_t>();
    b.tc_off = ctx->d_tc_off.as<int64_t>();
    b.out_part = ctx->bound_part ? ctx->bound_part : ctx->d_out_part.as<int32_t>();
    b.out_cons = ctx->bound_cons ? ctx->bound_cons : ctx->d_out_cons.as<int32_t>();
    b.out_total = ctx->want_totals ? ctx->d_out_total.as<int64_t>() : nullptr;
    b.out_count = ctx->want_totals ? ctx->d_out_count.as<int32_t>() : nullptr;
    return b;
}

BigTopics make_big(lagassign_ctx* ctx) {
    BigTopics bt{};
    bt.n_big = ctx->n_big;
    bt.n_tiles = ctx->n_tiles;
    bt.topic = ctx->d_bt_topic.as<int32_t>();
    bt.off = ctx->d_bt_off.as<int64_t>();
    bt.P = ctx->d_bt_P.as<int64_t>();
    bt.boff = ctx->d_bt_boff.as<int64_t>();
    bt.tile0 = ctx->d_bt_tile0.as<int32_t>();
    bt.tile_bt = ctx->d_tile_bt.as<int32_t>();
    bt.tile_desc = ctx->d_tile_desc.as<int4>();
    bt.keyA = ctx->d_keyA.as<uint64_t>(); bt.keyB = ctx->d_keyB.as<uint64_t>();
    bt.payA = ctx->d_payA.as<uint32_t>(); bt.payB = ctx->d_payB.as<uint32_t>();
    bt.range = ctx->d_range.as<uint64_t>();
    bt.pass_rec = ctx->d_pass_rec.as<uint32_t>();
    int8_t* sel = ctx->d_sel.as<int8_t>();
    bt.sel_src = sel;
    bt.sel_dst = sel + (size_t)ctx->n_big * N_PASS;
    bt.nxt_pass = sel + (size_t)ctx->n_big * N_PASS * 2;
    bt.sel_final = sel + (size_t)ctx->n_big * N_PASS * 3;
    bt.all_equal = sel + (size_t)ctx->n_big * (N_PASS * 3 + 1);
    bt.dig = ctx->d_dig.as<uint8_t>();
    bt.fill_ws_off = ctx->d_fill_ws_off.as<int64_t>();
    bt.fill_ws_key = ctx->d_fill_ws_key.as<uint64_t>();
    bt.fill_ws_idx = ctx->d_fill_ws_idx.as<uint32_t>();
    bt.need_tie = ctx->d_need_tie.as<int32_t>();
    bt.tile_hist = ctx->d_status.as<uint32_t>();
    bt.seg0 = ctx->d_bt_seg0.as<int32_t>();
    bt.seg_bt = ctx->d_seg_bt.as<int32_t>();
    bt.n_segs = ctx->n_segs;
    bt.seg_total = ctx->d_seg_total.as<uint32_t>();
    bt.ticket = ctx->d_ticket.as<uint32_t>();
    bt.passes_run = reinterpret_cast<int32_t*>(ctx->d_ticket.as<uint32_t>() + N_PASS);
    bt.pass_any = ctx->d_ticket.as<uint32_t>() + N_PASS + 1;
    bt.verify_fail = ctx->d_ticket.as<uint32_t>() + 2 * N_PASS + 1;
    return bt;
}

// Same shape as the batch the current plan was made for (see lagassign_ctx::PlanKey)?
bool plan_matches(const lagassign_ctx* ctx, const lagassign_batch* in, bool has_offsets) {
    const auto& k = ctx->plan_key;
    if (!ctx->plan_cache || !k.valid || !in || !in->topic_off || !in->cset_off) return false;
    if (k.has_part != (in->partition != nullptr) || k.has_offsets != has_offsets || k.want_totals != ctx->want_totals ||
        k.want_group != ctx->want_group || k.sort_only != ctx->sort_only)
        return false;
    if (in->n_topics < 0 || (size_t)in->n_topics + 1 != k.topic_off.size()) return false;
    if (in->n_csets < 0 || (size_t)in->n_csets + 1 != k.cset_off.size()) return false;
    if (!k.topic_cset.empty() && !in->topic_cset) return false;
    if (!k.cset_rank.empty() && !in->cset_rank) return false;
    if (memcmp(in->cset_off, k.cset_off.data(), k.cset_off.size() * sizeof(int64_t)) != 0) return false;   // R follows
    return memcmp(in->topic_off, k.topic_off.data(), k.topic_off.size() * sizeof(int64_t)) == 0 &&
           (k.topic_cset.empty() || memcmp(in->topic_cset, k.topic_cset.data(), k.topic_cset.size() * sizeof(int32_t)) == 0) &&
           (k.cset_rank.empty() || memcmp(in->cset_rank, k.cset_rank.data(), k.cset_rank.size() * sizeof(int32_t)) == 0);
}

int do_upload(lagassign_ctx* ctx, const lagassign_batch* in, const lagassign_offsets* offs, bool defer_data = false) {
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    if (!in) return fail(ctx, LAGASSIGN_ERR_ARG, "batch is NULL");
    ctx->uploaded = ctx->executed = false;
    ctx->stats = lagassign_stats{};
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<int32_t> l_warp, l_cta[N_CTA_CLASSES], l_none, bt_topic, bt_tile0, tile_bt, l_fill, bt_seg0, seg_bt;
    std::vector<TileDesc> tile_desc;
    std::vector<int64_t> bt_off, bt_P, bt_boff, tc_off, fill_ws_off;
    lagassign_batch tmp;
    if (offs) {
        // lags are produced on the device; planning only needs the shapes
        if (!offs->begin_offset || !offs->end_offset || !offs->committed || !offs->has_committed)
            return fail(ctx, LAGASSIGN_ERR_ARG, "offset arrays must not be NULL");
        tmp = *in;
        static const int64_t dummy = 0;
        tmp.lag = &dummy;   // not dereferenced by the planner
        in = &tmp;
    }
    if (plan_matches(ctx, in, offs != nullptr)) {
        const size_t N = (size_t)ctx->N;
        if (N > 0 && !in->lag) return fail(ctx, LAGASSIGN_ERR_ARG, "lag is NULL");
        ctx->stats = ctx->plan_key.stats;
        ctx->stats.h2d_bytes = 0;        // only what this call copies
        if (offs) {
            ctx->reset_latest = offs->reset_latest ? 1 : 0;
            TRY(upload_vec(ctx, ctx->d_begin, offs->begin_offset, N));
            TRY(upload_vec(ctx, ctx->d_end, offs->end_offset, N));
            TRY(upload_vec(ctx, ctx->d_committed, offs->committed, N));
            TRY(upload_vec(ctx, ctx->d_hascommitted, offs->has_committed, N));
        } else if (!defer_data) {
            TRY(upload_vec(ctx, ctx->d_lag, in->lag, N));
        }
        if (ctx->has_part && !defer_data) TRY(upload_vec(ctx, ctx->d_part, in->partition, N));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->grp) ctx->grp->uploaded = true;
        ctx->uploaded = !defer_data;
        return LAGASSIGN_OK;
    }
    ctx->plan_key.valid = false;
    TRY(validate_and_plan(ctx, in, l_warp, l_cta, l_none, bt_topic, bt_off, bt_P, bt_boff, bt_tile0, tile_bt, l_fill, tc_off, bt_seg0, seg_bt, tile_desc, fill_ws_off));
    const int64_t T = ctx->T, N = ctx->N;
    ctx->has_part = in->partition != nullptr;
    ctx->has_offsets = offs != nullptr;
    ctx->n_warp = (int)l_warp.size(); ctx->n_none = (int)l_none.size();
    for (int c = 0; c < N_CTA_CLASSES; ++c) ctx->n_cta[c] = (int)l_cta[c].size();
    ctx->n_big = (int)bt_topic.size(); ctx->n_tiles = (int)tile_bt.size(); ctx->n_fill = (int)l_fill.size();

    TRY(upload_vec(ctx, ctx->d_topic_off, in->topic_off, (size_t)T + 1));
    if (offs) {
        ctx->reset_latest = offs->reset_latest ? 1 : 0;
        TRY(upload_vec(ctx, ctx->d_begin, offs->begin_offset, (size_t)N));
        TRY(upload_vec(ctx, ctx->d_end, offs->end_offset, (size_t)N));
        TRY(upload_vec(ctx, ctx->d_committed, offs->committed, (size_t)N));
        TRY(upload_vec(ctx, ctx->d_hascommitted, offs->has_committed, (size_t)N));
        CUDA_TRY(ctx, ctx->d_lag.ensure(std::max<size_t>((size_t)N, 1) * sizeof(int64_t)));
    } else if (defer_data) {
        CUDA_TRY(ctx, ctx->d_lag.ensure(std::max<size_t>((size_t)N, 1) * sizeof(int64_t)));
        if (ctx->has_part) CUDA_TRY(ctx, ctx->d_part.ensure(std::max<size_t>((size_t)N, 1) * sizeof(int32_t)));
    } else {
        TRY(upload_vec(ctx, ctx->d_lag, in->lag, (size_t)N));
    }
    if (ctx->has_part && !defer_data) TRY(upload_vec(ctx, ctx->d_part, in->partition, (size_t)N));
    TRY(upload_vec(ctx, ctx->d_topic_cset, in->topic_cset, (size_t)T));
    TRY(upload_vec(ctx, ctx->d_cset_off, in->cset_off, (size_t)ctx->S + 1));
    TRY(upload_vec(ctx, ctx->d_cset_rank, in->cset_rank, (size_t)ctx->R));
    TRY(upload_vec(ctx, ctx->d_tc_off, tc_off.data(), tc_off.size()));
    TRY(upload_vec(ctx, ctx->d_list_warp, l_warp.data(), l_warp.size()));
    for (int c = 0; c < N_CTA_CLASSES; ++c) TRY(upload_vec(ctx, ctx->d_list_cta[c], l_cta[c].data(), l_cta[c].size()));
    TRY(upload_vec(ctx, ctx->d_list_none, l_none.data(), l_none.size()));
    TRY(upload_vec(ctx, ctx->d_list_fill, l_fill.data(), l_fill.size()));
    TRY(upload_vec(ctx, ctx->d_bt_topic, bt_topic.data(), bt_topic.size()));
    TRY(upload_vec(ctx, ctx->d_bt_off, bt_off.data(), bt_off.size()));
    TRY(upload_vec(ctx, ctx->d_bt_P, bt_P.data(), bt_P.size()));
    TRY(upload_vec(ctx, ctx->d_bt_boff, bt_boff.data(), bt_boff.size()));
    TRY(upload_vec(ctx, ctx->d_bt_tile0, bt_tile0.data(), bt_tile0.size()));
    TRY(upload_vec(ctx, ctx->d_bt_seg0, bt_seg0.data(), bt_seg0.size()));
    TRY(upload_vec(ctx, ctx->d_seg_bt, seg_bt.data(), seg_bt.size()));
    TRY(upload_vec(ctx, ctx->d_tile_desc, tile_desc.data(), tile_desc.size()));
    TRY(upload_vec(ctx, ctx->d_tile_bt, tile_bt.data(), tile_bt.size()));
    TRY(upload_vec(ctx, ctx->d_fill_ws_off, fill_ws_off.data(), fill_ws_off.size()));

    // outputs and workspaces
    if (!ctx->sort_only) {
        CUDA_TRY(ctx, ctx->d_out_part.ensure(std::max<size_t>((size_t)N, 1) * sizeof(int32_t)));
        CUDA_TRY(ctx, ctx->d_out_cons.ensure(std::max<size_t>((size_t)N, 1) * sizeof(int32_t)));
    }
    if (ctx->want_group && N > 0) {
        // the flat result as ONE topic of N elements, sorted by the child context
        if (N > RT_MAX_P) return fail(ctx, LAGASSIGN_ERR_UNSUPPORTED, "grouping needs fewer than 2^31 partitions per batch");
        if (!ctx->grp) {
            lagassign_ctx* child = new (std::nothrow) lagassign_ctx();
            if (!child) return fail(ctx, LAGASSIGN_ERR_OOM, "out of host memory");
            child->device = ctx->device; child->stream = ctx->stream; child->own_stream = false;
            child->sm_count = ctx->sm_count; child->sort_only = true;
            ctx->grp = child;
        }
        const int64_t g_topic_off[2] = {0, N};
        const int64_t g_cset_off[2] = {0, 0};
        const int32_t g_topic_cset[1] = {0};
        static const int64_t dummy_lag = 0;
        lagassign_batch gb{};
        gb.n_topics = 1; gb.topic_off = g_topic_off; gb.lag = &dummy_lag; gb.partition = nullptr;
        gb.topic_cset = g_topic_cset; gb.n_csets = 1; gb.cset_off = g_cset_off; gb.cset_rank = nullptr;
        int rc = do_upload(ctx->grp, &gb, nullptr, /*defer_data=*/true);      // plan only: its lags are written on the device
        if (rc != LAGASSIGN_OK) return fail(ctx, rc, "grouping plan: %s", ctx->grp->err.c_str());
        ctx->grp->uploaded = true;
        CUDA_TRY(ctx, ctx->d_grp_topic.ensure(std::max<size_t>((size_t)ctx->n_assigned, 1) * sizeof(int32_t)));
        CUDA_TRY(ctx, ctx->d_grp_part.ensure(std::max<size_t>((size_t)ctx->n_assigned, 1) * sizeof(int32_t)));
        CUDA_TRY(ctx, ctx->d_member_start.ensure((size_t)(ctx->n_members + 1) * sizeof(int64_t)));
        CUDA_TRY(ctx, ctx->d_member_off.ensure((size_t)(ctx->n_members + 1) * sizeof(int64_t)));
    }
    if (ctx->want_totals) {
        CUDA_TRY(ctx, ctx->d_out_total.ensure(std::max<size_t>((size_t)ctx->total_tc, 1) * sizeof(int64_t)));
        CUDA_TRY(ctx, ctx->d_out_count.ensure(std::max<size_t>((size_t)ctx->total_tc, 1) * sizeof(int32_t)));
    }
    if (ctx->n_big > 0) {
        const size_t nb = (size_t)ctx->big_elems;
        CUDA_TRY(ctx, ctx->d_keyA.ensure(nb * sizeof(uint64_t)));
        CUDA_TRY(ctx, ctx->d_keyB.ensure(nb * sizeof(uint64_t)));
        CUDA_TRY(ctx, ctx->d_payA.ensure(nb * sizeof(uint32_t)));
        CUDA_TRY(ctx, ctx->d_payB.ensure(nb * sizeof(uint32_t)));
        CUDA_TRY(ctx, ctx->d_range.ensure((size_t)ctx->n_big * 4 * sizeof(uint64_t)));
        CUDA_TRY(ctx, ctx->d_pass_rec.ensure((size_t)ctx->n_tiles * N_PASS * sizeof(uint32_t)));
        CUDA_TRY(ctx, ctx->d_sel.ensure((size_t)ctx->n_big * (3 * N_PASS + 2)));
        CUDA_TRY(ctx, ctx->d_dig.ensure(nb));
        CUDA_TRY(ctx, ctx->d_fill_ws_key.ensure(std::max<size_t>((size_t)ctx->fill_ws_elems, 1) * sizeof(uint64_t)));
        CUDA_TRY(ctx, ctx->d_fill_ws_idx.ensure(std::max<size_t>((size_t)ctx->fill_ws_elems, 1) * sizeof(uint32_t)));
        CUDA_TRY(ctx, ctx->d_need_tie.ensure((size_t)ctx->n_big * sizeof(int32_t)));
        CUDA_TRY(ctx, ctx->d_status.ensure((size_t)ctx->n_tiles * 256 * sizeof(uint32_t)));       // tile_hist
        CUDA_TRY(ctx, ctx->d_seg_total.ensure((size_t)ctx->n_segs * 256 * sizeof(uint32_t)));
        CUDA_TRY(ctx, ctx->d_ticket.ensure(RT_TICKET_WORDS * sizeof(uint32_t)));
    }
    // the host vectors above must outlive the copies; also honours "the library
    // copies what it needs before returning"
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.n_partitions = N;
    ctx->stats.n_assigned = ctx->n_assigned;
    ctx->stats.topics_warp = ctx->n_warp;
    ctx->stats.topics_cta = (int64_t)ctx->n_cta[0] + ctx->n_cta[1] + ctx->n_cta[2] + ctx->n_cta[3];
    ctx->stats.topics_device = ctx->n_big;
    ctx->stats.topics_unassigned = ctx->n_none;
    ctx->stats.device_sort_elems = ctx->big_elems;
    ctx->h_list_warp.swap(l_warp);
    for (int c = 0; c < N_CTA_CLASSES; ++c) ctx->h_list_cta[c].swap(l_cta[c]);
    ctx->h_list_none.swap(l_none);
    ctx->uploaded = !defer_data;
    {
        auto& k = ctx->plan_key;
        k.has_part = ctx->has_part; k.has_offsets = ctx->has_offsets;
        k.want_totals = ctx->want_totals; k.want_group = ctx->want_group; k.sort_only = ctx->sort_only;
        k.topic_off.assign(in->topic_off, in->topic_off + T + 1);
        k.cset_off.assign(in->cset_off, in->cset_off + ctx->S + 1);
        k.topic_cset.assign(in->topic_cset, in->topic_cset + T);
        k.cset_rank.assign(in->cset_rank, in->cset_rank + ctx->R);
        k.stats = ctx->stats;
        k.valid = true;
    }
    return LAGASSIGN_OK;
}

// Upload of the two large arrays when do_upload() was asked to defer them.
int upload_data_now(lagassign_ctx* ctx, const lagassign_batch* in) {
    const size_t N = (size_t)ctx->N;
    if (N) {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_lag.p, in->lag, N * 8, cudaMemcpyHostToDevice, ctx->stream));
        ctx->stats.h2d_bytes += (int64_t)N * 8;
        if (ctx->has_part) {
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_part.p, in->partition, N * 4, cudaMemcpyHostToDevice, ctx->stream));
            ctx->stats.h2d_bytes += (int64_t)N * 4;
        }
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->uploaded = true;
    return LAGASSIGN_OK;
}

// Kernels of the on-chip topics with index in [t_lo, t_hi) (the per-bin topic lists are ascending, so a topic range is
// a slice of each); returns the number of launches. Used by the copy/compute pipeline and by lagassign_execute_range.
int launch_topic_range(lagassign_ctx* ctx, const DevBatch& b, int64_t t_lo, int64_t t_hi, bool profile) {
    auto slice = [&](const std::vector<int32_t>& v, int& lo, int& hi) {
        lo = (int)(std::lower_bound(v.begin(), v.end(), (int32_t)t_lo) - v.begin());
        hi = (int)(std::lower_bound(v.begin(), v.end(), (int32_t)t_hi) - v.begin());
    };
    int lo, hi, launches = 0;
    slice(ctx->h_list_none, lo, hi);
    if (hi > lo) {
        ProfScope ps(profile ? ctx : nullptr, K_UNASSIGNED);
        launch_unassigned_topics(b, ctx->d_list_none.as<int32_t>() + lo, hi - lo, ctx->stream); ++launches;
    }
    slice(ctx->h_list_warp, lo, hi);
    if (hi > lo) {
        ProfScope ps(profile ? ctx : nullptr, K_WARP);
        launch_warp_topics(b, ctx->d_list_warp.as<int32_t>() + lo, hi - lo, ctx->stream); ++launches;
    }
    for (int c2 = 0; c2 < N_CTA_CLASSES; ++c2) {
        slice(ctx->h_list_cta[c2], lo, hi);
        if (hi > lo) {
            ProfScope ps(profile ? ctx : nullptr, K_CTA);
            launch_cta_topics(b, ctx->d_list_cta[c2].as<int32_t>() + lo, hi - lo, c2, ctx->sm_count, ctx->stream); ++launches;
        }
    }
    return launches;
}

constexpr int64_t PIPE_MIN_ELEMS = 1 << 21;      // below this one copy each way is as good
constexpr int64_t PIPE_CHUNK_ELEMS = 3 << 20;    // ~24 MB of lags per chunk

// lagassign_assign_batch for batches made of on-chip topics only: the batch is cut
// into contiguous topic ranges and the three legs (H2D of the lags, kernels, D2H of
// the results) of successive chunks overlap on three streams. PCIe is full duplex,
// so the call approaches max(H2D, D2H) instead of their sum.
int assign_batch_pipelined_impl(lagassign_ctx* ctx, const lagassign_batch* in, const OutSpec& out, std::vector<cudaEvent_t>& tev) {
    if (!ctx->stream_in) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream_in, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream_out, cudaStreamNonBlocking));
    }
    const int64_t N = ctx->N, T = ctx->T;
    // Chunk sizes: a short ramp at both ends (chunk/8, /4, /2) so that the first kernels and
    // the first D2H start early and the last D2H is short; full chunks in between.
    const int64_t chunk = ctx->pipe_chunk_elems > 0 ? ctx->pipe_chunk_elems : PIPE_CHUNK_ELEMS;
    std::vector<int64_t> cuts;      // cumulative partition counts at which chunks end (last = N)
    {
        const int64_t ramp[3] = {std::max<int64_t>(chunk / 8, 1), std::max<int64_t>(chunk / 4, 1), std::max<int64_t>(chunk / 2, 1)};
        const int64_t ramp_sum = ramp[0] + ramp[1] + ramp[2];
        if (N >= 2 * ramp_sum + chunk) {
            int64_t pos = 0;
            for (int i = 0; i < 3; ++i) { pos += ramp[i]; cuts.push_back(pos); }
            const int64_t body = N - 2 * ramp_sum;
            const int64_t nb = std::min<int64_t>(56, (body + chunk - 1) / chunk);
            for (int64_t i = 1; i <= nb; ++i) cuts.push_back(ramp_sum + body / nb * i);
            cuts.back() = ramp_sum + body;
            pos = ramp_sum + body;
            for (int i = 2; i >= 0; --i) { pos += ramp[i]; cuts.push_back(pos); }
        } else {
            const int64_t k = std::min<int64_t>(62, std::max<int64_t>(1, (N + chunk - 1) / chunk));
            for (int64_t i = 1; i <= k; ++i) cuts.push_back(N / k * i);
        }
        cuts.back() = N;
    }
    const int K = (int)cuts.size();
    while ((int)ctx->pipe_events.size() < 2 * K) {
        cudaEvent_t e = nullptr;
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_events.push_back(e);
    }
    // LAGASSIGN_TRACE: timing events at the end of every leg of every chunk (destroyed by the caller)
    if (ctx->trace) {
        tev.assign(1 + 3 * (size_t)K, nullptr);
        for (auto& e : tev) CUDA_TRY(ctx, cudaEventCreate(&e));
        CUDA_TRY(ctx, cudaEventRecord(tev[0], ctx->stream_in));
    }
    DevBatch b = make_dev_batch(ctx);
    const int32_t* src_part = b.out_part;
    const int32_t* src_cons = b.out_cons;
    int launches = 0;
    int64_t t_lo = 0;
    ctx->stats.d2h_bytes = 0;
    for (int c = 0; c < K; ++c) {
        int64_t t_hi = T;
        if (c + 1 < K) {
            const int64_t target = cuts[c];
            t_hi = std::lower_bound(in->topic_off, in->topic_off + T + 1, target) - in->topic_off;
            t_hi = std::min<int64_t>(std::max(t_hi, t_lo), T);
        }
        const int64_t e0 = in->topic_off[t_lo], e1 = in->topic_off[t_hi];
        const size_t n = (size_t)(e1 - e0);
        cudaEvent_t ev_in = ctx->pipe_events[2 * c], ev_out = ctx->pipe_events[2 * c + 1];
        if (n) {
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_lag.as<int64_t>() + e0, in->lag + e0, n * 8, cudaMemcpyHostToDevice, ctx->stream_in));
            ctx->stats.h2d_bytes += (int64_t)n * 8;
            if (ctx->has_part) {
                CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_part.as<int32_t>() + e0, in->partition + e0, n * 4, cudaMemcpyHostToDevice, ctx->stream_in));
                ctx->stats.h2d_bytes += (int64_t)n * 4;
            }
        }
        CUDA_TRY(ctx, cudaEventRecord(ev_in, ctx->stream_in));
        if (ctx->trace) CUDA_TRY(ctx, cudaEventRecord(tev[1 + 3 * c], ctx->stream_in));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ev_in, 0));
        launches += launch_topic_range(ctx, b, t_lo, t_hi, /*profile=*/false);
        const uint8_t* d2h_part = reinterpret_cast<const uint8_t*>(src_part + e0);
        const uint8_t* d2h_cons = reinterpret_cast<const uint8_t*>(src_cons + e0);
        if (n && out.narrow) {
            // the chunk's results in their narrow form (one small kernel behind the chunk's sort kernels)
            uint8_t* np = ctx->d_narrow_part.as<uint8_t>() + e0 * out.pw;
            uint8_t* nc = ctx->d_narrow_cons.as<uint8_t>() + e0 * out.cw;
            if (!launch_narrow_results(ctx->stream, (int64_t)n, src_part + e0, src_cons + e0, out.pw, out.cw, np, nc))
                return fail(ctx, LAGASSIGN_ERR_ARG, "narrow result widths %d / %d", out.pw, out.cw);
            ++launches;
            d2h_part = np; d2h_cons = nc;
        }
        CUDA_TRY(ctx, cudaEventRecord(ev_out, ctx->stream));
        if (ctx->trace) CUDA_TRY(ctx, cudaEventRecord(tev[2 + 3 * c], ctx->stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream_out, ev_out, 0));
        if (n && out.h_part) {
            CUDA_TRY(ctx, cudaMemcpyAsync(out.h_part + e0 * out.pw, d2h_part, n * out.pw, cudaMemcpyDeviceToHost, ctx->stream_out));
            ctx->stats.d2h_bytes += (int64_t)n * out.pw;
        }
        if (n && out.h_cons && out.cw) {
            CUDA_TRY(ctx, cudaMemcpyAsync(out.h_cons + e0 * out.cw, d2h_cons, n * out.cw, cudaMemcpyDeviceToHost, ctx->stream_out));
            ctx->stats.d2h_bytes += (int64_t)n * out.cw;
        }
        if (ctx->trace) CUDA_TRY(ctx, cudaEventRecord(tev[3 + 3 * c], ctx->stream_out));
        t_lo = t_hi;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream_out));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->trace) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream_in));
        fprintf(stderr, "[lagassign] pipeline of %d chunks, ms since the first copy was enqueued: chunk (partitions up to): h2d done / kernels done / d2h done\n", K);
        for (int c = 0; c < K; ++c) {
            float a = 0, k = 0, d = 0;
            cudaEventElapsedTime(&a, tev[0], tev[1 + 3 * c]);
            cudaEventElapsedTime(&k, tev[0], tev[2 + 3 * c]);
            cudaEventElapsedTime(&d, tev[0], tev[3 + 3 * c]);
            fprintf(stderr, "[lagassign]   %2d (%lld): %.3f / %.3f / %.3f\n", c, (long long)cuts[c], a, k, d);
        }
    }
    ctx->stats.kernel_launches = launches;
    ctx->uploaded = true;
    ctx->executed = true;
    ctx->prof.clear();
    ctx->prof_resolved = true;
    return LAGASSIGN_OK;
}

int assign_batch_pipelined(lagassign_ctx* ctx, const lagassign_batch* in, const OutSpec& out) {
    std::vector<cudaEvent_t> tev;
    const int rc = assign_batch_pipelined_impl(ctx, in, out, tev);
    if (rc != LAGASSIGN_OK) {
        // Copies into and out of the CALLER's buffers may still be in flight on the three streams: they must have
        // ended before the error is returned and the caller is free to release those buffers.
        if (ctx->stream_in) cudaStreamSynchronize(ctx->stream_in);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        if (ctx->stream_out) cudaStreamSynchronize(ctx->stream_out);
        cudaGetLastError();
    }
    for (cudaEvent_t e : tev)
        if (e) cudaEventDestroy(e);
    return rc;
}

int configure_kernels(lagassign_ctx* ctx) {
    CUDA_TRY(ctx, configure_cta_kernels());
    CUDA_TRY(ctx, configure_big_kernels());
    return LAGASSIGN_OK;
}

int create_common(lagassign_ctx** out, int device, void* stream, bool own) {
    if (!out) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx out-pointer is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, LAGASSIGN_ERR_NO_DEVICE,
                    "no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(nullptr, LAGASSIGN_ERR_ARG, "device %d out of range [0,%d)", device, n);
    lagassign_ctx* ctx = new (std::nothrow) lagassign_ctx();
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_OOM, "out of host memory");
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return fail(nullptr, LAGASSIGN_ERR_CUDA, "cudaSetDevice(%d) failed", device); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return fail(nullptr, LAGASSIGN_ERR_CUDA, "cudaGetDeviceProperties failed"); }
    ctx->sm_count = prop.multiProcessorCount;
    { const char* tr = getenv("LAGASSIGN_TRACE"); ctx->trace = tr && tr[0] && tr[0] != '0'; }
    if (own) {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return fail(nullptr, LAGASSIGN_ERR_CUDA, "cudaStreamCreate failed"); }
        ctx->own_stream = true;
    } else {
        ctx->stream = reinterpret_cast<cudaStream_t>(stream);
    }
    int rc = configure_kernels(ctx);
    if (rc != LAGASSIGN_OK) { std::string m = ctx->err; lagassign_destroy(ctx); return fail(nullptr, rc, "%s", m.c_str()); }
    {
        // the fast radix ranking relies on lane-ordered shared-memory atomics (certified per sort, radix.cuh 4d):
        // a device that does not behave so uses the exact ranking from the start
        static int probed[64];          // per device: 0 = not yet, 1 = lane order holds, 2 = it does not
        int& pr = probed[device & 63];
        if (pr == 0) {
            int in_order = 0;
            if (probe_atomic_lane_order(&in_order, ctx->stream) != cudaSuccess) { cudaGetLastError(); in_order = 0; }
            pr = in_order ? 1 : 2;
        }
        ctx->exact_rank = (pr == 2);
        const char* ex = getenv("LAGASSIGN_EXACT_RANK");
        if (ex && ex[0] && ex[0] != '0') ctx->exact_rank = true;
    }
    *out = ctx;
    return LAGASSIGN_OK;
}

int multi_set_option(lagassign_ctx* ctx, int option, int64_t value);
int multi_upload(lagassign_ctx* ctx, const lagassign_batch* in, const lagassign_offsets* offs);
int multi_execute(lagassign_ctx* ctx);
int multi_sync(lagassign_ctx* ctx);
int multi_download(lagassign_ctx* ctx, const lagassign_result* out);
int multi_assign_batch(lagassign_ctx* ctx, const lagassign_batch* in, const lagassign_result* out);
int multi_assign_batch_narrow(lagassign_ctx* ctx, const lagassign_batch* in, lagassign_result_narrow* out);

}  // namespace

extern "C" {

int lagassign_abi_version(void) { return LAGASSIGN_ABI_VERSION; }

int lagassign_device_count(int* n_devices) {
    if (!n_devices) return fail(nullptr, LAGASSIGN_ERR_ARG, "n_devices is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *n_devices = n;
    return LAGASSIGN_OK;
}

int lagassign_create(lagassign_ctx** ctx, int device) { return create_common(ctx, device, nullptr, true); }

int lagassign_create_on_stream(lagassign_ctx** ctx, int device, void* cuda_stream) {
    return create_common(ctx, device, cuda_stream, false);
}

int lagassign_set_stream(lagassign_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    if (!ctx->shards.empty()) return fail(ctx, LAGASSIGN_ERR_UNSUPPORTED, "not available on a multi-device context");
    if (ctx->own_stream) return fail(ctx, LAGASSIGN_ERR_STATE, "the context owns its stream (created with lagassign_create)");
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    if (ctx->grp) ctx->grp->stream = ctx->stream;
    return LAGASSIGN_OK;
}

void lagassign_destroy(lagassign_ctx* ctx) {
    if (!ctx) return;
    if (!ctx->shards.empty()) {
        for (lagassign_ctx* c : ctx->shards) lagassign_destroy(c);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->d_topic_off, &ctx->d_lag, &ctx->d_part, &ctx->d_topic_cset, &ctx->d_cset_off,
                      &ctx->d_cset_rank, &ctx->d_tc_off, &ctx->d_begin, &ctx->d_end, &ctx->d_committed,
                      &ctx->d_hascommitted, &ctx->d_list_warp, &ctx->d_list_cta[0], &ctx->d_list_cta[1], &ctx->d_list_cta[2], &ctx->d_list_cta[3], &ctx->d_list_none,
                      &ctx->d_list_fill, &ctx->d_bt_topic, &ctx->d_bt_off, &ctx->d_bt_P, &ctx->d_bt_boff,
                      &ctx->d_bt_tile0, &ctx->d_tile_bt, &ctx->d_keyA, &ctx->d_keyB, &ctx->d_payA,
                      &ctx->d_payB, &ctx->d_dig, &ctx->d_narrow_part, &ctx->d_narrow_cons, &ctx->d_fill_ws_off, &ctx->d_fill_ws_key, &ctx->d_fill_ws_idx, &ctx->d_range, &ctx->d_pass_rec, &ctx->d_sel, &ctx->d_need_tie,
                      &ctx->d_status, &ctx->d_seg_total, &ctx->d_bt_seg0, &ctx->d_seg_bt, &ctx->d_tile_desc, &ctx->d_ticket, &ctx->d_out_part, &ctx->d_out_cons,
                      &ctx->d_out_total, &ctx->d_out_count};
    for (DevBuf* b : bufs) b->release();
    if (ctx->grp) { lagassign_destroy(ctx->grp); ctx->grp = nullptr; }
    DevBuf* gbufs[] = {&ctx->d_grp_topic, &ctx->d_grp_part, &ctx->d_member_start, &ctx->d_member_off};
    for (DevBuf* gb : gbufs) gb->release();
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->pipe_events) cudaEventDestroy(e);
    if (ctx->ev_fork) { cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join[0]); cudaEventDestroy(ctx->ev_join[1]); }
    if (ctx->stream_in) cudaStreamDestroy(ctx->stream_in);
    if (ctx->stream_out) cudaStreamDestroy(ctx->stream_out);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* lagassign_last_error(const lagassign_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_thread_error.c_str();
}

int lagassign_set_option(lagassign_ctx* ctx, int option, int64_t value) {
    if (ctx && !ctx->shards.empty()) return multi_set_option(ctx, option, value);
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    switch (option) {
        case LAGASSIGN_OPT_PROFILE:
            ctx->profile = value != 0;
            return LAGASSIGN_OK;
        case LAGASSIGN_OPT_GROUP:
            ctx->want_group = value != 0;
            ctx->uploaded = false;   // the grouping plan and buffers are set up at upload
            return LAGASSIGN_OK;
        case LAGASSIGN_OPT_PLAN_CACHE:
            ctx->plan_cache = value != 0;
            return LAGASSIGN_OK;
        case LAGASSIGN_OPT_PIPE_CHUNK:
            if (value < 0) return fail(ctx, LAGASSIGN_ERR_ARG, "pipeline chunk size must be >= 0");
            ctx->pipe_chunk_elems = value;
            return LAGASSIGN_OK;
        case LAGASSIGN_OPT_EXACT_RANK:
            ctx->exact_rank = value != 0;
            return LAGASSIGN_OK;
        case LAGASSIGN_OPT_TOTALS:
            ctx->want_totals = value != 0;
            ctx->uploaded = false;   // output buffers are sized at upload
            return LAGASSIGN_OK;
        default:
            return fail(ctx, LAGASSIGN_ERR_ARG, "unknown option %d", option);
    }
}

int lagassign_upload(lagassign_ctx* ctx, const lagassign_batch* in) {
    if (ctx && !ctx->shards.empty()) return multi_upload(ctx, in, nullptr); return do_upload(ctx, in, nullptr); }

int lagassign_upload_offsets(lagassign_ctx* ctx, const lagassign_batch* in, const lagassign_offsets* offs) {
    if (!offs) return fail(ctx, LAGASSIGN_ERR_ARG, "offsets is NULL");
    if (ctx && !ctx->shards.empty()) return multi_upload(ctx, in, offs);
    return do_upload(ctx, in, offs);
}

int lagassign_execute(lagassign_ctx* ctx) {
    if (ctx && !ctx->shards.empty()) return multi_execute(ctx);
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    if (!ctx->uploaded) return fail(ctx, LAGASSIGN_ERR_STATE, "execute() before a successful upload()");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int launches = 0;
    bool verify_pending = false;
    ctx->prof.clear();
    ctx->ev_used = 0;
    ctx->prof_resolved = false;
    DevBatch b = make_dev_batch(ctx);
    const int sms = ctx->sm_count;

    // With raw offsets (lagassign_upload_offsets) there is no separate lag kernel: computePartitionLag (Main:376-404)
    // is evaluated inside the kernels below as they build their sort keys (k_warp_topics<.., LAG_OFFS>,
    // k_cta_topics, k_big_range).
    if (ctx->want_totals && ctx->total_tc > 0) {
        CUDA_TRY(ctx, cudaMemsetAsync(b.out_total, 0, (size_t)ctx->total_tc * sizeof(int64_t), st));
        CUDA_TRY(ctx, cudaMemsetAsync(b.out_count, 0, (size_t)ctx->total_tc * sizeof(int32_t), st));
    }
    // The size bins write disjoint slices of the outputs, so they may run side by side: when
    // more than one bin has work (and per-kernel profiling is off), the warp bin and the CTA
    // bin go to two side streams forked from / joined back into the context's stream, next to
    // the radix chain. Small launches then fill the SMs a tail would leave idle.
    int n_cta_all = 0;
    for (int c = 0; c < N_CTA_CLASSES; ++c) n_cta_all += ctx->n_cta[c];
    const int groups = (ctx->n_warp + ctx->n_none > 0) + (n_cta_all > 0) + (ctx->n_big > 0);
    const bool fork = !ctx->profile && groups > 1;
    cudaStream_t st_warp = st, st_cta = st;
    if (fork) {
        if (!ctx->stream_in) {
            CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream_in, cudaStreamNonBlocking));
            CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream_out, cudaStreamNonBlocking));
        }
        if (!ctx->ev_fork) {
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join[0], cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join[1], cudaEventDisableTiming));
        }
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, st));
        if (ctx->n_big > 0) {          // the radix chain keeps the context's stream
            st_warp = ctx->stream_in;
            st_cta = ctx->stream_out;
        } else {
            st_cta = ctx->stream_out;  // warp bin on the context's stream, CTA bin beside it
        }
        if (st_warp != st) CUDA_TRY(ctx, cudaStreamWaitEvent(st_warp, ctx->ev_fork, 0));
        CUDA_TRY(ctx, cudaStreamWaitEvent(st_cta, ctx->ev_fork, 0));
    }
    if (ctx->n_none > 0) {
        ProfScope ps(ctx, K_UNASSIGNED);
        launch_unassigned_topics(b, ctx->d_list_none.as<int32_t>(), ctx->n_none, st_warp);
        ++launches;
    }
    if (ctx->n_warp > 0) {
        ProfScope ps(ctx, K_WARP);
        launch_warp_topics(b, ctx->d_list_warp.as<int32_t>(), ctx->n_warp, st_warp);
        ++launches;
    }
    for (int c = 0; c < N_CTA_CLASSES; ++c) {
        if (ctx->n_cta[c] > 0) {
            ProfScope ps(ctx, K_CTA);
            launch_cta_topics(b, ctx->d_list_cta[c].as<int32_t>(), ctx->n_cta[c], c, sms, st_cta);
            ++launches;
        }
    }
    if (fork) {
        if (st_warp != st) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join[0], st_warp));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join[1], st_cta));
    }
    if (ctx->n_big > 0) {
        BigTopics bt = make_big(ctx);
        CUDA_TRY(ctx, cudaMemsetAsync(bt.range, 0, (size_t)ctx->n_big * 4 * sizeof(uint64_t), st));
        CUDA_TRY(ctx, cudaMemsetAsync(bt.need_tie, 0, (size_t)ctx->n_big * sizeof(int32_t), st));
        CUDA_TRY(ctx, cudaMemsetAsync(bt.ticket, 0, RT_TICKET_WORDS * sizeof(uint32_t), st));
        const int first_pass = ctx->has_part ? 0 : N_TIE_PASS;
        if (ctx->has_part) {
            ProfScope ps(ctx, K_ORDER);
            launch_big_order_check(bt, b.part, sms, st);
            ++launches;
        }
        {
            ProfScope ps(ctx, K_RANGE);
            launch_big_range(bt, b.lag, b.offs, b.part, sms, st);
            ++launches;
        }
        {
            ProfScope ps(ctx, K_SCAN);
            launch_big_scan(bt, first_pass, st);
            ++launches;
        }
        // Which digit passes does any topic need? Twelve words back to the host (one short
        // stream synchronisation, after the histogram + scan kernels) save the four launches
        // of every pass that would return at once: with low-entropy lags most passes do.
        uint32_t pass_any[N_PASS];
        bool any_pass = false;
        CUDA_TRY(ctx, cudaMemcpyAsync(pass_any, bt.pass_any, sizeof pass_any, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        for (int p = first_pass; p < N_PASS; ++p) {
            if (pass_any[p] == 0) continue;
            {
                ProfScope ps(ctx, K_TILE_HIST);
                launch_big_tile_hist(bt, p, b.lag, b.part, sms, st);
            }
            {
                ProfScope ps(ctx, K_PASS_SCANS);
                launch_big_pass_scans(bt, p, st);
            }
            {
                ProfScope ps(ctx, K_SCATTER);
                launch_big_scatter(bt, p, b.lag, b.part, ctx->exact_rank, sms, st);
            }
            launches += 4;
            any_pass = true;
        }
        // certificate of the fast ranking (radix.cuh 4d/4e); its verdict is read at the end of this call. It rides on
        // the finalize kernel's read of the sorted sequence; the sort-only mode (grouping) has no finalize.
        const bool certify = any_pass && !ctx->exact_rank;
        if (certify && ctx->sort_only) {
            ProfScope ps(ctx, K_VERIFY);
            launch_big_verify(bt, b.part, sms, st);
            ++launches;
        }
        verify_pending = certify;
        if (!ctx->sort_only) {
            ProfScope ps(ctx, K_FINALIZE);
            launch_big_finalize(bt, b, certify, sms, st);
            ++launches;
        }
        if (ctx->n_fill > 0 && !ctx->sort_only) {
            ProfScope ps(ctx, K_FILL);
            launch_big_fill(bt, ctx->d_list_fill.as<int32_t>(), ctx->n_fill, b, sms, st);
            ++launches;
        }
    }
    if (fork) {
        if (st_warp != st) CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_join[0], 0));
        CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_join[1], 0));
    }
    if (ctx->want_group && ctx->N > 0 && ctx->grp) {
        // stable partition of the flat result by member rank, see radix.cuh section 7
        lagassign_ctx* child = ctx->grp;
        {
            ProfScope ps(ctx, K_GROUP);
            launch_group_keys(ctx->N, b.out_cons, ctx->n_members, child->d_lag.as<int64_t>(), sms, st);
        }
        int rc = lagassign_execute(child);
        if (rc != LAGASSIGN_OK) return fail(ctx, rc, "grouping sort: %s", child->err.c_str());
        {
            ProfScope ps(ctx, K_GROUP);
            launch_group_finalize(make_big(child), child->d_lag.as<int64_t>(), ctx->n_members, ctx->N - ctx->n_assigned,
                                  b.out_part, b.topic_off, ctx->T, ctx->d_grp_topic.as<int32_t>(),
                                  ctx->d_grp_part.as<int32_t>(), ctx->d_member_start.as<int64_t>(),
                                  ctx->d_member_off.as<int64_t>(), ctx->n_assigned, sms, st);
        }
        launches += 4 + child->stats.kernel_launches;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->stats.kernel_launches = launches;
    if (verify_pending) {
        // The finalize / fill kernels above were launched optimistically. If the certificate failed (shared-memory
        // atomics of one instruction not served in lane order: never observed), everything is redone with the exact
        // ballot ranking, from now on.
        uint32_t failed = 0;
        CUDA_TRY(ctx, cudaMemcpyAsync(&failed, make_big(ctx).verify_fail, sizeof failed, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        if (failed) {
            ctx->exact_rank = true;
            ctx->rank_fallbacks += 1;
            const int rc = lagassign_execute(ctx);
            ctx->stats.kernel_launches += launches;
            return rc;
        }
    }
    ctx->executed = true;
    return LAGASSIGN_OK;
}

int lagassign_execute_range(lagassign_ctx* ctx, int64_t t_lo, int64_t t_hi) {
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    if (!ctx->shards.empty()) return fail(ctx, LAGASSIGN_ERR_UNSUPPORTED, "not available on a multi-device context");
    if (!ctx->uploaded) return fail(ctx, LAGASSIGN_ERR_STATE, "execute_range() before a successful upload()");
    if (t_lo < 0 || t_hi < t_lo || t_hi > ctx->T) return fail(ctx, LAGASSIGN_ERR_ARG, "topic range outside the batch");
    if (ctx->n_big > 0 || ctx->want_group || ctx->want_totals)
        return fail(ctx, LAGASSIGN_ERR_UNSUPPORTED, "execute_range covers batches of topics with at most 4096 partitions, without totals / grouping");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (t_lo == 0) {                       // a new pass over the batch: start a fresh per-kernel profile
        ctx->prof.clear();
        ctx->ev_used = 0;
        ctx->prof_resolved = false;
        ctx->stats.kernel_launches = 0;
    }
    DevBatch b = make_dev_batch(ctx);
    ctx->stats.kernel_launches += launch_topic_range(ctx, b, t_lo, t_hi, ctx->profile);
    CUDA_TRY(ctx, cudaGetLastError());
    if (t_hi == ctx->T) ctx->executed = true;
    return LAGASSIGN_OK;
}

int lagassign_sync(lagassign_ctx* ctx) {
    if (ctx && !ctx->shards.empty()) return multi_sync(ctx);
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LAGASSIGN_OK;
}

static int download_spec(lagassign_ctx* ctx, const OutSpec& out) {
    if (!ctx->executed) return fail(ctx, LAGASSIGN_ERR_STATE, "download() before execute()");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t N = (size_t)ctx->N;
    ctx->stats.d2h_bytes = 0;
    const int32_t* src_part = ctx->bound_part ? ctx->bound_part : ctx->d_out_part.as<int32_t>();
    const int32_t* src_cons = ctx->bound_cons ? ctx->bound_cons : ctx->d_out_cons.as<int32_t>();
    const uint8_t* d2h_part = reinterpret_cast<const uint8_t*>(src_part);
    const uint8_t* d2h_cons = reinterpret_cast<const uint8_t*>(src_cons);
    if (N && out.narrow) {
        if (!launch_narrow_results(ctx->stream, (int64_t)N, src_part, src_cons, out.pw, out.cw,
                                   ctx->d_narrow_part.as<uint8_t>(), ctx->d_narrow_cons.as<uint8_t>()))
            return fail(ctx, LAGASSIGN_ERR_ARG, "narrow result widths %d / %d", out.pw, out.cw);
        d2h_part = ctx->d_narrow_part.as<uint8_t>();
        d2h_cons = ctx->d_narrow_cons.as<uint8_t>();
    }
    if (N && out.h_part) {
        CUDA_TRY(ctx, cudaMemcpyAsync(out.h_part, d2h_part, N * out.pw, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += (int64_t)N * out.pw;
    }
    if (N && out.h_cons && out.cw) {
        CUDA_TRY(ctx, cudaMemcpyAsync(out.h_cons, d2h_cons, N * out.cw, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += (int64_t)N * out.cw;
    }
    if ((out.h_total || out.h_count) && !ctx->want_totals)
        return fail(ctx, LAGASSIGN_ERR_STATE, "out_total/out_count requested but LAGASSIGN_OPT_TOTALS was not enabled before upload()");
    const size_t TC = (size_t)ctx->total_tc;
    if (TC && out.h_total) {
        CUDA_TRY(ctx, cudaMemcpyAsync(out.h_total, ctx->d_out_total.p, TC * 8, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += (int64_t)TC * 8;
    }
    if (TC && out.h_count) {
        CUDA_TRY(ctx, cudaMemcpyAsync(out.h_count, ctx->d_out_count.p, TC * 4, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += (int64_t)TC * 4;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->n_big > 0) {
        int32_t pr = 0;
        CUDA_TRY(ctx, cudaMemcpy(&pr, ctx->d_ticket.as<uint32_t>() + N_PASS, sizeof pr, cudaMemcpyDeviceToHost));
        ctx->stats.radix_passes_run = pr;
    }
    return LAGASSIGN_OK;
}

int lagassign_download(lagassign_ctx* ctx, const lagassign_result* out) {
    if (ctx && !ctx->shards.empty()) return multi_download(ctx, out);
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    if (!out) return fail(ctx, LAGASSIGN_ERR_ARG, "result is NULL");
    return download_spec(ctx, spec_of(out));
}

// upload (plan cache) + kernels + download for one device; `narrow`: results leave in the narrow form, whose widths
// are decided here from the plan (or imposed by the multi-device parent) and reported back
static int assign_batch_one(lagassign_ctx* ctx, const lagassign_batch* in, OutSpec out, bool narrow, int* pw_out, int* cw_out) {
    const auto t0 = std::chrono::steady_clock::now();
    TRY(do_upload(ctx, in, nullptr, /*defer_data=*/true));     // plan + metadata; the large arrays follow
    if (narrow) {
        narrow_widths_of(ctx->has_part, ctx->plan_max_p, ctx->plan_multi_round, ctx->n_members, &out.pw, &out.cw);
        if (ctx->forced_pw) { out.pw = ctx->forced_pw; out.cw = ctx->forced_cw; }
        out.narrow = !(out.pw == 4 && out.cw == 4);
        if (pw_out) *pw_out = out.pw;
        if (cw_out) *cw_out = out.cw;
        if (out.narrow) {
            CUDA_TRY(ctx, ctx->d_narrow_part.ensure(std::max<size_t>((size_t)ctx->N * out.pw, 16)));
            CUDA_TRY(ctx, ctx->d_narrow_cons.ensure(std::max<size_t>((size_t)ctx->N * out.cw, 16)));
        }
    }
    if (ctx->n_big == 0 && !ctx->want_totals && !ctx->want_group && ctx->N >= PIPE_MIN_ELEMS && !out.h_total && !out.h_count) {
        const auto t1 = std::chrono::steady_clock::now();
        const int rc = assign_batch_pipelined(ctx, in, out);
        if (ctx->trace) {
            const auto t2 = std::chrono::steady_clock::now();
            fprintf(stderr, "[lagassign] assign_batch: plan+metadata %.3f ms, pipelined copies+kernels %.3f ms\n",
                    std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
        }
        return rc;
    }
    TRY(upload_data_now(ctx, in));
    TRY(lagassign_execute(ctx));
    return download_spec(ctx, out);
}

int lagassign_assign_batch(lagassign_ctx* ctx, const lagassign_batch* in, const lagassign_result* out) {
    if (ctx && !ctx->shards.empty()) return multi_assign_batch(ctx, in, out);
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    if (!out) return fail(ctx, LAGASSIGN_ERR_ARG, "result is NULL");
    return assign_batch_one(ctx, in, spec_of(out), /*narrow=*/false, nullptr, nullptr);
}

int lagassign_assign_batch_narrow(lagassign_ctx* ctx, const lagassign_batch* in, lagassign_result_narrow* out) {
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    if (!out) return fail(ctx, LAGASSIGN_ERR_ARG, "result is NULL");
    if (!ctx->shards.empty()) return multi_assign_batch_narrow(ctx, in, out);
    OutSpec o;
    o.h_part = static_cast<uint8_t*>(out->out_partition);
    o.h_cons = static_cast<uint8_t*>(out->out_consumer_rank);
    int pw = 4, cw = 4;
    const int rc = assign_batch_one(ctx, in, o, /*narrow=*/true, &pw, &cw);
    out->part_width = pw;
    out->cons_width = cw;
    return rc;
}

int lagassign_bind_device_outputs(lagassign_ctx* ctx, int32_t* d_out_partition, int32_t* d_out_consumer_rank) {
    if (ctx && !ctx->shards.empty()) return fail(ctx, LAGASSIGN_ERR_UNSUPPORTED, "not available on a multi-device context");
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    if ((d_out_partition == nullptr) != (d_out_consumer_rank == nullptr))
        return fail(ctx, LAGASSIGN_ERR_ARG, "bind both device outputs or neither");
    ctx->bound_part = d_out_partition;
    ctx->bound_cons = d_out_consumer_rank;
    return LAGASSIGN_OK;
}

int lagassign_download_lags(lagassign_ctx* ctx, int64_t* lag_out) {
    if (ctx && !ctx->shards.empty()) return fail(ctx, LAGASSIGN_ERR_UNSUPPORTED, "not available on a multi-device context");
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    if (!lag_out) return fail(ctx, LAGASSIGN_ERR_ARG, "lag_out is NULL");
    if (!ctx->executed) return fail(ctx, LAGASSIGN_ERR_STATE, "download_lags() before execute()");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->has_offsets && ctx->N) {
        // The kernels compute the lags on the fly and never store them (except for radix-path topics): materialise
        // them here, for this check only, with the same formula (lag_from_offsets).
        const int grid = (int)std::min<int64_t>((ctx->N + 255) / 256, (int64_t)ctx->sm_count * 16);
        k_compute_lag<<<grid, 256, 0, ctx->stream>>>(ctx->N, ctx->d_begin.as<int64_t>(), ctx->d_end.as<int64_t>(),
                                                     ctx->d_committed.as<int64_t>(), ctx->d_hascommitted.as<uint8_t>(),
                                                     ctx->reset_latest, ctx->d_lag.as<int64_t>());
    }
    if (ctx->N) CUDA_TRY(ctx, cudaMemcpyAsync(lag_out, ctx->d_lag.p, (size_t)ctx->N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LAGASSIGN_OK;
}

int lagassign_group_sizes(const lagassign_ctx* ctx, int64_t* n_members, int64_t* n_assigned) {
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    if (n_members) *n_members = ctx->n_members;
    if (n_assigned) *n_assigned = ctx->n_assigned;
    return LAGASSIGN_OK;
}

int lagassign_download_grouping(lagassign_ctx* ctx, const lagassign_grouping* out) {
    if (ctx && !ctx->shards.empty()) return fail(ctx, LAGASSIGN_ERR_UNSUPPORTED, "not available on a multi-device context");
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx is NULL");
    if (!out || !out->member_off) return fail(ctx, LAGASSIGN_ERR_ARG, "grouping output is NULL");
    if (!ctx->want_group) return fail(ctx, LAGASSIGN_ERR_STATE, "LAGASSIGN_OPT_GROUP was not enabled before upload()");
    if (!ctx->executed) return fail(ctx, LAGASSIGN_ERR_STATE, "download_grouping() before execute()");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t M = (size_t)ctx->n_members, NA = (size_t)ctx->n_assigned;
    if (ctx->N == 0) {
        for (size_t m = 0; m <= M; ++m) out->member_off[m] = 0;
        return LAGASSIGN_OK;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(out->member_off, ctx->d_member_off.p, (M + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (NA && out->topic) CUDA_TRY(ctx, cudaMemcpyAsync(out->topic, ctx->d_grp_topic.p, NA * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (NA && out->partition) CUDA_TRY(ctx, cudaMemcpyAsync(out->partition, ctx->d_grp_part.p, NA * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return LAGASSIGN_OK;
}

int lagassign_get_stats(const lagassign_ctx* ctx, lagassign_stats* stats) {
    if (!ctx || !stats) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx or stats is NULL");
    *stats = ctx->stats;
    return LAGASSIGN_OK;
}

static int resolve_profile(lagassign_ctx* ctx) {
    if (ctx->prof_resolved) return LAGASSIGN_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < K_KINDS; ++k) { ctx->prof_ms[k] = 0.f; ctx->prof_n[k] = 0; }
    for (const ProfRec& r : ctx->prof) {
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, r.a, r.b));
        ctx->prof_ms[r.kind] += ms;
        ctx->prof_n[r.kind] += 1;
    }
    ctx->prof_resolved = true;
    return LAGASSIGN_OK;
}

int lagassign_profile_count(lagassign_ctx* ctx, int* n_out) {
    if (!ctx || !n_out) return fail(ctx, LAGASSIGN_ERR_ARG, "ctx or n_out is NULL");
    if (!ctx->shards.empty()) { *n_out = 0; return LAGASSIGN_OK; }      // per-kernel times belong to one device
    TRY(resolve_profile(ctx));
    *n_out = K_KINDS;
    return LAGASSIGN_OK;
}

int lagassign_profile_entry(lagassign_ctx* ctx, int i, const char** name_out, float* ms_out, int* launches_out) {
    if (!ctx || i < 0 || i >= K_KINDS) return fail(ctx, LAGASSIGN_ERR_ARG, "bad profile index");
    TRY(resolve_profile(ctx));
    if (name_out) *name_out = kKernelNames[i];
    if (ms_out) *ms_out = ctx->prof_ms[i];
    if (launches_out) *launches_out = ctx->prof_n[i];
    return LAGASSIGN_OK;
}

int lagassign_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(nullptr, LAGASSIGN_ERR_ARG, "ptr is NULL");
    *ptr = nullptr;
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, LAGASSIGN_ERR_OOM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    return LAGASSIGN_OK;
}

int lagassign_host_free(void* ptr) {
    if (!ptr) return LAGASSIGN_OK;
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, LAGASSIGN_ERR_CUDA, "cudaFreeHost failed: %s", cudaGetErrorString(e)); }
    return LAGASSIGN_OK;
}

}  // extern "C"

// ---- multi-device contexts (lagassign_create_multi) -------------------------------------------------------
// One process, several GPUs, no collective: topics are independent (Main:215-225), so a batch is cut into
// contiguous topic ranges balanced by partition count, one per device; every device runs the whole path on its
// range — own stream, own PCIe link — and copies its results into the caller's host arrays at the range's
// offset. The host is the gather point, which is where the reference's caller (Main:149, :152-156) needs the
// map anyway. A topic is never split (its fill is one sequential chain).
namespace {

int multi_fail_from(lagassign_ctx* ctx, const std::vector<int>& rcs) {
    for (size_t i = 0; i < rcs.size(); ++i)
        if (rcs[i] != LAGASSIGN_OK)
            return fail(ctx, rcs[i], "device %d: %s", ctx->shards[i]->device, ctx->shards[i]->err.c_str());
    return LAGASSIGN_OK;
}

template <class F>
int multi_for_each(lagassign_ctx* ctx, F f) {
    const size_t n = ctx->shards.size();
    std::vector<int> rcs(n, LAGASSIGN_OK);
    std::vector<std::thread> th;
    for (size_t i = 1; i < n; ++i) th.emplace_back([&, i] { rcs[i] = f(i); });
    rcs[0] = f(0);
    for (auto& t : th) t.join();
    return multi_fail_from(ctx, rcs);
}

int multi_split(lagassign_ctx* ctx, const lagassign_batch* in) {
    if (!in || !in->topic_off || in->n_topics < 0) return fail(ctx, LAGASSIGN_ERR_ARG, "batch / topic_off is NULL");
    const int64_t T = in->n_topics;
    for (int64_t t = 0; t < T; ++t)
        if (in->topic_off[t + 1] < in->topic_off[t]) return fail(ctx, LAGASSIGN_ERR_ARG, "topic_off not monotone at %lld", (long long)t);
    if (T > 0 && (!in->topic_cset || !in->cset_off)) return fail(ctx, LAGASSIGN_ERR_ARG, "topic_cset / cset_off is NULL");
    const int64_t N = in->topic_off[T];
    const size_t n = ctx->shards.size();
    ctx->shard_ranges.assign(n, lagassign_ctx::ShardRange{});
    int64_t t_lo = 0, tc = 0;
    for (size_t i = 0; i < n; ++i) {
        auto& r = ctx->shard_ranges[i];
        int64_t t_hi = T;
        if (i + 1 < n) {
            const int64_t target = (int64_t)((__int128)N * (int64_t)(i + 1) / (int64_t)n);
            t_hi = std::lower_bound(in->topic_off, in->topic_off + T + 1, target) - in->topic_off;
            t_hi = std::min<int64_t>(std::max(t_hi, t_lo), T);
        }
        r.t0 = t_lo; r.t1 = t_hi; r.e0 = in->topic_off[t_lo]; r.e1 = in->topic_off[t_hi]; r.tc0 = tc;
        r.topic_off.resize((size_t)(t_hi - t_lo) + 1);
        for (int64_t t = t_lo; t <= t_hi; ++t) r.topic_off[(size_t)(t - t_lo)] = in->topic_off[t] - r.e0;
        for (int64_t t = t_lo; t < t_hi; ++t) {
            const int32_t s = in->topic_cset[t];
            if (s < 0 || s >= in->n_csets) return fail(ctx, LAGASSIGN_ERR_ARG, "topic_cset[%lld] out of range", (long long)t);
            tc += in->cset_off[s + 1] - in->cset_off[s];
        }
        t_lo = t_hi;
    }
    ctx->T = T; ctx->N = N; ctx->total_tc = tc;
    return LAGASSIGN_OK;
}

lagassign_batch shard_batch(const lagassign_ctx* ctx, const lagassign_batch* in, size_t i) {
    const auto& r = ctx->shard_ranges[i];
    lagassign_batch b = *in;
    b.n_topics = r.t1 - r.t0;
    b.topic_off = r.topic_off.data();
    b.lag = in->lag ? in->lag + r.e0 : nullptr;
    b.partition = in->partition ? in->partition + r.e0 : nullptr;
    b.topic_cset = in->topic_cset ? in->topic_cset + r.t0 : nullptr;
    return b;
}

lagassign_result shard_result(const lagassign_ctx* ctx, const lagassign_result* out, size_t i) {
    const auto& r = ctx->shard_ranges[i];
    lagassign_result o{};
    o.out_partition = out->out_partition ? out->out_partition + r.e0 : nullptr;
    o.out_consumer_rank = out->out_consumer_rank ? out->out_consumer_rank + r.e0 : nullptr;
    o.out_total = out->out_total ? out->out_total + r.tc0 : nullptr;
    o.out_count = out->out_count ? out->out_count + r.tc0 : nullptr;
    return o;
}

void multi_collect_stats(lagassign_ctx* ctx) {
    lagassign_stats a{};
    for (lagassign_ctx* c : ctx->shards) {
        const lagassign_stats& s = c->stats;
        a.n_partitions += s.n_partitions; a.n_assigned += s.n_assigned; a.topics_warp += s.topics_warp;
        a.topics_cta += s.topics_cta; a.topics_device += s.topics_device; a.topics_unassigned += s.topics_unassigned;
        a.device_sort_elems += s.device_sort_elems; a.radix_passes_run = std::max(a.radix_passes_run, s.radix_passes_run);
        a.kernel_launches += s.kernel_launches; a.h2d_bytes += s.h2d_bytes; a.d2h_bytes += s.d2h_bytes;
    }
    ctx->stats = a;
}

int multi_set_option(lagassign_ctx* ctx, int option, int64_t value) {
    if (option == LAGASSIGN_OPT_GROUP && value != 0)
        return fail(ctx, LAGASSIGN_ERR_UNSUPPORTED, "grouping by member is not available on a multi-device context");
    for (lagassign_ctx* c : ctx->shards) {
        const int rc = lagassign_set_option(c, option, value);
        if (rc != LAGASSIGN_OK) return fail(ctx, rc, "device %d: %s", c->device, c->err.c_str());
    }
    if (option == LAGASSIGN_OPT_TOTALS) ctx->want_totals = value != 0;
    return LAGASSIGN_OK;
}

int multi_upload(lagassign_ctx* ctx, const lagassign_batch* in, const lagassign_offsets* offs) {
    ctx->uploaded = ctx->executed = false;
    TRY(multi_split(ctx, in));
    const int rc = multi_for_each(ctx, [&](size_t i) {
        const lagassign_batch b = shard_batch(ctx, in, i);
        if (!offs) return lagassign_upload(ctx->shards[i], &b);
        const int64_t e0 = ctx->shard_ranges[i].e0;
        lagassign_offsets o = *offs;
        if (!o.begin_offset || !o.end_offset || !o.committed || !o.has_committed) return lagassign_upload_offsets(ctx->shards[i], &b, &o);
        o.begin_offset += e0; o.end_offset += e0; o.committed += e0; o.has_committed += e0;
        return lagassign_upload_offsets(ctx->shards[i], &b, &o);
    });
    multi_collect_stats(ctx);
    ctx->uploaded = rc == LAGASSIGN_OK;
    return rc;
}

int multi_execute(lagassign_ctx* ctx) {
    if (!ctx->uploaded) return fail(ctx, LAGASSIGN_ERR_STATE, "execute() before a successful upload()");
    const int rc = multi_for_each(ctx, [&](size_t i) { return lagassign_execute(ctx->shards[i]); });
    multi_collect_stats(ctx);
    ctx->executed = rc == LAGASSIGN_OK;
    return rc;
}

int multi_sync(lagassign_ctx* ctx) {
    return multi_for_each(ctx, [&](size_t i) { return lagassign_sync(ctx->shards[i]); });
}

int multi_download(lagassign_ctx* ctx, const lagassign_result* out) {
    if (!out) return fail(ctx, LAGASSIGN_ERR_ARG, "result is NULL");
    if (!ctx->executed) return fail(ctx, LAGASSIGN_ERR_STATE, "download() before execute()");
    const int rc = multi_for_each(ctx, [&](size_t i) {
        const lagassign_result o = shard_result(ctx, out, i);
        return lagassign_download(ctx->shards[i], &o);
    });
    multi_collect_stats(ctx);
    return rc;
}

int multi_assign_batch(lagassign_ctx* ctx, const lagassign_batch* in, const lagassign_result* out) {
    if (!out) return fail(ctx, LAGASSIGN_ERR_ARG, "result is NULL");
    ctx->uploaded = ctx->executed = false;
    TRY(multi_split(ctx, in));
    const int rc = multi_for_each(ctx, [&](size_t i) {
        const lagassign_batch b = shard_batch(ctx, in, i);
        const lagassign_result o = shard_result(ctx, out, i);
        return lagassign_assign_batch(ctx->shards[i], &b, &o);
    });
    multi_collect_stats(ctx);
    ctx->uploaded = ctx->executed = rc == LAGASSIGN_OK;
    return rc;
}

// Narrow form over several devices: the widths are those of the WHOLE batch (every shard writes entries of the same
// size into the caller's arrays at its range's offset).
int multi_assign_batch_narrow(lagassign_ctx* ctx, const lagassign_batch* in, lagassign_result_narrow* out) {
    ctx->uploaded = ctx->executed = false;
    TRY(multi_split(ctx, in));
    int64_t max_p = 0, n_members = 0;
    bool multi_round = false;
    for (int64_t t = 0; t < in->n_topics; ++t) {
        const int64_t P = in->topic_off[t + 1] - in->topic_off[t];
        const int32_t s = in->topic_cset[t];
        const int64_t C = in->cset_off[s + 1] - in->cset_off[s];
        max_p = std::max(max_p, P);
        multi_round = multi_round || (C > 0 && P > C);
    }
    if (in->cset_rank)
        for (int64_t j = 0; j < in->cset_off[in->n_csets]; ++j) n_members = std::max<int64_t>(n_members, (int64_t)in->cset_rank[j] + 1);
    int pw = 4, cw = 4;
    narrow_widths_of(in->partition != nullptr, max_p, multi_round, n_members, &pw, &cw);
    out->part_width = pw;
    out->cons_width = cw;
    const int rc = multi_for_each(ctx, [&](size_t i) {
        const lagassign_batch b = shard_batch(ctx, in, i);
        const auto& r = ctx->shard_ranges[i];
        lagassign_result_narrow o = *out;
        o.out_partition = out->out_partition ? static_cast<uint8_t*>(out->out_partition) + r.e0 * pw : nullptr;
        o.out_consumer_rank = out->out_consumer_rank ? static_cast<uint8_t*>(out->out_consumer_rank) + r.e0 * cw : nullptr;
        ctx->shards[i]->forced_pw = pw;
        ctx->shards[i]->forced_cw = cw;
        const int rc1 = lagassign_assign_batch_narrow(ctx->shards[i], &b, &o);
        ctx->shards[i]->forced_pw = 0;
        return rc1;
    });
    multi_collect_stats(ctx);
    ctx->uploaded = ctx->executed = rc == LAGASSIGN_OK;
    return rc;
}

}  // namespace

extern "C" int lagassign_create_multi(lagassign_ctx** ctx, const int* devices, int n_devices) {
    if (!ctx) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx out-pointer is NULL");
    *ctx = nullptr;
    if (!devices || n_devices < 1) return fail(nullptr, LAGASSIGN_ERR_ARG, "devices is NULL or n_devices < 1");
    lagassign_ctx* parent = new (std::nothrow) lagassign_ctx();
    if (!parent) return fail(nullptr, LAGASSIGN_ERR_OOM, "out of host memory");
    parent->device = devices[0];
    for (int i = 0; i < n_devices; ++i) {
        lagassign_ctx* c = nullptr;
        const int rc = lagassign_create(&c, devices[i]);
        if (rc != LAGASSIGN_OK) {
            const std::string m = g_thread_error;
            for (lagassign_ctx* d : parent->shards) lagassign_destroy(d);
            delete parent;
            return fail(nullptr, rc, "device %d: %s", devices[i], m.c_str());
        }
        parent->shards.push_back(c);
    }
    *ctx = parent;
    return LAGASSIGN_OK;
}

extern "C" int lagassign_shard_count(const lagassign_ctx* ctx, int* n_out) {
    if (!ctx || !n_out) return fail(nullptr, LAGASSIGN_ERR_ARG, "ctx or n_out is NULL");
    *n_out = ctx->shards.empty() ? 1 : (int)ctx->shards.size();
    return LAGASSIGN_OK;
}
