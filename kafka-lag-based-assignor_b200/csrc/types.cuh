// types.cuh — plain structs and constants shared by the host API and the kernels.
#pragma once
#include <cstddef>
#include <cstdint>

#include "common.cuh"

namespace lagassign {

// Raw offsets of the fused lag computation (lagassign_upload_offsets; reference computePartitionLag,
// LagBasedPartitionAssignor.java:376-404). end == nullptr: the batch carries lags.
struct LagOffsets {
    const int64_t* begin;       // [N]
    const int64_t* end;         // [N]
    const int64_t* committed;   // [N]
    const uint8_t* has;         // [N]
    int reset_latest;
};

struct DevBatch {
    const int64_t* topic_off;   // [T+1]
    const int64_t* lag;         // [N]  (with offsets: the context's own buffer; only the radix path's topics are
                                //       materialised there, by k_big_range)
    LagOffsets offs;            // see above
    const int32_t* part;        // [N] or null
    const int32_t* topic_cset;  // [T]
    const int64_t* cset_off;    // [S+1]
    const int32_t* cset_rank;   // [R]
    const int64_t* tc_off;      // [T+1] offsets of the totals/counts blocks (or null)
    int32_t* out_part;          // [N]
    int32_t* out_cons;          // [N]
    int64_t* out_total;         // null unless enabled
    int32_t* out_count;
};

// Main:376-404 for element e: max(end - (has ? committed : (latest ? end : begin)), 0), wrapping subtract
__device__ __forceinline__ int64_t lag_from_offsets(const LagOffsets& o, int64_t e) {
    const int64_t end = ld_stream(o.end + e);
    int64_t next;
    if (o.has[e]) next = ld_stream(o.committed + e);          // Main:384-386
    else next = o.reset_latest ? end : ld_stream(o.begin + e);   // Main:391-396
    const int64_t d = (int64_t)((uint64_t)end - (uint64_t)next);
    return d > 0 ? d : 0;                                        // Main:402
}
// Where a kernel's lags come from: LAG_PLAIN = the lag array, LAG_OFFS = computed from the raw offsets while the
// sort keys are built (no lag array exists), LAG_EITHER = decided at run time (uniform over the grid).
enum : int { LAG_PLAIN = 0, LAG_OFFS = 1, LAG_EITHER = 2 };
template <int MODE>
__device__ __forceinline__ int64_t lag_load(const DevBatch& b, int64_t e) {
    if constexpr (MODE == LAG_PLAIN) return ld_stream(b.lag + e);
    else if constexpr (MODE == LAG_OFFS) return lag_from_offsets(b.offs, e);
    else return b.offs.end != nullptr ? lag_from_offsets(b.offs, e) : ld_stream(b.lag + e);
}

#ifndef LAGASSIGN_RT_THREADS
#define LAGASSIGN_RT_THREADS 256
#endif
#ifndef LAGASSIGN_RT_IPT
#define LAGASSIGN_RT_IPT 16
#endif
#ifndef LAGASSIGN_RT_MINBLOCKS
#define LAGASSIGN_RT_MINBLOCKS 2
#endif
constexpr int RT_THREADS = LAGASSIGN_RT_THREADS;
constexpr int RT_IPT = LAGASSIGN_RT_IPT;
constexpr int RT_TILE = RT_THREADS * RT_IPT;   // 4096 elements per tile
constexpr int RT_WARPS = RT_THREADS / 32;
constexpr int N_TIE_PASS = 4;
constexpr int N_LAG_PASS = 8;
constexpr int N_PASS = N_TIE_PASS + N_LAG_PASS;

constexpr int RT_SEG = 64;              // tiles per scan segment
constexpr int RT_TICKET_WORDS = 2 * 12 + 2;
constexpr int RT_BUF_ALIGN = 16;        // topics start at multiples of 16 elements in the sort buffers (16-byte aligned bulk copies / digit-byte vectors)
constexpr int64_t RT_MAX_P = (1ll << 31) - 1;   // positions and digit counts are 32-bit

enum : int { SEL_RAW = 0, SEL_A = 1, SEL_B = 2, SEL_SKIP = -1 };

// One 32-byte record per tile, built on the host, so a tile's geometry is one load
// instead of a chain of dependent ones (tile -> topic -> five per-topic arrays).
struct TileDesc {
    int32_t b;               // large-topic index
    int32_t tile_in_topic;
    int32_t n_valid;         // elements in this tile (RT_TILE except for a topic's last tile)
    int32_t seg;             // global scan-segment id
    int64_t buf_off;         // boff + tile start: element offset in the compact sort buffers
    int64_t raw_off;         // topic_off + tile start: element offset in the caller's arrays
};
static_assert(sizeof(TileDesc) == 32, "TileDesc is loaded as two 16-byte words");

struct BigTopics {
    int n_big;                 // B
    int n_tiles;               // G
    const int32_t* topic;      // [B] topic id
    const int64_t* off;        // [B] offset in the caller's arrays (= topic_off[t])
    const int64_t* P;          // [B]
    const int64_t* boff;       // [B] offset in the compact sort buffers (a multiple of RT_BUF_ALIGN elements)
    const int32_t* tile0;      // [B] first tile of the topic
    const int32_t* tile_bt;    // [G] tile -> b
    const int4* tile_desc;     // [G][2] per-tile descriptor, see TileDesc
    uint64_t* keyA; uint64_t* keyB;   // [sum P]
    uint32_t* payA; uint32_t* payB;
    uint8_t* dig;              // [sum P] digit of every element in the topic's NEXT executed pass, in the order of the
                               //         buffer the last scatter wrote (input of k_big_tile_hist)
    uint64_t* range;           // [B][4] OR of keys, OR of ~keys, OR of id keys, OR of ~id keys (k_big_range)
    uint32_t* pass_rec;        // [N_PASS][G] per (pass, tile): source | destination << 8 | next executed pass << 16 (k_big_scan)
    int8_t* sel_src;           // [B][N_PASS]  SEL_* (SEL_SKIP = pass skipped for this topic)
    int8_t* sel_dst;           // [B][N_PASS]
    int8_t* nxt_pass;          // [B][N_PASS]  the topic's next executed pass after this one (-1: none)
    int8_t* sel_final;         // [B] where the sorted data of the topic ended up
    int8_t* all_equal;         // [B] every lag of the topic has the same value (all eight digit passes constant)
    int32_t* need_tie;         // [B] partition ids not ascending
    uint32_t* tile_hist;       // [G][256] digit counts of each tile, then exclusive prefix inside its segment
    const int32_t* seg0;       // [B] first scan segment of the topic (segments of RT_SEG tiles)
    const int32_t* seg_bt;     // [n_segs] segment -> b
    int n_segs;
    uint32_t* seg_total;       // [n_segs][256] segment totals, then start of (segment, digit) inside the topic
    uint32_t* ticket;          // [RT_TICKET_WORDS] small per-execute words: [N_PASS] = passes_run, then pass_any[N_PASS], then verify_fail
    int32_t* passes_run;       // [1] max over topics of executed passes (stat)
    uint32_t* pass_any;        // [N_PASS] does any topic run this pass? (else the launch returns at once)
    uint32_t* verify_fail;     // [1] set by k_big_verify when a sorted sequence is not strictly increasing
    const int64_t* fill_ws_off; // [B] slot offset of the topic's round-fill workspace (C > FILL_CTA_MAX_C), else -1
    uint64_t* fill_ws_key;     // consumers of such topics: (total key, index) in global memory,
    uint32_t* fill_ws_idx;     //   pow2 >= C slots per topic
};


constexpr int BIGFILL_THREADS = 256;   // (1024 threads were measured on B200: 4.3 -> 6.1 ms for config 2's round chain)
constexpr int BIGFILL_MAXJ = FILL_CTA_MAX_C / BIGFILL_THREADS;
constexpr int BIGFILL_REG_MAX_C = 4096;                      // consumers kept in registers up to here
constexpr size_t BIGFILL_SMEM = (size_t)FILL_CTA_MAX_C * 12; // >= 2*3*4096 words of exchange buffers

// ---- launchers (one translation unit per kernel family, see launch_*.cu) ----------------
cudaError_t configure_cta_kernels();
cudaError_t configure_big_kernels();
void launch_unassigned_topics(const DevBatch& b, const int32_t* list, int n, cudaStream_t st);
void launch_warp_topics(const DevBatch& b, const int32_t* list, int n, cudaStream_t st);
constexpr int N_CTA_CLASSES = 4;   // topics of 257..512 / ..1024 / ..2048 / ..4096 partitions
inline int cta_class_of(int64_t P) { return P <= 512 ? 0 : P <= 1024 ? 1 : P <= 2048 ? 2 : 3; }
void launch_cta_topics(const DevBatch& b, const int32_t* list, int n, int cls, int sms, cudaStream_t st);
void launch_big_order_check(const BigTopics& bt, const int32_t* part, int sms, cudaStream_t st);
void launch_big_range(const BigTopics& bt, const int64_t* lag, const LagOffsets& offs, const int32_t* part, int sms, cudaStream_t st);
void launch_big_scan(const BigTopics& bt, int first_pass, cudaStream_t st);
void launch_big_tile_hist(const BigTopics& bt, int pass, const int64_t* lag, const int32_t* part, int sms, cudaStream_t st);
void launch_big_pass_scans(const BigTopics& bt, int pass, cudaStream_t st);   // two kernels
void launch_big_scatter(const BigTopics& bt, int pass, const int64_t* lag, const int32_t* part, bool exact, int sms, cudaStream_t st);
void launch_big_verify(const BigTopics& bt, const int32_t* part, int sms, cudaStream_t st);
cudaError_t probe_atomic_lane_order(int* in_order, cudaStream_t st);
void launch_big_finalize(const BigTopics& bt, const DevBatch& b, bool verify, int sms, cudaStream_t st);
void launch_big_fill(const BigTopics& bt, const int32_t* fill_list, int n_fill, const DevBatch& b, int sms, cudaStream_t st);
void launch_group_keys(int64_t n, const int32_t* out_cons, int64_t n_members, int64_t* glag, int sms, cudaStream_t st);
void launch_group_finalize(const BigTopics& bt, const int64_t* glag, int64_t n_members, int64_t n_unassigned,
                           const int32_t* out_part, const int64_t* topic_off, int64_t n_topics, int32_t* grp_topic,
                           int32_t* grp_part, int64_t* member_start, int64_t* member_off, int64_t n_assigned, int sms,
                           cudaStream_t st);

// narrow copy of a result range (exchange.cu): entries of pw / cw bytes; cw == 0: consumers are not copied
bool launch_narrow_results(cudaStream_t st, int64_t n, const int32_t* part, const int32_t* cons, int pw, int cw,
                           uint8_t* dpart, uint8_t* dcons);

}  // namespace lagassign
