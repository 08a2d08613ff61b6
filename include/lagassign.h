/*
 * lagassign.h — C ABI of the B200-native lag-balancing partition assignor.
 *
 * This is the drop-in boundary for ONE hot path of
 * grantneale/kafka-lag-based-assignor: the pure function
 *
 *     static Map<String, List<TopicPartition>> assign(
 *         Map<String, List<TopicPartitionLag>> partitionLagPerTopic,
 *         Map<String, List<String>> subscriptions)
 *
 * (reference: src/main/java/com/github/grantneale/kafka/
 *  LagBasedPartitionAssignor.java:166-188, which drives assignTopic :204-277
 *  and consumersPerTopic :410-426; the plugin entry that calls it is :149).
 *
 * The reference is a JVM class with no native code, so this header is what a
 * JNI shim would bind. Plain C: pointers and sizes only,
 * no C++ types, no exceptions, no torch types. All entry points return 0 on
 * success or a negative LAGASSIGN_ERR_* code; the message is available from
 * lagassign_last_error(). There is NO CPU fallback: without a CUDA device
 * lagassign_create() fails with LAGASSIGN_ERR_NO_DEVICE.
 *
 * Flat layout ("CSR") replacing the reference's maps of boxed objects:
 *
 *   topics        t = 0..T-1, in the order the caller wants results grouped
 *   topic_off     int64[T+1]  partitions of topic t are [topic_off[t], topic_off[t+1])
 *   lag           int64[N]    TopicPartitionLag.getLag()        (Main:451-453)
 *   partition     int32[N]    TopicPartitionLag.getPartition()  (Main:447-449);
 *                             NULL means ids 0..P-1 in order
 *   consumer sets s = 0..S-1: the distinct values of consumersPerTopic (Main:410-426)
 *   cset_off      int64[S+1]
 *   cset_rank     int32[R]    members of set s, as GLOBAL member ranks, strictly
 *                             ascending inside each set. rank r = position of the
 *                             member id in String.compareTo order (Main:258-259)
 *   topic_cset    int32[T]    which consumer set subscribes topic t
 *
 * Result, per topic, in ASSIGNMENT ORDER (the order of the loop Main:237, i.e.
 * lag descending / partition ascending, Main:228-235):
 *
 *   out_partition[topic_off[t] + k]      partition id of the k-th assigned partition
 *   out_consumer_rank[topic_off[t] + k]  global rank of the member chosen for it
 *                                        (Collections.min of Main:240-263)
 *
 * A topic whose consumer set is empty is never visited by the reference
 * (Main:177 iterates consumersPerTopic only): its partitions come back in input
 * order with out_consumer_rank = -1.
 *
 * Optional check outputs (reference: consumerTotalLags / consumerTotalPartitions
 * at the end of assignTopic, Main:216-225, :265-266), laid out per topic then
 * per member of the topic's consumer set (set order), at offset
 * tc_off[t] = sum_{u<t} |cset(topic_cset[u])|:
 *
 *   out_total  int64[sum_t C_t]   wrapping signed 64-bit sums
 *   out_count  int32[sum_t C_t]
 */
#ifndef LAGASSIGN_H
#define LAGASSIGN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAGASSIGN_ABI_VERSION 3

#define LAGASSIGN_OK               0
#define LAGASSIGN_ERR_ARG         -1  /* NULL pointer, non-monotone offsets, unsorted set ... */
#define LAGASSIGN_ERR_CUDA        -2  /* a CUDA runtime call or kernel failed */
#define LAGASSIGN_ERR_OOM         -3  /* host or device allocation failed */
#define LAGASSIGN_ERR_STATE       -4  /* call order violated (execute before upload ...) */
#define LAGASSIGN_ERR_NO_DEVICE   -5  /* no usable CUDA device: there is no CPU fallback */
#define LAGASSIGN_ERR_UNSUPPORTED -6  /* shape outside what the kernels cover */

typedef struct lagassign_ctx lagassign_ctx;

/* Inputs of one assign() call. All pointers are HOST pointers owned by the
 * caller; the library copies what it needs before returning. */
typedef struct lagassign_batch {
    int64_t        n_topics;    /* T */
    const int64_t *topic_off;   /* [T+1], topic_off[0] == 0, non-decreasing */
    const int64_t *lag;         /* [N], N = topic_off[T] */
    const int32_t *partition;   /* [N] or NULL */
    const int32_t *topic_cset;  /* [T], values in [0, n_csets) */
    int64_t        n_csets;     /* S */
    const int64_t *cset_off;    /* [S+1], cset_off[0] == 0, non-decreasing */
    const int32_t *cset_rank;   /* [cset_off[S]], strictly ascending inside a set, >= 0 */
} lagassign_batch;

/* Outputs of one assign() call. HOST pointers owned by the caller. out_total
 * and out_count may be NULL (they are only produced when enabled with
 * lagassign_set_option(LAGASSIGN_OPT_TOTALS, 1) before upload). */
typedef struct lagassign_result {
    int32_t *out_partition;      /* [N] */
    int32_t *out_consumer_rank;  /* [N] */
    int64_t *out_total;          /* [sum_t C_t] or NULL */
    int32_t *out_count;          /* [sum_t C_t] or NULL */
} lagassign_result;

/* Raw offsets for the fused lag computation (reference: computePartitionLag,
 * Main:376-404, called per partition at Main:344-356). When given to
 * lagassign_upload_offsets() the device computes
 *   lag = max(end - (has_committed ? committed : (reset_latest ? end : begin)), 0)
 * with wrapping int64 subtraction (one elementwise kernel ahead of the sort; the lags
 * never exist on the host). */
typedef struct lagassign_offsets {
    const int64_t *begin_offset;  /* [N]  topicBeginOffsets.getOrDefault(p, 0)  Main:350 */
    const int64_t *end_offset;    /* [N]  topicEndOffsets.getOrDefault(p, 0)    Main:351 */
    const int64_t *committed;     /* [N]  OffsetAndMetadata.offset()            Main:386 */
    const uint8_t *has_committed; /* [N]  partitionMetadata != null             Main:384 */
    int32_t        reset_latest;  /* auto.offset.reset equalsIgnoreCase "latest" Main:391 */
} lagassign_offsets;

/* Counters of the last execute(), for benchmarks and tests. */
typedef struct lagassign_stats {
    int64_t n_partitions;        /* N */
    int64_t n_assigned;          /* partitions that received a consumer */
    int64_t topics_warp;         /* topics sorted+filled by one warp  (P <= 256) */
    int64_t topics_cta;          /* topics sorted+filled by one CTA   (P <= 4096) */
    int64_t topics_device;       /* topics sorted by the device-wide radix path */
    int64_t topics_unassigned;   /* topics with an empty consumer set */
    int64_t device_sort_elems;   /* elements that went through the radix path */
    int32_t radix_passes_run;    /* digit passes that moved data (max over topics) */
    int32_t kernel_launches;     /* kernels launched by the last execute() */
    int64_t h2d_bytes;           /* bytes copied host->device by the last upload() */
    int64_t d2h_bytes;           /* bytes copied device->host by the last download() */
} lagassign_stats;

#define LAGASSIGN_OPT_TOTALS  1  /* produce out_total / out_count (default 0) */
#define LAGASSIGN_OPT_PROFILE 2  /* bracket every kernel launch with CUDA events (default 0) */
#define LAGASSIGN_OPT_GROUP   3  /* also group the result by member on the device (default 0) */
#define LAGASSIGN_OPT_PIPE_CHUNK 4 /* partitions per chunk of the copy/compute pipeline of
                                    * lagassign_assign_batch (0 = library default) */
#define LAGASSIGN_OPT_PLAN_CACHE 5 /* reuse the plan and the device-resident metadata when a batch has the
                                    * shape (topic_off, topic_cset, consumer sets, flags) of the previous
                                    * one and only lags / ids / offsets differ (default 1) */

#define LAGASSIGN_OPT_EXACT_RANK 6 /* device-wide radix path: 1 = rank with the exact ballot form from the start instead
                                    * of the fast atomic form + certificate (default 0, or 1 when the probe at
                                    * lagassign_create finds that the device does not serve same-address shared
                                    * atomics of one instruction in lane order); see csrc/radix.cuh section 4d */

int  lagassign_abi_version(void);

/* Number of CUDA devices visible; 0 devices is not an error here. */
int  lagassign_device_count(int *n_devices);

/* One context = one device + one stream + reusable device workspaces.
 * Mirrors one LagBasedPartitionAssignor instance (Main:83-89). Not re-entrant:
 * like the reference's instance state (Main:87-89) a context must be used from
 * one thread at a time; distinct contexts are independent. */
int  lagassign_create(lagassign_ctx **ctx, int device);

/* Same, but all work is enqueued on the caller's CUDA stream (a cudaStream_t
 * passed as void*), so a caller that times with its own events on that stream
 * sees the kernels. */
int  lagassign_create_on_stream(lagassign_ctx **ctx, int device, void *cuda_stream);

/* Move a context created with lagassign_create_on_stream to another stream of the caller (for example the stream a
 * CUDA graph is being captured on: lagassign_execute_range enqueues kernels only, so a caller may capture its steps). */
int  lagassign_set_stream(lagassign_ctx *ctx, void *cuda_stream);

/* One context over SEVERAL devices of this process ("topics shard naturally across the 8 GPUs"):
 * the batch of lagassign_assign_batch / upload is cut into contiguous topic ranges balanced by partition count,
 * one per device; every device runs the whole path on its range concurrently (its own stream and PCIe link) and
 * copies its results straight into the caller's host arrays at the range's offset. No collective: the host is
 * the gather point, which is where the JVM leader that calls Main:149 needs the map. A topic is never split.
 * Not available on such a context: lagassign_bind_device_outputs, lagassign_download_lags, LAGASSIGN_OPT_GROUP,
 * per-kernel profiles (lagassign_profile_count reports 0 kinds). A device may be listed more than once (several
 * streams on one GPU; the tests use it to exercise the split on a one-GPU box). */
int  lagassign_create_multi(lagassign_ctx **ctx, const int *devices, int n_devices);
/* Number of devices behind a context (1 for lagassign_create / lagassign_create_on_stream). */
int  lagassign_shard_count(const lagassign_ctx *ctx, int *n_out);

void lagassign_destroy(lagassign_ctx *ctx);

/* Last error text of this context (or of the calling thread when ctx is NULL). */
const char *lagassign_last_error(const lagassign_ctx *ctx);

int  lagassign_set_option(lagassign_ctx *ctx, int option, int64_t value);

/* The whole hot path with host buffers in and out: what replaces the call at
 * Main:149. Equivalent to upload + execute + download. Blocking. */
int  lagassign_assign_batch(lagassign_ctx *ctx, const lagassign_batch *in,
                            const lagassign_result *out);

/* The same call with the results in their NARROW form (optional; for callers that read bytes as cheaply as ints,
 * e.g. a JVM over direct buffers): at BASELINE config 4 the device-to-host copy shrinks from 8 bytes per partition
 * to 1, which leaves the host-to-device copy of the lags as the only PCIe leg of the call.
 *   out_partition      [N] entries of part_width bytes (unsigned): 1 when the batch has no explicit ids and every
 *                      topic has at most 256 partitions, 2 up to 65536, else 4 (always 4 with explicit ids)
 *   out_consumer_rank  [N] entries of cons_width bytes: 0 — nothing is written — when every topic of the batch has
 *                      P <= C or no subscriber: the consumer of sorted partition k is then cset_rank[cset_off[s] + k]
 *                      (Main:240-263 on all-zero totals; -1 for a topic without subscribers), which
 *                      the caller reads from its own table; else 2 (all ones = no subscriber) when there are fewer
 *                      than 65535 members, else 4 (int32)
 * Both buffers must hold 4 * N bytes unless the caller knows the widths from an earlier call on the same shape;
 * part_width / cons_width are set by the call. Works on single- and multi-device contexts. */
typedef struct lagassign_result_narrow {
    void   *out_partition;
    void   *out_consumer_rank;
    int32_t part_width;          /* out */
    int32_t cons_width;          /* out */
} lagassign_result_narrow;
int  lagassign_assign_batch_narrow(lagassign_ctx *ctx, const lagassign_batch *in,
                                   lagassign_result_narrow *out);

/* Split form. upload: validate, plan (bin topics by size), copy inputs to HBM.
 * execute: enqueue the kernels (asynchronous, except that batches with topics of
 * more than 4096 partitions wait once for the digit histograms to decide which radix
 * passes to launch and once, at the end, for the certificate of the sort).
 * download: wait and copy results to host. execute may be repeated
 * on the same uploaded batch. */
int  lagassign_upload(lagassign_ctx *ctx, const lagassign_batch *in);
int  lagassign_upload_offsets(lagassign_ctx *ctx, const lagassign_batch *in,
                              const lagassign_offsets *offs);
int  lagassign_execute(lagassign_ctx *ctx);
/* execute() for the topics [t_lo, t_hi) of the uploaded batch only (asynchronous). For callers that overlap the
 * result exchange of one topic range with the kernels of the next (distributed.py). Batches whose topics all have at
 * most 4096 partitions, without totals / grouping; LAGASSIGN_ERR_UNSUPPORTED otherwise (use execute()). The per-kernel
 * profile accumulates over the ranges of one pass (it restarts at t_lo == 0). */
int  lagassign_execute_range(lagassign_ctx *ctx, int64_t t_lo, int64_t t_hi);
int  lagassign_sync(lagassign_ctx *ctx);
int  lagassign_download(lagassign_ctx *ctx, const lagassign_result *out);

/* Multi-GPU: make execute() write out_partition / out_consumer_rank straight
 * into caller-owned DEVICE memory (e.g. this rank's slice of an all-gather
 * buffer). Pass NULLs to go back to the context's own buffers. */
int  lagassign_bind_device_outputs(lagassign_ctx *ctx, int32_t *d_out_partition,
                                   int32_t *d_out_consumer_rank);

/* Result exchange for the one-process-per-GPU arrangement (topics sharded over ranks, every rank finally holding
 * the whole map as two int32 arrays: what the reference returns at Main:187). Only the narrow form of a rank's
 * results crosses NVLink: partition ids in part_width = 1 / 2 / 4 bytes (every id of the batch below 256 / 65536 /
 * any), consumers in cons_width = 2 / 4 bytes, or cons_width = 0 when every topic of the batch has P <= C or no
 * subscriber — then the consumer of sorted partition k is the k-th member of the topic's set (SURVEY 0 fact 5;
 * Main:240-263 on all-zero totals) and every rank writes it from the replicated table.
 *   pack:   this rank's int32 results (n entries, device pointers) -> its slot (slot_bytes(n, ...) bytes, 16-aligned)
 *   gather: the caller's collective moves the slots (slots of slot_bytes(max n over ranks) each, rank order)
 *   unpack: all slots -> out_partition / out_consumer_rank of the WHOLE batch, one warp per topic; needs the whole
 *           batch's topic_off / topic_cset / consumer sets and the first topic of every rank (world + 1 entries)
 *           in device memory.
 * All pointers are DEVICE pointers; work is enqueued on cuda_stream. */
int64_t lagassign_exchange_slot_bytes(int64_t n, int part_width, int cons_width);
int  lagassign_exchange_pack(void *cuda_stream, int64_t n, const int32_t *d_part, const int32_t *d_cons,
                             int part_width, int cons_width, void *d_slot);
/* unpack for ARBITRARY topic ranges per rank (a chunk of every rank's shard): rank r's slot holds the results of the
 * topics [d_range_lo[r], d_range_hi[r]); skip_rank (or -1) names a rank whose range is already in place in the output
 * arrays (the caller's own, written there by its kernels). */
int  lagassign_exchange_unpack_ranges(void *cuda_stream, const int64_t *d_topic_off, const int32_t *d_topic_cset,
                                      const int64_t *d_cset_off, const int32_t *d_cset_rank,
                                      const int64_t *d_range_lo, const int64_t *d_range_hi, int world, int skip_rank,
                                      int64_t max_topics_per_rank, const void *d_slots, int64_t slot_bytes,
                                      int part_width, int cons_width,
                                      int32_t *d_out_partition, int32_t *d_out_consumer_rank);
int  lagassign_exchange_unpack(void *cuda_stream, int64_t n_topics, const int64_t *d_topic_off,
                               const int32_t *d_topic_cset, const int64_t *d_cset_off, const int32_t *d_cset_rank,
                               const int64_t *d_shard_first_topic, int world, const void *d_slots,
                               int64_t slot_bytes, int part_width, int cons_width,
                               int32_t *d_out_partition, int32_t *d_out_consumer_rank);

/* consumersPerTopic (Main:410-426, called at Main:176) and the de-duplication of consumer sets, on the device
 * (SURVEY 8f row 3): from the subscriptions as integers — members in rank order (String.compareTo), every entry the
 * index of a subscribed topic in the caller's topic order — to the consumer-set part of lagassign_batch:
 *   topic_cset [T]          set of every topic
 *   cset_off   [n_csets+1]  (capacity T + 1)
 *   cset_rank  [...]        members of each set, ascending rank (capacity: number of subscription entries)
 * Sets are numbered in order of first appearance over the topics; a member that lists a topic twice counts once (as
 * in the reference, whose per-topic maps are keyed by member id, Main:216-225); a topic nobody lists gets the empty set.
 * Stateless: allocates and frees its own device memory on `device`. Host pointers in and out. */
typedef struct lagassign_subscriptions {
    int64_t        n_members;   /* M */
    int64_t        n_topics;    /* T */
    const int64_t *sub_off;     /* [M+1] entries of member (rank) m are [sub_off[m], sub_off[m+1]) */
    const int32_t *sub_topic;   /* [sub_off[M]] topic indices in [0, T) */
} lagassign_subscriptions;
typedef struct lagassign_consumer_sets {
    int64_t  n_csets;           /* out */
    int32_t *topic_cset;        /* [T] */
    int64_t *cset_off;          /* capacity T + 1 */
    int32_t *cset_rank;         /* capacity sub_off[M] */
} lagassign_consumer_sets;
int  lagassign_invert_subscriptions(int device, const lagassign_subscriptions *in, lagassign_consumer_sets *out);
const char *lagassign_subscriptions_last_error(void);

/* Device copy of the computed lags (valid after execute(); for tests of the
 * fused lag computation). */
int  lagassign_download_lags(lagassign_ctx *ctx, int64_t *lag_out);

int  lagassign_get_stats(const lagassign_ctx *ctx, lagassign_stats *stats);

/* Per-member view of the result (enabled by LAGASSIGN_OPT_GROUP before upload), i.e. the
 * reference's return value Map<member, List<TopicPartition>> (Main:166-188) in CSR form: the
 * list of the member with global rank m is
 *     (topic[i], partition[i])  for i in [member_off[m], member_off[m+1])
 * with topic[] = index of the topic in the batch, in the order the reference appends them
 * (Main:264: topics in batch order, inside a topic in assignment order). n_members =
 * 1 + the largest rank in cset_rank; member_off has n_members + 1 entries; topic and
 * partition have one entry per assigned partition (lagassign_stats.n_assigned). A JVM caller
 * only has to wrap slices instead of scattering 8 bytes per partition into per-member lists. */
typedef struct lagassign_grouping {
    int64_t *member_off;   /* [n_members + 1] */
    int32_t *topic;        /* [n_assigned] */
    int32_t *partition;    /* [n_assigned] */
} lagassign_grouping;
int  lagassign_group_sizes(const lagassign_ctx *ctx, int64_t *n_members, int64_t *n_assigned);
int  lagassign_download_grouping(lagassign_ctx *ctx, const lagassign_grouping *out);

/* Per-kernel device times of the last execute() when LAGASSIGN_OPT_PROFILE is on
 * (CUDA events on the context's stream around each launch). Entries are kernel
 * kinds; a kind launched several times (radix passes) accumulates. Waits for the
 * stream. Returns the number of kinds in *n_out; names are static strings. */
int  lagassign_profile_count(lagassign_ctx *ctx, int *n_out);
int  lagassign_profile_entry(lagassign_ctx *ctx, int i, const char **name_out,
                             float *ms_out, int *launches_out);

/* Pinned host memory, so uploads/downloads run at full PCIe speed. A JNI
 * caller wraps these with NewDirectByteBuffer. */
int  lagassign_host_alloc(void **ptr, size_t bytes);
int  lagassign_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* LAGASSIGN_H */
