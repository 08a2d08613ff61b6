/*
 * oracle.h — CPU restatement of the reference's lag-based assignment.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path (the package
 * kafka-lag-based-assignor_b200/ or include/lagassign.h) links, loads or calls
 * this. It is used by tests/, by __graft_entry__.smoke() and by bench.py's
 * cpu_baseline / --impl reference legs, as the checker and as the timed CPU
 * baseline.
 *
 * Reference: /root/reference/src/main/java/com/github/grantneale/kafka/
 * LagBasedPartitionAssignor.java ("Main" below). The reference is a JVM class;
 * no JDK exists in this image, so it cannot be executed here (no oracle/_ref).
 * Parity is pinned by the reference's own test vectors instead: Test:82-132
 * (full, order-sensitive assignment), README:42-57 / Main:57-69 (worked
 * example) and Test:21-80 (computePartitionLag); see tests/golden/.
 *
 * Flat layout is the one of include/lagassign.h.
 */
#ifndef LAGASSIGN_ORACLE_H
#define LAGASSIGN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_LITERAL 0  /* O(P log P + P*C): Main:228-277 statement by statement */
#define ORACLE_ROUNDS  1  /* O(P log P + P log C): equivalent round formulation   */

/* One topic. part may be NULL (ids 0..P-1). crank: the topic's consumers as
 * ascending global ranks. out_total/out_count (size C) may be NULL.
 * Returns 0, or -1 on allocation failure. */
int oracle_assign_topic(int mode, int64_t P, const int64_t *lag, const int32_t *part,
                        int64_t C, const int32_t *crank,
                        int32_t *out_part, int32_t *out_cons,
                        int64_t *out_total, int32_t *out_count);

/* Whole batch (Main:166-188 with the maps flattened). threads <= 1 runs the
 * topics sequentially like the reference (Main:177-184); threads > 1 spreads
 * topics over OpenMP threads. topic_lo/topic_hi restrict the run to a range of
 * topics (outputs of other topics are left untouched) so a bounded sample can
 * be timed. Returns the number of partitions that received a consumer, or -1. */
int64_t oracle_assign_batch(int mode, int threads,
                            int64_t n_topics, const int64_t *topic_off,
                            const int64_t *lag, const int32_t *partition,
                            const int32_t *topic_cset, int64_t n_csets,
                            const int64_t *cset_off, const int32_t *cset_rank,
                            int64_t topic_lo, int64_t topic_hi,
                            int32_t *out_partition, int32_t *out_consumer_rank,
                            int64_t *out_total, int32_t *out_count);

/* Main:376-404. reset_mode is compared with equalsIgnoreCase("latest"). */
int64_t oracle_compute_partition_lag(int has_committed, int64_t committed,
                                     int64_t begin_offset, int64_t end_offset,
                                     const char *reset_mode);

/* java.lang.String.compareTo on two UTF-8 encoded ids: lexicographic over
 * UTF-16 code units, then length (Main:258-259). Returns <0, 0, >0. */
int oracle_compare_member_ids(const char *a_utf8, const char *b_utf8);

/* rank_out[i] = position of ids[i] in String.compareTo order (ids distinct). */
int oracle_rank_members(int64_t n, const char *const *ids_utf8, int32_t *rank_out);

int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
