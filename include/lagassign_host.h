/*
 * lagassign_host.h — C ABI of the host-side mirror of the reference plugin class.
 *
 * The reference's host code is a JVM class (LagBasedPartitionAssignor.java). No
 * JDK exists in the build image, so the host side above lagassign.h is written
 * in C++ (kafka-lag-based-assignor_b200/csrc/host/) with the same names and
 * semantics as the reference's plugin surface for this path:
 *
 *   configure(Map)          Main:97-130   (group.id required -> error)
 *   name() -> "lag"         Main:132-135
 *   assign(Map, Map)        Main:166-188  (maps of strings in, per-member lists out)
 *   computePartitionLag     Main:376-404
 *
 * This header exposes that C++ class through plain C so that tests (ctypes) and
 * non-JVM Kafka clients can drive it. A "group" object plays the role of one
 * assign(partitionLagPerTopic, subscriptions) call: fill in the two maps, call
 * lagassign_group_assign(), read the per-member lists.
 *
 * All strings are UTF-8. Member ids are ordered like java.lang.String.compareTo
 * (UTF-16 code units, Main:258-259). Topics are visited in the iteration order
 * of the java.util.HashMap built by consumersPerTopic (Main:412-424, iterated at
 * Main:177), which is what fixes the order of each member's result list.
 */
#ifndef LAGASSIGN_HOST_H
#define LAGASSIGN_HOST_H

#include <stdint.h>

#include "lagassign.h"   /* lagassign_batch, error codes */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lagassign_group lagassign_group;

/* One assignor instance bound to CUDA device `device` (Main:83-89). */
int  lagassign_group_create(lagassign_group **g, int device);
void lagassign_group_destroy(lagassign_group *g);
const char *lagassign_group_last_error(const lagassign_group *g);

/* Configurable.configure (Main:97-130). Returns LAGASSIGN_ERR_ARG when group.id is
 * absent (the reference throws IllegalArgumentException, Main:107-113). */
int  lagassign_group_configure(lagassign_group *g, const char *const *keys,
                               const char *const *values, int n);
/* client.id derived for the metadata consumer: "<group.id>.assignor" (Main:119-120). */
const char *lagassign_group_metadata_client_id(const lagassign_group *g);

/* ConsumerPartitionAssignor.name() (Main:132-135): "lag". */
const char *lagassign_group_name(const lagassign_group *g);

/* Start a new assign() call: forget maps and results. */
int  lagassign_group_clear(lagassign_group *g);

/* subscriptions.put(member, topics) — entries must be added in the iteration
 * order of the caller's Map (Main:172, :413). */
int  lagassign_group_add_subscription(lagassign_group *g, const char *member_id,
                                      const char *const *topics, int n_topics);

/* partitionLagPerTopic.put(topic, [TopicPartitionLag(topic, partitions[i], lags[i])])
 * (Main:167, :431-455). partitions may be NULL for ids 0..n-1. */
int  lagassign_group_add_topic_lags(lagassign_group *g, const char *topic,
                                    const int32_t *partitions, const int64_t *lags, int64_t n);

/* Raw offsets of a topic instead of its lags: what readTopicPartitionLags gathers per partition (Main:338-356).
 * begin / end: topicBeginOffsets / topicEndOffsets.getOrDefault(p, 0); committed + has_committed: the
 * OffsetAndMetadata of the partition, if any. lagassign_group_assign() then lets the DEVICE compute every lag
 * (computePartitionLag, Main:376-404, inside the sort kernels: lagassign_upload_offsets). A group is filled either
 * with lags or with offsets. auto.offset.reset: "latest" (the default, Main:346-347), anything else = earliest. */
int  lagassign_group_add_topic_offsets(lagassign_group *g, const char *topic, const int32_t *partitions,
                                       const int64_t *begin_offset, const int64_t *end_offset,
                                       const int64_t *committed, const uint8_t *has_committed, int64_t n);
int  lagassign_group_set_auto_offset_reset(lagassign_group *g, const char *mode);

/* assign(partitionLagPerTopic, subscriptions) (Main:166-188) on the GPU. */
int  lagassign_group_assign(lagassign_group *g);

/* The two host-side halves of lagassign_group_assign, for a caller that runs the device step
 * itself (its own context / stream, or several groups in one lagassign_assign_batch) and for
 * the CPU tests of the host logic:
 *   flatten:   ranks members (String.compareTo, Main:258-259), builds consumersPerTopic and its
 *              HashMap iteration order (Main:410-426, :177), de-duplicates consumer sets and lays
 *              lags / ids out flat. *batch_out points into storage owned by the group, valid until
 *              the next flatten / clear / destroy; batch_out->partition is NULL when every topic
 *              lists ids 0..P-1 in order. Needs no GPU.
 *   unflatten: takes the device step's two output arrays (N entries each, as lagassign_result)
 *              for the batch of the last flatten and fills the group's results exactly as
 *              lagassign_group_assign does (Main:264 append order). */
int  lagassign_group_flatten(lagassign_group *g, lagassign_batch *batch_out);
int  lagassign_group_unflatten(lagassign_group *g, const int32_t *out_partition,
                               const int32_t *out_consumer_rank);
/* Members in String.compareTo order: the meaning of the ranks in cset_rank / out_consumer_rank. */
int64_t lagassign_group_member_count(const lagassign_group *g);
const char *lagassign_group_member_at_rank(const lagassign_group *g, int64_t rank);

/* Results: Map<member, List<TopicPartition>>. */
int64_t lagassign_group_result_size(const lagassign_group *g, const char *member_id); /* -1: unknown member */
int  lagassign_group_result_get(const lagassign_group *g, const char *member_id, int64_t i,
                                const char **topic_out, int32_t *partition_out);

/* Topics in the order they were assigned (HashMap iteration order). */
int64_t lagassign_group_topic_count(const lagassign_group *g);
const char *lagassign_group_topic_at(const lagassign_group *g, int64_t i);

/* computePartitionLag (Main:376-404), scalar form. */
int64_t lagassign_compute_partition_lag(int has_committed, int64_t committed, int64_t begin_offset,
                                        int64_t end_offset, const char *auto_offset_reset_mode);

#ifdef __cplusplus
}
#endif
#endif
