// lag_based_partition_assignor.hpp — C++ host mirror of the reference plugin
// class com.github.grantneale.kafka.LagBasedPartitionAssignor for the assign()
// path: same names, same argument meaning, same error behaviour; the per-topic
// arithmetic (LagBasedPartitionAssignor.java:204-277) runs on the GPU through the
// C ABI of include/lagassign.h. There is no CPU implementation of it here.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

struct lagassign_ctx;

namespace lagassign_host {

struct TopicPartition {            // org.apache.kafka.common.TopicPartition
    std::string topic;
    int32_t partition;
};

struct TopicPartitionLag {         // Main:431-455
    std::string topic;
    int32_t partition;
    int64_t lag;
};

// What readTopicPartitionLags gathers per partition before it calls computePartitionLag (Main:338-356): the begin
// and end offsets (0 when the broker did not report them, Main:350-351) and the committed offset if there is one.
struct TopicPartitionOffsets {
    std::string topic;
    int32_t partition;
    int64_t beginOffset, endOffset, committedOffset;
    bool hasCommitted;              // partitionMetadata != null (Main:384)
};
using PartitionOffsetsPerTopic = std::unordered_map<std::string, std::vector<TopicPartitionOffsets>>;

// Map<String, List<String>> member -> topics, as its entries in ITERATION ORDER
// (Main:168; the order matters through consumersPerTopic, Main:413).
using Subscriptions = std::vector<std::pair<std::string, std::vector<std::string>>>;
// Map<String, List<TopicPartitionLag>> (Main:167); only looked up by key (Main:182).
using PartitionLagPerTopic = std::unordered_map<std::string, std::vector<TopicPartitionLag>>;
// Map<String, List<TopicPartition>> member -> assignment (Main:171), in the
// order of `subscriptions`.
using Assignment = std::vector<std::pair<std::string, std::vector<TopicPartition>>>;

// The arguments of lagassign_assign_batch() for one assign() call, plus what is needed to
// turn the flat result back into per-member lists. Built by flatten(), consumed by unflatten().
struct FlatBatch {
    std::vector<std::string> members;     // result order = order of `subscriptions` (first occurrence)
    std::vector<int32_t> by_rank;         // String.compareTo rank -> index into members
    std::vector<std::string> topics;      // assignment order = HashMap iteration order (Main:177)
    std::vector<int64_t> topic_off;       // [T+1]
    std::vector<int64_t> lag;             // [N]
    std::vector<int32_t> partition;       // [N]
    bool implicit_ids = true;             // every topic lists partitions 0..P-1 in order
    std::vector<int32_t> topic_cset;      // [T]
    std::vector<int64_t> cset_off;        // [S+1]
    std::vector<int32_t> cset_rank;       // member ranks, ascending inside a set
};

class LagBasedPartitionAssignor {
 public:
    explicit LagBasedPartitionAssignor(int device = 0);
    ~LagBasedPartitionAssignor();
    LagBasedPartitionAssignor(const LagBasedPartitionAssignor&) = delete;
    LagBasedPartitionAssignor& operator=(const LagBasedPartitionAssignor&) = delete;

    // Configurable.configure (Main:97-130). Throws std::invalid_argument when
    // group.id is missing (IllegalArgumentException at Main:108-113).
    void configure(const std::map<std::string, std::string>& configs);
    const std::map<std::string, std::string>& metadataConsumerProps() const { return metadata_props_; }

    // ConsumerPartitionAssignor.name (Main:132-135).
    std::string name() const { return "lag"; }

    // static assign(Map, Map) (Main:166-188). Throws std::runtime_error when the
    // GPU path fails (no fallback). `topicOrderOut`, if given, receives the
    // topics in the order they were assigned.
    Assignment assign(const PartitionLagPerTopic& partitionLagPerTopic, const Subscriptions& subscriptions,
                      std::vector<std::string>* topicOrderOut = nullptr);

    // The same assignment from RAW OFFSETS: readTopicPartitionLags' loop body (Main:344-356) moves to the device —
    // every partition's lag is computePartitionLag(committed?, begin, end, autoOffsetReset) (Main:376-404), evaluated
    // by the sort kernels as they build their keys (lagassign_upload_offsets); no lag exists on the host.
    Assignment assignFromOffsets(const PartitionOffsetsPerTopic& offsetsPerTopic, const std::string& autoOffsetReset,
                                 const Subscriptions& subscriptions, std::vector<std::string>* topicOrderOut = nullptr);

    // The two host-side halves of assign(): everything before and after the device call.
    // flatten: member ranks (Main:258-259), consumersPerTopic and its iteration order
    // (Main:410-426, :177), consumer sets de-duplicated, lags/ids laid out flat.
    // unflatten: Main:264's appends, topics in assignment order, inside a topic in
    // assignment order; out_consumer_rank < 0 = not assigned.
    static FlatBatch flatten(const PartitionLagPerTopic& partitionLagPerTopic, const Subscriptions& subscriptions);
    static Assignment unflatten(const FlatBatch& flat, const int32_t* out_partition, const int32_t* out_consumer_rank);

    // static computePartitionLag (Main:376-404); committed == nullptr means no
    // committed offset (partitionMetadata == null).
    static int64_t computePartitionLag(const int64_t* committed, int64_t beginOffset, int64_t endOffset,
                                       const std::string& autoOffsetResetMode);

 private:
    lagassign_ctx* ctx_ = nullptr;
    int device_;
    std::map<std::string, std::string> consumer_group_props_;
    std::map<std::string, std::string> metadata_props_;
};

}  // namespace lagassign_host
