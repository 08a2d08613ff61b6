#include "lag_based_partition_assignor.hpp"

#include <algorithm>
#include <cctype>
#include <numeric>
#include <stdexcept>

#include "../../../include/lagassign.h"
#include "java_compat.hpp"

namespace lagassign_host {

LagBasedPartitionAssignor::LagBasedPartitionAssignor(int device) : device_(device) {}

LagBasedPartitionAssignor::~LagBasedPartitionAssignor() {
    if (ctx_) lagassign_destroy(ctx_);
}

void LagBasedPartitionAssignor::configure(const std::map<std::string, std::string>& configs) {
    consumer_group_props_ = configs;                                   // Main:101-104
    auto it = consumer_group_props_.find("group.id");                  // Main:107
    if (it == consumer_group_props_.end()) {
        throw std::invalid_argument(                                   // Main:108-113
            "group.id cannot be null when using partition.assignment.strategy="
            "com.github.grantneale.kafka.LagBasedPartitionAssignor");
    }
    metadata_props_ = consumer_group_props_;                           // Main:116-117
    metadata_props_["enable.auto.commit"] = "false";                   // Main:118
    metadata_props_["client.id"] = it->second + ".assignor";           // Main:119-120
}

int64_t LagBasedPartitionAssignor::computePartitionLag(const int64_t* committed, int64_t beginOffset,
                                                       int64_t endOffset, const std::string& mode) {
    int64_t next;
    if (committed != nullptr) {
        next = *committed;                                             // Main:386
    } else {
        std::string m = mode;
        std::transform(m.begin(), m.end(), m.begin(), [](unsigned char c) { return (char)std::tolower(c); });
        next = (m == "latest") ? endOffset : beginOffset;              // Main:391-396
    }
    const int64_t d = (int64_t)((uint64_t)endOffset - (uint64_t)next); // Java long '-' wraps
    return d > 0 ? d : 0;                                              // Main:402
}

Assignment LagBasedPartitionAssignor::assign(const PartitionLagPerTopic& lags, const Subscriptions& subscriptions,
                                             std::vector<std::string>* topicOrderOut) {
    // Main:171-174: every member gets a list, even with no topics.
    Assignment result;
    result.reserve(subscriptions.size());
    std::unordered_map<std::string, int> member_index;
    for (const auto& e : subscriptions) {
        if (member_index.emplace(e.first, (int)result.size()).second) result.push_back({e.first, {}});
    }
    const int M = (int)result.size();

    // rank members by String.compareTo (Main:258-259)
    std::vector<std::vector<uint16_t>> units(M);
    for (int i = 0; i < M; ++i) units[i] = utf16_units(result[i].first);
    std::vector<int> by_rank(M);
    std::iota(by_rank.begin(), by_rank.end(), 0);
    std::sort(by_rank.begin(), by_rank.end(), [&](int a, int b) { return java_compare_to(units[a], units[b]) < 0; });
    std::vector<int32_t> rank_of(M);
    for (int r = 0; r < M; ++r) rank_of[by_rank[r]] = r;

    // consumersPerTopic (Main:410-426) and its HashMap iteration order (Main:177)
    JavaHashMapOrder topic_map;
    std::vector<std::vector<int32_t>> topic_members;   // ranks, per topic id
    for (const auto& e : subscriptions) {
        const int32_t r = rank_of[member_index[e.first]];
        for (const std::string& topic : e.second) {
            const int id = topic_map.put_if_absent(topic);
            if (id == (int)topic_members.size()) topic_members.emplace_back();
            topic_members[id].push_back(r);                             // Main:419
        }
    }
    const std::vector<int> order = topic_map.order();
    const int64_t T = (int64_t)order.size();

    // flatten: topics in iteration order, consumer sets de-duplicated
    std::vector<int64_t> topic_off(T + 1, 0);
    std::vector<int32_t> topic_cset(T);
    std::vector<int64_t> cset_off{0};
    std::vector<int32_t> cset_rank;
    std::map<std::vector<int32_t>, int32_t> set_ids;
    int64_t N = 0;
    for (int64_t k = 0; k < T; ++k) {
        const int id = order[k];
        auto it = lags.find(topic_map.key(id));                         // Main:182 getOrDefault(.., emptyList)
        N += it == lags.end() ? 0 : (int64_t)it->second.size();
        topic_off[k + 1] = N;
        std::vector<int32_t> set = topic_members[id];
        std::sort(set.begin(), set.end());
        set.erase(std::unique(set.begin(), set.end()), set.end());      // a member listing a topic twice is one map key (Main:216-219)
        auto ins = set_ids.emplace(set, (int32_t)set_ids.size());
        if (ins.second) {
            cset_rank.insert(cset_rank.end(), set.begin(), set.end());
            cset_off.push_back((int64_t)cset_rank.size());
        }
        topic_cset[k] = ins.first->second;
    }
    std::vector<int64_t> lag((size_t)N);
    std::vector<int32_t> part((size_t)N);
    bool implicit_ids = true;
    for (int64_t k = 0; k < T; ++k) {
        auto it = lags.find(topic_map.key(order[k]));
        if (it == lags.end()) continue;
        int64_t o = topic_off[k];
        int32_t expect = 0;
        for (const TopicPartitionLag& e : it->second) {
            lag[o] = e.lag;
            part[o] = e.partition;
            implicit_ids = implicit_ids && e.partition == expect;
            ++expect; ++o;
        }
    }
    if (topicOrderOut) {
        topicOrderOut->clear();
        for (int64_t k = 0; k < T; ++k) topicOrderOut->push_back(topic_map.key(order[k]));
    }
    if (T == 0 || N == 0) return result;

    // the GPU call that replaces Main:149 / the loop Main:177-184
    if (!ctx_) {
        int rc = lagassign_create(&ctx_, device_);
        if (rc != LAGASSIGN_OK) {
            ctx_ = nullptr;
            throw std::runtime_error(std::string("lagassign_create failed: ") + lagassign_last_error(nullptr));
        }
    }
    std::vector<int32_t> out_part((size_t)N), out_cons((size_t)N);
    lagassign_batch in{};
    in.n_topics = T;
    in.topic_off = topic_off.data();
    in.lag = lag.data();
    in.partition = implicit_ids ? nullptr : part.data();
    in.topic_cset = topic_cset.data();
    in.n_csets = (int64_t)cset_off.size() - 1;
    in.cset_off = cset_off.data();
    in.cset_rank = cset_rank.data();
    lagassign_result out{};
    out.out_partition = out_part.data();
    out.out_consumer_rank = out_cons.data();
    int rc = lagassign_assign_batch(ctx_, &in, &out);
    if (rc != LAGASSIGN_OK)
        throw std::runtime_error(std::string("lagassign_assign_batch failed: ") + lagassign_last_error(ctx_));

    // unflatten in topic order, then assignment order (Main:264 appends)
    for (int64_t k = 0; k < T; ++k) {
        const std::string& topic = topic_map.key(order[k]);
        for (int64_t i = topic_off[k]; i < topic_off[k + 1]; ++i) {
            const int32_t r = out_cons[(size_t)i];
            if (r < 0) continue;
            result[(size_t)by_rank[(size_t)r]].second.push_back({topic, out_part[(size_t)i]});
        }
    }
    return result;
}

}  // namespace lagassign_host
