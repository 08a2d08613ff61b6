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

FlatBatch LagBasedPartitionAssignor::flatten(const PartitionLagPerTopic& lags, const Subscriptions& subscriptions) {
    FlatBatch fb;
    // Main:171-174: every member gets a list, even with no topics.
    std::unordered_map<std::string, int> member_index;
    std::vector<int> entry_member(subscriptions.size());
    for (size_t i = 0; i < subscriptions.size(); ++i) {
        auto ins = member_index.emplace(subscriptions[i].first, (int)fb.members.size());
        if (ins.second) fb.members.push_back(subscriptions[i].first);
        entry_member[i] = ins.first->second;
    }
    const int M = (int)fb.members.size();

    // rank members by String.compareTo (Main:258-259)
    std::vector<std::vector<uint16_t>> units(M);
    for (int i = 0; i < M; ++i) units[i] = utf16_units(fb.members[i]);
    fb.by_rank.resize(M);
    std::iota(fb.by_rank.begin(), fb.by_rank.end(), 0);
    std::sort(fb.by_rank.begin(), fb.by_rank.end(),
              [&](int32_t a, int32_t b) { return java_compare_to(units[a], units[b]) < 0; });

    // consumersPerTopic (Main:410-426). Pass 1, in the iteration order of `subscriptions`: the
    // HashMap's keys (its iteration order, Main:177, depends on the order of insertion).
    JavaHashMapOrder topic_map;
    std::vector<std::vector<int>> member_topics(M);     // topic ids per member, as listed
    for (size_t i = 0; i < subscriptions.size(); ++i) {
        auto& mine = member_topics[entry_member[i]];
        for (const std::string& topic : subscriptions[i].second) mine.push_back(topic_map.put_if_absent(topic));   // Main:419
    }
    const std::vector<int> order = topic_map.order();
    const int64_t T = (int64_t)order.size();
    // Pass 2, members in RANK order: every topic's member list comes out ascending in rank, and
    // a member that lists a topic twice (one map key, Main:216-219) is dropped by looking at the tail.
    std::vector<std::vector<int32_t>> topic_members((size_t)T);   // by topic id
    for (int32_t r = 0; r < M; ++r) {
        for (int id : member_topics[fb.by_rank[r]]) {
            auto& v = topic_members[(size_t)id];
            if (v.empty() || v.back() != r) v.push_back(r);
        }
    }

    // topics in iteration order, consumer sets de-duplicated (hash of the rank list, then compare)
    fb.topics.reserve((size_t)T);
    fb.topic_off.assign((size_t)T + 1, 0);
    fb.topic_cset.resize((size_t)T);
    fb.cset_off.assign(1, 0);
    std::unordered_map<uint64_t, std::vector<int32_t>> sets_by_hash;   // hash -> set ids
    int64_t N = 0;
    for (int64_t k = 0; k < T; ++k) {
        const int id = order[(size_t)k];
        fb.topics.push_back(topic_map.key(id));
        auto it = lags.find(fb.topics.back());                          // Main:182 getOrDefault(.., emptyList)
        N += it == lags.end() ? 0 : (int64_t)it->second.size();
        fb.topic_off[(size_t)k + 1] = N;
        const std::vector<int32_t>& set = topic_members[(size_t)id];
        uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)set.size();
        for (int32_t r : set) { h ^= (uint64_t)(uint32_t)r; h *= 0x100000001b3ull; h ^= h >> 29; }
        std::vector<int32_t>& candidates = sets_by_hash[h];
        int32_t found = -1;
        for (int32_t sid : candidates) {
            const int64_t lo = fb.cset_off[(size_t)sid], hi = fb.cset_off[(size_t)sid + 1];
            if (hi - lo == (int64_t)set.size() && std::equal(set.begin(), set.end(), fb.cset_rank.begin() + lo)) { found = sid; break; }
        }
        if (found < 0) {
            found = (int32_t)fb.cset_off.size() - 1;
            candidates.push_back(found);
            fb.cset_rank.insert(fb.cset_rank.end(), set.begin(), set.end());
            fb.cset_off.push_back((int64_t)fb.cset_rank.size());
        }
        fb.topic_cset[(size_t)k] = found;
    }
    fb.lag.resize((size_t)N);
    fb.partition.resize((size_t)N);
    fb.implicit_ids = true;
    for (int64_t k = 0; k < T; ++k) {
        auto it = lags.find(fb.topics[(size_t)k]);
        if (it == lags.end()) continue;
        int64_t o = fb.topic_off[(size_t)k];
        int32_t expect = 0;
        for (const TopicPartitionLag& e : it->second) {
            fb.lag[(size_t)o] = e.lag;
            fb.partition[(size_t)o] = e.partition;
            fb.implicit_ids = fb.implicit_ids && e.partition == expect;
            ++expect; ++o;
        }
    }
    return fb;
}

Assignment LagBasedPartitionAssignor::unflatten(const FlatBatch& fb, const int32_t* out_partition,
                                                const int32_t* out_consumer_rank) {
    Assignment result;
    result.reserve(fb.members.size());
    for (const std::string& m : fb.members) result.push_back({m, {}});
    const int64_t T = (int64_t)fb.topics.size();
    if (T == 0 || fb.topic_off[(size_t)T] == 0) return result;
    // topic order, then assignment order (Main:264 appends)
    for (int64_t k = 0; k < T; ++k) {
        const std::string& topic = fb.topics[(size_t)k];
        for (int64_t i = fb.topic_off[(size_t)k]; i < fb.topic_off[(size_t)k + 1]; ++i) {
            const int32_t r = out_consumer_rank[(size_t)i];
            if (r < 0) continue;
            result[(size_t)fb.by_rank[(size_t)r]].second.push_back({topic, out_partition[(size_t)i]});
        }
    }
    return result;
}

Assignment LagBasedPartitionAssignor::assign(const PartitionLagPerTopic& lags, const Subscriptions& subscriptions,
                                             std::vector<std::string>* topicOrderOut) {
    const FlatBatch fb = flatten(lags, subscriptions);
    if (topicOrderOut) *topicOrderOut = fb.topics;
    const int64_t T = (int64_t)fb.topics.size();
    const int64_t N = fb.topic_off[(size_t)T];
    if (T == 0 || N == 0) return unflatten(fb, nullptr, nullptr);

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
    in.topic_off = fb.topic_off.data();
    in.lag = fb.lag.data();
    in.partition = fb.implicit_ids ? nullptr : fb.partition.data();
    in.topic_cset = fb.topic_cset.data();
    in.n_csets = (int64_t)fb.cset_off.size() - 1;
    in.cset_off = fb.cset_off.data();
    in.cset_rank = fb.cset_rank.data();
    lagassign_result out{};
    out.out_partition = out_part.data();
    out.out_consumer_rank = out_cons.data();
    int rc = lagassign_assign_batch(ctx_, &in, &out);
    if (rc != LAGASSIGN_OK)
        throw std::runtime_error(std::string("lagassign_assign_batch failed: ") + lagassign_last_error(ctx_));
    return unflatten(fb, out_part.data(), out_cons.data());
}

Assignment LagBasedPartitionAssignor::assignFromOffsets(const PartitionOffsetsPerTopic& offsets, const std::string& autoOffsetReset,
                                                        const Subscriptions& subscriptions, std::vector<std::string>* topicOrderOut) {
    // layout (member ranks, topic order, consumer sets, partition ids) exactly as for lags; the lag column stays empty
    PartitionLagPerTopic shape;
    for (const auto& kv : offsets) {
        std::vector<TopicPartitionLag>& v = shape[kv.first];
        v.reserve(kv.second.size());
        for (const TopicPartitionOffsets& e : kv.second) v.push_back({e.topic, e.partition, 0});
    }
    const FlatBatch fb = flatten(shape, subscriptions);
    if (topicOrderOut) *topicOrderOut = fb.topics;
    const int64_t T = (int64_t)fb.topics.size();
    const int64_t N = fb.topic_off[(size_t)T];
    if (T == 0 || N == 0) return unflatten(fb, nullptr, nullptr);
    std::vector<int64_t> begin((size_t)N), end((size_t)N), committed((size_t)N);
    std::vector<uint8_t> has((size_t)N);
    for (int64_t k = 0; k < T; ++k) {
        auto it = offsets.find(fb.topics[(size_t)k]);
        if (it == offsets.end()) continue;
        int64_t o = fb.topic_off[(size_t)k];
        for (const TopicPartitionOffsets& e : it->second) {
            begin[(size_t)o] = e.beginOffset;
            end[(size_t)o] = e.endOffset;
            committed[(size_t)o] = e.hasCommitted ? e.committedOffset : 0;
            has[(size_t)o] = e.hasCommitted ? 1 : 0;
            ++o;
        }
    }
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
    in.topic_off = fb.topic_off.data();
    in.lag = nullptr;
    in.partition = fb.implicit_ids ? nullptr : fb.partition.data();
    in.topic_cset = fb.topic_cset.data();
    in.n_csets = (int64_t)fb.cset_off.size() - 1;
    in.cset_off = fb.cset_off.data();
    in.cset_rank = fb.cset_rank.data();
    lagassign_offsets offs{};
    offs.begin_offset = begin.data();
    offs.end_offset = end.data();
    offs.committed = committed.data();
    offs.has_committed = has.data();
    std::string mode = autoOffsetReset;
    for (char& c : mode) c = (char)std::tolower((unsigned char)c);
    offs.reset_latest = mode == "latest" ? 1 : 0;                     // Main:391 equalsIgnoreCase("latest"), else "assume earliest"
    lagassign_result out{};
    out.out_partition = out_part.data();
    out.out_consumer_rank = out_cons.data();
    int rc = lagassign_upload_offsets(ctx_, &in, &offs);
    if (rc == LAGASSIGN_OK) rc = lagassign_execute(ctx_);
    if (rc == LAGASSIGN_OK) rc = lagassign_download(ctx_, &out);
    if (rc != LAGASSIGN_OK)
        throw std::runtime_error(std::string("assignment from offsets failed: ") + lagassign_last_error(ctx_));
    return unflatten(fb, out_part.data(), out_cons.data());
}

}  // namespace lagassign_host
