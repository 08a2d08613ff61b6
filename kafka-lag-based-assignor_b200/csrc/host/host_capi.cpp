// host_capi.cpp — plain-C view (include/lagassign_host.h) of the C++ host mirror.
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/lagassign.h"
#include "../../../include/lagassign_host.h"
#include "lag_based_partition_assignor.hpp"

using namespace lagassign_host;

struct lagassign_group {
    std::unique_ptr<LagBasedPartitionAssignor> assignor;
    Subscriptions subscriptions;
    PartitionLagPerTopic lags;
    PartitionOffsetsPerTopic offsets;
    std::string auto_offset_reset = "latest";       // ConsumerConfig default, Main:346-347
    Assignment result;
    FlatBatch flat;
    bool flat_valid = false;
    std::map<std::string, size_t> result_index;
    std::vector<std::string> topic_order;
    std::string err, client_id, name;
};

namespace {
thread_local std::string g_err;
int fail(lagassign_group* g, int code, const std::string& msg) {
    if (g) g->err = msg;
    g_err = msg;
    return code;
}
}  // namespace

extern "C" {

int lagassign_group_create(lagassign_group** g, int device) {
    if (!g) return fail(nullptr, LAGASSIGN_ERR_ARG, "group out-pointer is NULL");
    *g = nullptr;
    try {
        auto* grp = new lagassign_group();
        grp->assignor.reset(new LagBasedPartitionAssignor(device));
        grp->name = grp->assignor->name();
        *g = grp;
        return LAGASSIGN_OK;
    } catch (const std::exception& e) {
        return fail(nullptr, LAGASSIGN_ERR_OOM, e.what());
    }
}

void lagassign_group_destroy(lagassign_group* g) { delete g; }

const char* lagassign_group_last_error(const lagassign_group* g) { return g ? g->err.c_str() : g_err.c_str(); }

int lagassign_group_configure(lagassign_group* g, const char* const* keys, const char* const* values, int n) {
    if (!g) return fail(nullptr, LAGASSIGN_ERR_ARG, "group is NULL");
    if (n < 0 || (n > 0 && (!keys || !values))) return fail(g, LAGASSIGN_ERR_ARG, "keys / values is NULL");
    try {
        std::map<std::string, std::string> cfg;
        for (int i = 0; i < n; ++i) {
            if (!keys[i]) return fail(g, LAGASSIGN_ERR_ARG, "config key is NULL");
            if (values[i]) cfg[keys[i]] = values[i];      // a NULL value is a missing key (Main:107-113 rejects a null group.id)
        }
        g->assignor->configure(cfg);
        g->client_id = g->assignor->metadataConsumerProps().at("client.id");
    } catch (const std::invalid_argument& e) {
        return fail(g, LAGASSIGN_ERR_ARG, e.what());
    } catch (const std::exception& e) {                   // nothing may cross the extern "C" boundary
        return fail(g, LAGASSIGN_ERR_ARG, e.what());
    }
    return LAGASSIGN_OK;
}

const char* lagassign_group_metadata_client_id(const lagassign_group* g) { return g ? g->client_id.c_str() : ""; }

const char* lagassign_group_name(const lagassign_group* g) { return g ? g->name.c_str() : "lag"; }

int lagassign_group_clear(lagassign_group* g) {
    if (!g) return fail(nullptr, LAGASSIGN_ERR_ARG, "group is NULL");
    g->subscriptions.clear(); g->lags.clear(); g->offsets.clear(); g->result.clear(); g->result_index.clear(); g->topic_order.clear();
    g->flat = FlatBatch(); g->flat_valid = false;
    return LAGASSIGN_OK;
}

int lagassign_group_add_subscription(lagassign_group* g, const char* member_id, const char* const* topics, int n_topics) {
    if (!g || !member_id) return fail(g, LAGASSIGN_ERR_ARG, "group or member id is NULL");
    std::vector<std::string> ts;
    for (int i = 0; i < n_topics; ++i) ts.emplace_back(topics[i]);
    g->subscriptions.emplace_back(member_id, std::move(ts));
    return LAGASSIGN_OK;
}

int lagassign_group_add_topic_lags(lagassign_group* g, const char* topic, const int32_t* partitions,
                                   const int64_t* lags, int64_t n) {
    if (!g || !topic || (n > 0 && !lags)) return fail(g, LAGASSIGN_ERR_ARG, "group, topic or lags is NULL");
    std::vector<TopicPartitionLag> v;
    v.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) v.push_back({topic, partitions ? partitions[i] : (int32_t)i, lags[i]});
    g->lags[topic] = std::move(v);
    return LAGASSIGN_OK;
}

int lagassign_group_add_topic_offsets(lagassign_group* g, const char* topic, const int32_t* partitions,
                                      const int64_t* begin_offset, const int64_t* end_offset, const int64_t* committed,
                                      const uint8_t* has_committed, int64_t n) {
    if (!g || !topic || (n > 0 && (!begin_offset || !end_offset || !committed || !has_committed)))
        return fail(g, LAGASSIGN_ERR_ARG, "group, topic or an offset array is NULL");
    std::vector<TopicPartitionOffsets> v;
    v.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i)
        v.push_back({topic, partitions ? partitions[i] : (int32_t)i, begin_offset[i], end_offset[i], committed[i], has_committed[i] != 0});
    g->offsets[topic] = std::move(v);
    return LAGASSIGN_OK;
}

int lagassign_group_set_auto_offset_reset(lagassign_group* g, const char* mode) {
    if (!g || !mode) return fail(g, LAGASSIGN_ERR_ARG, "group or mode is NULL");
    g->auto_offset_reset = mode;
    return LAGASSIGN_OK;
}

int lagassign_group_assign(lagassign_group* g) {
    if (!g) return fail(nullptr, LAGASSIGN_ERR_ARG, "group is NULL");
    if (!g->offsets.empty() && !g->lags.empty())
        return fail(g, LAGASSIGN_ERR_ARG, "a group takes either lags or raw offsets, not both");
    try {
        if (!g->offsets.empty())
            g->result = g->assignor->assignFromOffsets(g->offsets, g->auto_offset_reset, g->subscriptions, &g->topic_order);
        else
        g->result = g->assignor->assign(g->lags, g->subscriptions, &g->topic_order);
    } catch (const std::exception& e) {
        return fail(g, LAGASSIGN_ERR_CUDA, e.what());
    }
    g->result_index.clear();
    for (size_t i = 0; i < g->result.size(); ++i) g->result_index[g->result[i].first] = i;
    return LAGASSIGN_OK;
}

int lagassign_group_flatten(lagassign_group* g, lagassign_batch* batch_out) {
    if (!g || !batch_out) return fail(g, LAGASSIGN_ERR_ARG, "group or batch_out is NULL");
    g->flat_valid = false;
    try {
        g->flat = LagBasedPartitionAssignor::flatten(g->lags, g->subscriptions);
    } catch (const std::exception& e) {
        return fail(g, LAGASSIGN_ERR_UNSUPPORTED, e.what());
    }
    g->flat_valid = true;
    g->topic_order = g->flat.topics;
    const FlatBatch& fb = g->flat;
    std::memset(batch_out, 0, sizeof *batch_out);
    batch_out->n_topics = (int64_t)fb.topics.size();
    batch_out->topic_off = fb.topic_off.data();
    batch_out->lag = fb.lag.data();
    batch_out->partition = fb.implicit_ids ? nullptr : fb.partition.data();
    batch_out->topic_cset = fb.topic_cset.data();
    batch_out->n_csets = (int64_t)fb.cset_off.size() - 1;
    batch_out->cset_off = fb.cset_off.data();
    batch_out->cset_rank = fb.cset_rank.data();
    return LAGASSIGN_OK;
}

int lagassign_group_unflatten(lagassign_group* g, const int32_t* out_partition, const int32_t* out_consumer_rank) {
    if (!g) return fail(nullptr, LAGASSIGN_ERR_ARG, "group is NULL");
    if (!g->flat_valid) return fail(g, LAGASSIGN_ERR_STATE, "unflatten() before flatten()");
    const int64_t N = g->flat.topic_off.empty() ? 0 : g->flat.topic_off.back();
    if (N > 0 && (!out_partition || !out_consumer_rank)) return fail(g, LAGASSIGN_ERR_ARG, "output arrays are NULL");
    const int32_t M = (int32_t)g->flat.members.size();
    for (int64_t i = 0; i < N; ++i)
        if (out_consumer_rank[i] >= M) return fail(g, LAGASSIGN_ERR_ARG, "out_consumer_rank holds a rank beyond the member count");
    g->result = LagBasedPartitionAssignor::unflatten(g->flat, out_partition, out_consumer_rank);
    g->result_index.clear();
    for (size_t i = 0; i < g->result.size(); ++i) g->result_index[g->result[i].first] = i;
    return LAGASSIGN_OK;
}

int64_t lagassign_group_member_count(const lagassign_group* g) { return (g && g->flat_valid) ? (int64_t)g->flat.members.size() : 0; }

const char* lagassign_group_member_at_rank(const lagassign_group* g, int64_t rank) {
    if (!g || !g->flat_valid || rank < 0 || rank >= (int64_t)g->flat.by_rank.size()) return nullptr;
    return g->flat.members[(size_t)g->flat.by_rank[(size_t)rank]].c_str();
}

int64_t lagassign_group_result_size(const lagassign_group* g, const char* member_id) {
    if (!g || !member_id) return -1;
    auto it = g->result_index.find(member_id);
    return it == g->result_index.end() ? -1 : (int64_t)g->result[it->second].second.size();
}

int lagassign_group_result_get(const lagassign_group* g, const char* member_id, int64_t i,
                               const char** topic_out, int32_t* partition_out) {
    if (!g || !member_id) return LAGASSIGN_ERR_ARG;
    auto it = g->result_index.find(member_id);
    if (it == g->result_index.end()) return LAGASSIGN_ERR_ARG;
    const auto& v = g->result[it->second].second;
    if (i < 0 || i >= (int64_t)v.size()) return LAGASSIGN_ERR_ARG;
    if (topic_out) *topic_out = v[(size_t)i].topic.c_str();
    if (partition_out) *partition_out = v[(size_t)i].partition;
    return LAGASSIGN_OK;
}

int64_t lagassign_group_topic_count(const lagassign_group* g) { return g ? (int64_t)g->topic_order.size() : 0; }

const char* lagassign_group_topic_at(const lagassign_group* g, int64_t i) {
    if (!g || i < 0 || i >= (int64_t)g->topic_order.size()) return nullptr;
    return g->topic_order[(size_t)i].c_str();
}

int64_t lagassign_compute_partition_lag(int has_committed, int64_t committed, int64_t begin_offset,
                                        int64_t end_offset, const char* mode) {
    return LagBasedPartitionAssignor::computePartitionLag(has_committed ? &committed : nullptr, begin_offset,
                                                          end_offset, mode ? mode : "latest");
}

}  // extern "C"
