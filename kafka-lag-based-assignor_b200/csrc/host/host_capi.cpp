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
    Assignment result;
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
    std::map<std::string, std::string> cfg;
    for (int i = 0; i < n; ++i) cfg[keys[i]] = values[i] ? values[i] : "";
    try {
        g->assignor->configure(cfg);
    } catch (const std::invalid_argument& e) {
        return fail(g, LAGASSIGN_ERR_ARG, e.what());
    }
    g->client_id = g->assignor->metadataConsumerProps().at("client.id");
    return LAGASSIGN_OK;
}

const char* lagassign_group_metadata_client_id(const lagassign_group* g) { return g ? g->client_id.c_str() : ""; }

const char* lagassign_group_name(const lagassign_group* g) { return g ? g->name.c_str() : "lag"; }

int lagassign_group_clear(lagassign_group* g) {
    if (!g) return fail(nullptr, LAGASSIGN_ERR_ARG, "group is NULL");
    g->subscriptions.clear(); g->lags.clear(); g->result.clear(); g->result_index.clear(); g->topic_order.clear();
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

int lagassign_group_assign(lagassign_group* g) {
    if (!g) return fail(nullptr, LAGASSIGN_ERR_ARG, "group is NULL");
    try {
        g->result = g->assignor->assign(g->lags, g->subscriptions, &g->topic_order);
    } catch (const std::exception& e) {
        return fail(g, LAGASSIGN_ERR_CUDA, e.what());
    }
    g->result_index.clear();
    for (size_t i = 0; i < g->result.size(); ++i) g->result_index[g->result[i].first] = i;
    return LAGASSIGN_OK;
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
