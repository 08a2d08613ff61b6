// emu_onchip.cpp — TEST INFRASTRUCTURE: the on-chip kernels of the product (k_warp_topics,
// k_cta_topics<EPT>; kernels_warp.cuh / kernels_cta.cuh and everything they include) compiled
// for the HOST with the stand-in of cuda_emu.hpp and run one CTA at a time. Built and driven by
// tests/test_kernels_emulated_cpu.py; never part of the library.
#define LAGASSIGN_HOST_EMULATION 1
#include "cuda_emu.hpp"

#include "kernels_warp.cuh"
#include "kernels_cta.cuh"

#include <vector>

using namespace lagassign;

namespace {
LagOffsets g_offs{};   // raw offsets for the NEXT emu_assign_onchip call (fused computePartitionLag), then cleared
template <int EPT>
void run_cta_class(const DevBatch& b, const std::vector<int32_t>& list) {
    if (list.empty()) return;
    const int grid = (int)std::min<size_t>(list.size(), 3);     // grid-stride loop over topics, like the product
    emu::launch(grid, CTA_THREADS, [&] { k_cta_topics<EPT>(b, list.data(), (int)list.size()); });
}
}  // namespace

// Topics with 1..4096 partitions and a non-empty consumer set are computed; the others are left
// untouched (the caller pre-fills the outputs). Returns the number of topics it ran, or -1 for a
// topic the on-chip kernels do not cover (P > C > 8192 cannot occur for P <= 4096).
extern "C" void emu_onchip_set_offsets(const int64_t* begin, const int64_t* end, const int64_t* committed,
                                       const uint8_t* has, int reset_latest) {
    g_offs = LagOffsets{begin, end, committed, has, reset_latest};
}

extern "C" int emu_assign_onchip(int64_t n_topics, const int64_t* topic_off, const int64_t* lag, const int32_t* partition,
                                 const int32_t* topic_cset, const int64_t* cset_off, const int32_t* cset_rank,
                                 const int64_t* tc_off, int32_t* out_partition, int32_t* out_consumer_rank,
                                 int64_t* out_total, int32_t* out_count) {
    DevBatch b{};
    b.topic_off = topic_off; b.lag = lag; b.part = partition; b.topic_cset = topic_cset;
    b.cset_off = cset_off; b.cset_rank = cset_rank; b.tc_off = tc_off;
    b.out_part = out_partition; b.out_cons = out_consumer_rank; b.out_total = out_total; b.out_count = out_count;
    b.offs = g_offs;
    g_offs = LagOffsets{};
    std::vector<int32_t> l_warp, l_cta[N_CTA_CLASSES];
    for (int64_t t = 0; t < n_topics; ++t) {
        const int64_t P = topic_off[t + 1] - topic_off[t];
        const int32_t s = topic_cset[t];
        const int64_t C = cset_off[s + 1] - cset_off[s];
        if (P == 0 || C == 0 || P > CTA_MAX_P) continue;
        if (P <= WARP_MAX_P) l_warp.push_back((int32_t)t);
        else l_cta[cta_class_of(P)].push_back((int32_t)t);
    }
    int ran = 0;
    if (!l_warp.empty()) {
        const int n = (int)l_warp.size();
        const int grid = (n + SMALL_WARPS - 1) / SMALL_WARPS;
        const bool plain = partition == nullptr && out_total == nullptr;     // as launch_warp_topics() chooses
        if (plain && b.offs.end == nullptr)
            emu::launch(grid, SMALL_WARPS * 32, [&] { k_warp_topics<true, LAG_PLAIN>(b, l_warp.data(), n); });
        else if (plain)
            emu::launch(grid, SMALL_WARPS * 32, [&] { k_warp_topics<true, LAG_OFFS>(b, l_warp.data(), n); });
        else
            emu::launch(grid, SMALL_WARPS * 32, [&] { k_warp_topics<false, LAG_EITHER>(b, l_warp.data(), n); });
        ran += n;
    }
    run_cta_class<2>(b, l_cta[0]);
    run_cta_class<4>(b, l_cta[1]);
    run_cta_class<8>(b, l_cta[2]);
    run_cta_class<16>(b, l_cta[3]);
    for (auto& l : l_cta) ran += (int)l.size();
    return ran;
}
