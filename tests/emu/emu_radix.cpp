// emu_radix.cpp — TEST INFRASTRUCTURE: the device-wide path of the product (radix.cuh: k_big_order_check,
// k_big_range, k_big_scan, k_big_tile_hist, k_big_seg_scan, k_big_topic_scan, k_big_scatter, k_big_finalize,
// k_big_fill) compiled for the HOST with the stand-in of cuda_emu.hpp and run one CTA at a time, in the
// launch order of lagassign_execute(). The planning below restates the big-topic part of validate_and_plan()
// (lagassign_api.cu) and is test code: it may disagree with the product, which is the point of checking.
// Built and driven by tests/test_kernels_emulated_cpu.py; never part of the library.
#define LAGASSIGN_HOST_EMULATION 1
#include "cuda_emu.hpp"

#include "radix.cuh"

#include <algorithm>
#include <vector>

using namespace lagassign;

namespace {
LagOffsets g_offs{};   // raw offsets for the NEXT emu_assign_big call (k_big_hist then computes and stores the lags)
template <class K> void launch(int grid, int block, K k) { if (grid > 0) emu::launch(grid, block, k); }
}

// Every topic with more than `min_big_p` partitions and a non-empty consumer set goes through the radix
// path (the product uses it above 4096; smaller thresholds exercise short last tiles and one-tile topics).
// `sms` scales the grids like the SM count does in the product (small values force grid-stride loops).
// Returns the number of topics processed, -1 for an unsupported shape.
extern "C" void emu_big_set_offsets(const int64_t* begin, const int64_t* end, const int64_t* committed,
                                    const uint8_t* has, int reset_latest) {
    g_offs = LagOffsets{begin, end, committed, has, reset_latest};
}

extern "C" int emu_assign_big(int64_t n_topics, const int64_t* topic_off, const int64_t* lag, const int32_t* partition,
                              const int32_t* topic_cset, const int64_t* cset_off, const int32_t* cset_rank,
                              const int64_t* tc_off, int32_t* out_partition, int32_t* out_consumer_rank,
                              int64_t* out_total, int32_t* out_count, int64_t min_big_p, int sms, int* passes_run_out) {
    const LagOffsets offs = g_offs;
    g_offs = LagOffsets{};
    std::vector<int32_t> bt_topic, bt_tile0, tile_bt, l_fill, bt_seg0, seg_bt;
    std::vector<int64_t> bt_off, bt_P, bt_boff, fill_ws_off;
    int64_t fill_ws_elems = 0;
    std::vector<TileDesc> tile_desc;
    int64_t boff = 0, tiles = 0;
    int n_segs = 0;
    for (int64_t t = 0; t < n_topics; ++t) {
        const int64_t P = topic_off[t + 1] - topic_off[t];
        const int32_t s = topic_cset[t];
        const int64_t C = cset_off[s + 1] - cset_off[s];
        if (P <= min_big_p || C == 0) continue;
        const int64_t nt = (P + RT_TILE - 1) / RT_TILE;
        const int32_t b = (int32_t)bt_topic.size();
        bt_topic.push_back((int32_t)t); bt_off.push_back(topic_off[t]); bt_P.push_back(P); bt_boff.push_back(boff);
        bt_tile0.push_back((int32_t)tiles); bt_seg0.push_back(n_segs);
        const int nseg = (int)((nt + RT_SEG - 1) / RT_SEG);
        for (int k = 0; k < nseg; ++k) seg_bt.push_back(b);
        for (int64_t k = 0; k < nt; ++k) {
            tile_bt.push_back(b);
            TileDesc d;
            d.b = b; d.tile_in_topic = (int32_t)k; d.n_valid = (int32_t)std::min<int64_t>(RT_TILE, P - k * RT_TILE);
            d.seg = (int32_t)(n_segs + k / RT_SEG); d.buf_off = boff + k * RT_TILE; d.raw_off = topic_off[t] + k * RT_TILE;
            tile_desc.push_back(d);
        }
        n_segs += nseg;
        if (P > C) l_fill.push_back(b);
        int64_t ws = -1;
        if (P > C && C > FILL_CTA_MAX_C) { int64_t nc = 1; while (nc < C) nc <<= 1; ws = fill_ws_elems; fill_ws_elems += nc; }
        fill_ws_off.push_back(ws);
        boff += (P + RT_BUF_ALIGN - 1) / RT_BUF_ALIGN * RT_BUF_ALIGN; tiles += nt;
    }
    const int B = (int)bt_topic.size();
    if (B == 0) return 0;
    static_assert(sizeof(TileDesc) == 32, "tile descriptors are read as two int4");

    // (the product's device buffers carry allocation slack; bulk copies round sizes up to 16 bytes)
    std::vector<uint64_t> keyA((size_t)boff + 4), keyB((size_t)boff + 4);
    std::vector<uint32_t> payA((size_t)boff + 4), payB((size_t)boff + 4);
    std::vector<uint8_t> digbuf((size_t)boff + 16, 0xEE);
    // the caller's lag array, in a 16-byte aligned copy with slack at both ends like the product's device copy;
    // `lag_shift` (0 or 1 element) makes the emulated device pointer 16-byte aligned or not
    const int64_t n_total = topic_off[n_topics];
    std::vector<int64_t> lag_store((size_t)n_total + 8, 0);
    const int lag_shift = (int)(n_total & 1);
    int64_t* lag_base = lag_store.data() + 2;
    if ((reinterpret_cast<uintptr_t>(lag_base) & 15u) != 0) ++lag_base;          // 16-byte aligned
    lag_base += lag_shift;                                                       // or deliberately only 8-byte aligned
    if (offs.end == nullptr) std::copy(lag, lag + n_total, lag_base);   // with raw offsets the buffer starts zeroed
    lag = lag_base;
    std::vector<uint64_t> range((size_t)B * 4, 0);
    std::vector<uint32_t> pass_rec((size_t)tiles * N_PASS, 0xEEEEEEEEu);
    std::vector<int8_t> sel((size_t)B * (3 * N_PASS + 2), 0);
    std::vector<int32_t> need_tie((size_t)B, 0);
    std::vector<uint32_t> tile_hist((size_t)tiles * 256, 0), seg_total((size_t)n_segs * 256, 0), ticket(RT_TICKET_WORDS, 0);

    BigTopics bt{};
    bt.n_big = B; bt.n_tiles = (int)tiles;
    bt.topic = bt_topic.data(); bt.off = bt_off.data(); bt.P = bt_P.data(); bt.boff = bt_boff.data();
    bt.tile0 = bt_tile0.data(); bt.tile_bt = tile_bt.data();
    bt.tile_desc = reinterpret_cast<const int4*>(tile_desc.data());
    bt.keyA = keyA.data(); bt.keyB = keyB.data(); bt.payA = payA.data(); bt.payB = payB.data();
    bt.range = range.data(); bt.pass_rec = pass_rec.data();
    bt.sel_src = sel.data(); bt.sel_dst = sel.data() + (size_t)B * N_PASS;
    bt.nxt_pass = sel.data() + (size_t)B * N_PASS * 2;
    bt.sel_final = sel.data() + (size_t)B * N_PASS * 3; bt.all_equal = sel.data() + (size_t)B * (N_PASS * 3 + 1);
    bt.dig = digbuf.data();
    std::vector<uint64_t> ws_key((size_t)fill_ws_elems + 1);
    std::vector<uint32_t> ws_idx((size_t)fill_ws_elems + 1);
    bt.fill_ws_off = fill_ws_off.data(); bt.fill_ws_key = ws_key.data(); bt.fill_ws_idx = ws_idx.data();
    bt.need_tie = need_tie.data(); bt.tile_hist = tile_hist.data();
    bt.seg0 = bt_seg0.data(); bt.seg_bt = seg_bt.data(); bt.n_segs = n_segs; bt.seg_total = seg_total.data();
    bt.ticket = ticket.data(); bt.passes_run = reinterpret_cast<int32_t*>(ticket.data() + N_PASS);
    bt.pass_any = ticket.data() + N_PASS + 1;
    bt.verify_fail = ticket.data() + 2 * N_PASS + 1;

    DevBatch db{};
    db.topic_off = topic_off; db.lag = lag; db.part = partition; db.topic_cset = topic_cset;
    db.cset_off = cset_off; db.cset_rank = cset_rank; db.tc_off = tc_off;
    db.out_part = out_partition; db.out_cons = out_consumer_rank; db.out_total = out_total; db.out_count = out_count;

    // the launch sequence of lagassign_execute() for the device bin (launch_big.cu for the grid sizes)
    const bool has_part = partition != nullptr;
    // The emulator's "lanes" are free-running host threads, so the fast ranking (one atomic per element, relying on
    // lane-ordered service) is NOT stable here: the exact form is the default, and EMU_FAST_RANK=1 is used to check
    // that the certificate (k_big_verify) catches what then goes wrong.
    const bool exact = getenv("EMU_FAST_RANK") == nullptr;
    const int first_pass = has_part ? 0 : N_TIE_PASS;
    const int n_tiles = (int)tiles;
    if (has_part) launch(std::min(n_tiles, sms * 8), 256, [&] { k_big_order_check(bt, partition); });
    if (has_part) launch(std::min(n_tiles, sms * 8), RT_THREADS, [&] { k_big_range<true>(bt, lag_base, offs, partition); });
    else launch(std::min(n_tiles, sms * 8), RT_THREADS, [&] { k_big_range<false>(bt, lag_base, offs, nullptr); });
    launch(B, 256, [&] { k_big_scan(bt, first_pass); });
    for (int p = first_pass; p < N_PASS; ++p) {
        if (bt.pass_any[p] == 0) continue;
        if (has_part) launch(std::min(n_tiles, sms * 8), RT_THREADS, [&] { k_big_tile_hist<true>(bt, p, lag, partition); });
        else launch(std::min(n_tiles, sms * 8), RT_THREADS, [&] { k_big_tile_hist<false>(bt, p, lag, nullptr); });
        launch(n_segs, 256, [&] { k_big_seg_scan(bt, p); });
        launch(B, 256, [&] { k_big_topic_scan(bt, p); });
        if (has_part) launch(std::min(n_tiles, sms * 2), SC_THREADS, [&] { (exact ? k_big_scatter<true, true>(bt, p, lag, partition) : k_big_scatter<true, false>(bt, p, lag, partition)); });
        else launch(std::min(n_tiles, sms * 2), SC_THREADS, [&] { (exact ? k_big_scatter<false, true>(bt, p, lag, nullptr) : k_big_scatter<false, false>(bt, p, lag, nullptr)); });
    }
    // the certificate rides on the finalize kernel (the stand-alone k_big_verify serves the sort-only mode: both are run
    // here and must agree)
    launch(std::min(n_tiles, sms * 8), 256, [&] { k_big_verify(bt, partition); });
    const uint32_t standalone_verdict = *bt.verify_fail;
    *bt.verify_fail = 0;
    launch(std::min(n_tiles, sms * 8), 256, [&] {
        k_big_finalize<true>(bt, lag, partition, topic_cset, cset_off, cset_rank, tc_off, out_partition, out_consumer_rank, out_total, out_count);
    });
    if ((*bt.verify_fail != 0) != (standalone_verdict != 0)) return -3;   // the two forms of the certificate disagree
    if (*bt.verify_fail != 0) return -2;      // the certificate of the sort failed
    if (!l_fill.empty()) {
        const int n_fill = (int)l_fill.size();
        launch(std::min(n_fill, sms), BIGFILL_THREADS, [&] {
            k_big_fill(bt, l_fill.data(), n_fill, lag, topic_cset, cset_off, cset_rank, tc_off, out_consumer_rank, out_total, out_count);
        });
    }
    if (passes_run_out) *passes_run_out = *bt.passes_run;
    return B;
}
