"""Host-side multi-GPU logic on CPU: topic sharding and the result all-gather,
world_size 2 and 3 over gloo. The per-rank compute is stood in for by the oracle
(the kernels need a GPU); what is under test is kafka_lag_based_assignor_b200
.distributed — shard bounds, sub-batch slicing, equal and ragged gathers."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from helpers import args_of, random_flat, uniform_batch
        from kafka_lag_based_assignor_b200 import distributed as D
        rng = np.random.default_rng(123)     # same batch on every rank
        if kind == "equal":
            batch = uniform_batch(12, 20, 7, rng.integers(0, 1000, size=240, dtype=np.int64))
        else:
            batch = random_flat(rng, T=23, pmin=0, pmax=90, C=9, lag_kind="dups", explicit_ids="shuffled", n_sets=3)
        bounds = D.shard_bounds(batch["topic_off"], world)
        counts = D.partition_counts(batch["topic_off"], bounds)
        assert sum(counts) == int(batch["topic_off"][-1])
        assert bounds[0][0] == 0 and bounds[-1][1] == len(batch["topic_off"]) - 1
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
        lo, hi = bounds[rank]
        local = D.shard_batch(batch, lo, hi)
        loc = oracle.assign_batch(*args_of(local))          # stand-in for the GPU kernels
        fp, fc = D.all_gather_results(torch.from_numpy(loc["out_partition"]),
                                      torch.from_numpy(loc["out_consumer_rank"]), counts)
        full = oracle.assign_batch(*args_of(batch))
        ok = np.array_equal(fp.numpy(), full["out_partition"]) and np.array_equal(fc.numpy(), full["out_consumer_rank"])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("kind", ["equal", "ragged"])
def test_shard_and_gather_gloo(world, kind):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, kind, ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}


def test_shard_bounds_properties():
    from kafka_lag_based_assignor_b200 import distributed as D
    rng = np.random.default_rng(0)
    for _ in range(50):
        T = int(rng.integers(0, 40))
        P = rng.integers(0, 1000, size=T)
        off = np.concatenate([[0], np.cumsum(P)]).astype(np.int64)
        for world in (1, 2, 4, 8):
            b = D.shard_bounds(off, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == T
            assert all(lo <= hi for lo, hi in b)
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    # equal topics split equally (BASELINE config 4 shape)
    off = np.arange(0, 100001, dtype=np.int64) * 256
    for world in (2, 4, 8):
        assert D.partition_counts(off, D.shard_bounds(off, world)) == [25600000 // world] * world
    # one giant topic is never split
    off = np.array([0, 10, 1 << 20, (1 << 20) + 5], dtype=np.int64)
    b = D.shard_bounds(off, 4)
    assert sum(hi - lo for lo, hi in b) == 3


def test_exchange_format_choices():
    """Width of the narrow result form (include/lagassign.h "Result exchange"), decided from the whole batch."""
    from helpers import baseline_config, uniform_batch
    from kafka_lag_based_assignor_b200 import distributed as D
    b4, _ = baseline_config(4, scale=0.001)
    assert D.exchange_format(b4) == (1, 0)                    # 256 partitions, P < C: ids in a byte, consumers derived
    b3, _ = baseline_config(3, scale=0.01)
    assert D.exchange_format(b3) == (1, 0)
    b7, _ = baseline_config(7, scale=0.01)
    assert D.exchange_format(b7) == (1, 2)                    # P > C: consumers travel, 16 members fit 2 bytes
    assert D.exchange_format(uniform_batch(2, 257, 300, np.zeros(514, np.int64))) == (2, 0)
    assert D.exchange_format(uniform_batch(1, 70000, 10, np.zeros(70000, np.int64))) == (4, 2)
    wide = uniform_batch(1, 10, 3, np.zeros(10, np.int64))
    wide["cset_rank"] = np.array([5, 70000, 70001], np.int32)
    assert D.exchange_format(wide) == (1, 4)
    neg = uniform_batch(1, 4, 8, np.zeros(4, np.int64))
    neg["partition"] = np.array([-1, 0, 1, 2], np.int32)
    assert D.exchange_format(neg) == (4, 0)
    unsub = uniform_batch(2, 5, 0, np.zeros(10, np.int64))    # nobody subscribed: nothing to send for consumers
    assert D.exchange_format(unsub) == (1, 0)


def test_chunk_ranges_cover_every_shard_once():
    """The exchange pipeline cuts every rank's topic range into the same number of contiguous chunks."""
    from kafka_lag_based_assignor_b200 import distributed as D
    rng = np.random.default_rng(4)
    for _ in range(30):
        T = int(rng.integers(0, 60))
        P = rng.integers(0, 500, size=T)
        off = np.concatenate([[0], np.cumsum(P)]).astype(np.int64)
        for world in (1, 2, 3, 8):
            bounds = D.shard_bounds(off, world)
            for k in (1, 2, 4):
                chunks = D.chunk_ranges(bounds, k)
                assert len(chunks) == k and all(len(c) == world for c in chunks)
                for r, (lo, hi) in enumerate(bounds):
                    pieces = [chunks[c][r] for c in range(k)]
                    assert pieces[0][0] == lo and pieces[-1][1] == hi
                    assert all(pieces[i][1] == pieces[i + 1][0] for i in range(k - 1))
                    assert all(a <= b for a, b in pieces)
    assert D.pick_chunks(12_800_000, False) == 1 and D.pick_chunks(50_000_000, True) == 1      # see pick_chunks: measured
