"""PCIe probe: what the host<->device legs of assign_batch can reach on this box (pinned memory)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import kafka_lag_based_assignor_b200 as pkg
from helpers import baseline_config, args_of

def ms(fn, n=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3

nbytes = 25_600_000 * 8
h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def h2d(): d_in.copy_(h_in, non_blocking=True)
def d2h(): h_out.copy_(d_out, non_blocking=True)
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    t = ms(fn)
    print(f"{name}: {t:.3f} ms  {nbytes / t / 1e6:.1f} GB/s per direction")

b, desc = baseline_config(4)
a = args_of(b)
N = len(b["lag"])
with pkg.FlatAssignor(0) as f:
    lag = pkg.pinned_empty(N, np.int64); lag[:] = b["lag"]
    op = pkg.pinned_empty(N, np.int32); oc = pkg.pinned_empty(N, np.int32)
    b2 = dict(b); b2["lag"] = lag
    a2 = args_of(b2)
    def e2e(): f.assign_batch(*a2, out_partition=op, out_consumer_rank=oc)
    try:
        t = ms(e2e)
        print(f"assign_batch e2e: {t:.3f} ms -> {N / t / 1e6:.2f} G/s")
    except TypeError as e:
        print("assign_batch signature:", e)
    for ch in (1 << 20, 2 << 20, 3 << 20, 6 << 20, 13 << 20, 26 << 20):
        f.set_option(pkg.OPT_PIPE_CHUNK, ch)
        t = ms(e2e)
        print(f"  chunk {ch >> 20}M: {t:.3f} ms -> {N / t / 1e6:.2f} G/s")
    f.set_option(pkg.OPT_PIPE_CHUNK, 0)
    os.environ["LAGASSIGN_TRACE"] = "1"
    with pkg.FlatAssignor(0) as g:
        for _ in range(3): g.assign_batch(*a2, out_partition=op, out_consumer_rank=oc)
    def up(): f.upload(*a2)
    print(f"upload: {ms(up):.3f} ms")
    def ex(): f.execute(); f.sync()
    print(f"execute+sync: {ms(ex):.3f} ms")
    def dn(): f.download(out_partition=op, out_consumer_rank=oc)
    try: print(f"download: {ms(dn):.3f} ms")
    except TypeError as e: print("download signature:", e)

# warp kernel: aligned (16-byte stores from registers) vs unaligned (staged) topics, and a ragged mix
from helpers import uniform_batch, random_flat
rng = np.random.default_rng(7)
for (T, P, C) in ((100000, 256, 4096), (100000, 255, 4096), (100000, 252, 4096), (100000, 200, 4096), (200000, 100, 4096)):
    bb = uniform_batch(T, P, C, rng.integers(0, 1 << 40, size=T * P, dtype=np.int64))
    with pkg.FlatAssignor(0) as f:
        f.set_option(pkg.OPT_PROFILE, 1)
        f.upload(*args_of(bb))
        for _ in range(5): f.execute()
        f.sync()
        prof = f.profile()
    t = prof["k_warp_topics"][0] / prof["k_warp_topics"][1]
    print(f"T={T} P={P}: k_warp_topics {t:.4f} ms -> {T * P / t / 1e6:.1f} G/s, {T * P * 16 / t / 1e6:.0f} GB/s")
