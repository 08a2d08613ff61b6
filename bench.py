#!/usr/bin/env python
"""bench.py — partition-assignments/sec of the lag-based assignment hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--impl reference]

A step is one pass of the hot path (sort each topic's partitions by descending
lag, round-fill them onto the topic's consumers) over one synthetic batch.
Default workload: BASELINE.json config 4 (100 000 topics x 256 partitions, 4096
consumers), the configuration the throughput target is quoted on; for N > 1 its
topics are sharded over the ranks and the result all-gathered (strong scaling).

  value     whole-job partition-assignments/s, inputs resident in HBM, timed with
            CUDA events on the launching stream (max over ranks), all-gather
            included when N > 1
  e2e       same metric through lagassign_assign_batch() with pinned HOST buffers
            in and out (H2D + plan + kernels + D2H inside the timed region)
  roofline  dominant kernel of the step vs the measured HBM peak
  cpu_baseline  the reference algorithm (oracle, C restatement: no JVM in this
            image) on one host core, on a bounded sample of the same workload

--impl reference times the CPU restatement of the reference on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import args_of, baseline_config  # noqa: E402

METRIC = "partition-assignments/sec"
L2_BYTES = 126 * 1024 * 1024
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md


def algorithmic_bytes(batch) -> int:
    """per partition 8 B lag in + 4 B partition + 4 B consumer out
    (+ 4 B explicit id), per topic 12 B of offsets / set id."""
    N = int(batch["topic_off"][-1])
    T = len(batch["topic_off"]) - 1
    per = 16 + (4 if batch["partition"] is not None else 0)
    return N * per + T * 12


def workload_config(batch, desc, cfg, world):
    """The `config` object of the JSON line: the same for both arms (a function of the workload and N only)."""
    from kafka_lag_based_assignor_b200.distributed import shard_batch, shard_bounds
    bounds = shard_bounds(batch["topic_off"], world)
    local_bytes = max(algorithmic_bytes(shard_batch(batch, lo, hi)) for lo, hi in bounds)
    flush_needed = local_bytes < 2 * L2_BYTES
    return {"workload": desc, "config": cfg,
            "parallelism": f"topic-sharded x{world}" if world > 1 else "single GPU",
            "l2": "L2 flushed between timed steps (256 MiB write)" if flush_needed else "inputs larger than L2 (126 MB)",
            "collective": ("narrow result form (1-4 B partition id, 0-4 B consumer per partition) all-gathered with NCCL and "
                           "widened, inside the timed step") if world > 1 else "none",
            "partitions_per_step": int(batch["topic_off"][-1])}, flush_needed


# --impl reference: topics per step, fixed per config (the same work on every box)
REFERENCE_TOPICS_PER_STEP = {2: 1, 3: 10000, 4: 4096, 5: 64, 6: 1, 7: 20000}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(batch, mode, threads, budget_s):
    """Time the oracle on a bounded prefix of topics; returns (assignments/s, sample text)."""
    import oracle
    off = batch["topic_off"]
    T = len(off) - 1
    # calibrate on a small prefix, then size the sample for ~budget_s
    probe = max(1, min(T, 64 * max(1, threads)))
    t0 = time.perf_counter()
    oracle.assign_batch(*args_of(batch), mode=mode, threads=threads, topic_lo=0, topic_hi=probe)
    dt = max(time.perf_counter() - t0, 1e-6)
    n_probe = int(off[probe])
    want = int(min(T, max(probe, probe * budget_s / dt)))
    t0 = time.perf_counter()
    oracle.assign_batch(*args_of(batch), mode=mode, threads=threads, topic_lo=0, topic_hi=want)
    dt = time.perf_counter() - t0
    n = int(off[want])
    if n == 0:
        n, dt = n_probe, dt
    return n / dt, f"first {want} of {T} topics ({n} partitions) in {dt:.2f} s"


def run_reference(args, batch, desc):
    """--impl reference: the CPU restatement of LagBasedPartitionAssignor.assign on all host threads."""
    import oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle.build()
    # all host threads (torchrun exports OMP_NUM_THREADS=1, so do not ask OpenMP)
    threads = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    off = batch["topic_off"]
    T = len(off) - 1
    # one step = a FIXED prefix of the workload's topics (the same work on every box)
    per_step = int(min(T, REFERENCE_TOPICS_PER_STEP[args.config]))
    n = int(off[per_step])
    for _ in range(args.warmup):
        oracle.assign_batch(*args_of(batch), mode=oracle.LITERAL, threads=threads, topic_lo=0, topic_hi=per_step)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.assign_batch(*args_of(batch), mode=oracle.LITERAL, threads=threads, topic_lo=0, topic_hi=per_step)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"each step: first {per_step} of {T} topics ({n} partitions)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "assignments/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": workload_config(batch, desc, args.config, args.gpus)[0],
        "reference_note": "reference algorithm (Main:166-277, O(P log P + P*C)) restated in C (oracle/oracle.c, literal "
                          "mode), topics spread over all host threads; the JVM original cannot run in this image",
        "cpu_baseline": {"value": value, "unit": "assignments/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "assignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5, 6, 7])
    ap.add_argument("--scale", type=float, default=1.0, help="scale the workload down (debugging only)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hot-path-only", action="store_true",
                    help="skip the e2e calls and the next-rows section: only execute() steps launch kernels (for ncu launch lists)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    batch, desc = baseline_config(args.config, args.scale)
    if args.impl == "reference":
        run_reference(args, batch, desc)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__
    __graft_entry__.build()
    import kafka_lag_based_assignor_b200 as pkg
    from kafka_lag_based_assignor_b200.distributed import ShardedAssignor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N_total = int(batch["topic_off"][-1])
    stream = torch.cuda.current_stream()
    f = pkg.FlatAssignor(local_rank, stream=stream.cuda_stream)
    sh = ShardedAssignor(f, rank, world)
    sh.upload(batch)
    local_alg_bytes = algorithmic_bytes(sh.local)
    config_obj, flush_needed = workload_config(batch, desc, args.config, world)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda") if flush_needed else None

    def flush():
        if flush_buf is not None:
            flush_buf.fill_(1)

    # ---- warm-up ------------------------------------------------------------------------
    for _ in range(args.warmup):
        flush()
        sh.execute()
    barrier()

    # ---- timed: K steps as a caller runs them, device events around each step on the launching stream ---------
    # (N > 1: kernels, narrowing, all-gather and widening: distributed.py)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b_ in ev:
        flush()
        a.record(stream)
        sh.execute()
        b_.record(stream)
        torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    t_total_ms = sum(a.elapsed_time(b_) for a, b_ in ev)
    t_unprofiled_ms = t_total_ms / args.steps
    # ---- the same steps with CUDA events around every kernel launch (LAGASSIGN_OPT_PROFILE): the per-kernel times
    # behind `roofline` and `kernel_ms_per_step`. Not the headline: the events serialise the size bins of a mixed
    # batch, which otherwise run on concurrent streams.
    f.set_option(pkg.OPT_PROFILE, 1)
    n_prof = max(3, min(args.steps, 10))
    prof_acc = {}
    t_profiled_ms = 0.0
    for _ in range(n_prof):
        flush()
        a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True)
        a.record(stream)
        sh.execute()
        b_.record(stream)
        torch.cuda.synchronize()
        t_profiled_ms += a.elapsed_time(b_) / n_prof
        for k, (ms, n) in f.profile().items():
            acc = prof_acc.setdefault(k, [0.0, 0])
            acc[0] += ms; acc[1] += n
    f.set_option(pkg.OPT_PROFILE, 0)
    exchange_ms = None
    if world > 1:          # where the exchange's time goes (eager steps with extra events)
        acc = {}
        for _ in range(5):
            flush()
            evs = []
            sh.execute(events=evs)
            torch.cuda.synchronize()
            for name, e0, e1 in evs:
                acc[name] = acc.get(name, 0.0) + e0.elapsed_time(e1) / 5
        exchange_ms = dict(acc, chunks=sh.n_chunks,
                           note="per eager step, summed over the chunks; all_gather_wait = time the compute stream waits for "
                                "a chunk's collective after its own work is done (the rest of the collective is hidden behind "
                                "the next chunk's kernels)")
    kernel_ms_step = sum(v[0] for v in prof_acc.values()) / n_prof
    t_compute_ms = kernel_ms_step * args.steps
    launches_per_step = f.stats()["kernel_launches"]
    stats = f.stats()

    # ---- parity spot check (every rank checks a sample of its own shard) --------------------
    import oracle
    oracle.build()
    full_p, full_c = sh.results()
    n_chk = min(len(sh.local["topic_off"]) - 1, 200)
    exp = oracle.assign_batch(*args_of(sh.local), mode=oracle.ROUNDS, topic_lo=0, topic_hi=n_chk)
    n_el = int(sh.local["topic_off"][n_chk])
    got_p = full_p[sh.start:sh.start + n_el].cpu().numpy()
    got_c = full_c[sh.start:sh.start + n_el].cpu().numpy()
    parity_ok = bool(np.array_equal(got_p, exp["out_partition"][:n_el]) and
                     np.array_equal(got_c, exp["out_consumer_rank"][:n_el]))
    if world > 1 and rank == 0:
        # the gathered map must hold every rank's shard: check the last rank's first topics too
        lo_last = sh.bounds[-1][0]
        hi_last = min(lo_last + 50, sh.bounds[-1][1])
        if hi_last > lo_last:
            e2 = oracle.assign_batch(*args_of(batch), mode=oracle.ROUNDS, topic_lo=lo_last, topic_hi=hi_last)
            s0, s1 = int(batch["topic_off"][lo_last]), int(batch["topic_off"][hi_last])
            parity_ok = parity_ok and bool(np.array_equal(full_p[s0:s1].cpu().numpy(), e2["out_partition"][s0:s1]))

    # ---- e2e: the public C-ABI call with pinned host buffers --------------------------------
    loc = sh.local
    f2 = pkg.FlatAssignor(local_rank)
    h_lag = pkg.pinned_empty(len(loc["lag"]), np.int64); h_lag[:] = loc["lag"]
    h_part = None
    if loc["partition"] is not None:
        h_part = pkg.pinned_empty(len(loc["partition"]), np.int32); h_part[:] = loc["partition"]
    n_loc = len(loc["lag"])
    h_op = pkg.pinned_empty(n_loc, np.int32)
    h_oc = pkg.pinned_empty(n_loc, np.int32)
    h_npart = pkg.pinned_empty(4 * n_loc, np.uint8)      # narrow form: at most 4 bytes per entry
    h_ncons = pkg.pinned_empty(4 * n_loc, np.uint8)
    a_in = (loc["topic_off"], h_lag, h_part, loc["topic_cset"], loc["cset_off"], loc["cset_rank"])
    t_e2e_narrow = float("inf")
    narrow_fmt = None
    if args.hot_path_only:
        # ncu launch lists: nothing but the execute() steps above may launch kernels
        e2e_steps = 1
        t_e2e = t_e2e_replan = float("inf")
        st2 = st2_replan = st2_narrow = {"h2d_bytes": 0, "d2h_bytes": 0}
    else:
        e2e_steps = max(3, min(args.steps, 10))

        def timed_calls(call):
            for _ in range(2):
                call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                call()
            torch.cuda.synchronize()
            return time.perf_counter() - t0
        # (a) the int32 result arrays of lagassign_result
        t_e2e = timed_calls(lambda: f2.assign_batch(*a_in, h_op, h_oc))
        st2 = f2.stats()
        parity_ok = parity_ok and bool(np.array_equal(h_op[:n_el], exp["out_partition"][:n_el])) and \
            bool(np.array_equal(h_oc[:n_el], exp["out_consumer_rank"][:n_el]))
        # (b) the narrow result form (lagassign_assign_batch_narrow): same information, fewer bytes device -> host
        got_n = {}
        t_e2e_narrow = timed_calls(lambda: got_n.update(f2.assign_batch_narrow(*a_in, h_npart, h_ncons)))
        st2_narrow = f2.stats()
        narrow_fmt = (got_n["part_width"], got_n["cons_width"])
        parity_ok = parity_ok and bool(np.array_equal(got_n["out_partition"][:n_el].astype(np.int64), exp["out_partition"][:n_el]))
        if got_n["cons_width"]:
            c_n = got_n["out_consumer_rank"][:n_el].astype(np.int64)
            c_n[c_n == (1 << (8 * got_n["cons_width"])) - 1] = -1
            parity_ok = parity_ok and bool(np.array_equal(c_n, exp["out_consumer_rank"][:n_el]))
        # (c) the int32 call with the plan cache off: every call validates, plans and uploads the metadata again
        f2.set_option(pkg.OPT_PLAN_CACHE, 0)
        t_e2e_replan = timed_calls(lambda: f2.assign_batch(*a_in, h_op, h_oc))
        st2_replan = f2.stats()

    # ---- the two "next" rows of SURVEY 8(f) that are built, measured beside the hot path (N = 1) ----
    next_rows = None
    if world == 1 and not args.hot_path_only:
        def timed(fa, steps=5):
            for _ in range(2):
                flush(); fa.execute()
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(steps):
                flush()
                a = torch.cuda.Event(enable_timing=True); z = torch.cuda.Event(enable_timing=True)
                a.record(stream); fa.execute(); z.record(stream); torch.cuda.synchronize()
                tot += a.elapsed_time(z)
            return tot / steps
        ms_plain = timed(f)
        fg = pkg.FlatAssignor(local_rank, stream=stream.cuda_stream)
        fg.set_option(pkg.OPT_GROUP, 1)
        fg.upload(*args_of(batch))
        ms_group = timed(fg)
        g = fg.download_grouping()
        grouped_ok = bool(g["member_off"][-1] == fg.stats()["n_assigned"])
        fg.close()
        fo = pkg.FlatAssignor(local_rank, stream=stream.cuda_stream)
        fo.set_option(pkg.OPT_PROFILE, 1)
        lagv = np.maximum(batch["lag"], 0)
        zeros = np.zeros(len(lagv), np.int64)
        fo.upload_offsets(batch["topic_off"], zeros, lagv, zeros, np.ones(len(lagv), np.uint8), 1, batch["partition"],
                          batch["topic_cset"], batch["cset_off"], batch["cset_rank"])
        ms_off = timed(fo)
        fo.close()
        next_rows = {"grouping_by_member": {"ms_per_step_with_grouping": ms_group, "ms_per_step_without": ms_plain,
                                            "check": "ok" if grouped_ok else "MISMATCH",
                                            "note": "SURVEY 8f row 2: stable partition of the flat result by member rank on the device"},
                     "device_lag_computation": {"ms_per_step_with_offsets": ms_off,
                                                "ratio_to_plain_step": ms_off / ms_plain,
                                                "note": "SURVEY 8f row 1: computePartitionLag (Main:376-404) evaluated inside the sort kernels as "
                                                        "they build their keys (no lag array, no extra launch): the step reads 25 B/partition of "
                                                        "raw offsets instead of 8 B of lags"}}

    # ---- reduce over ranks --------------------------------------------------------------------
    times = torch.tensor([t_total_ms, t_compute_ms, t_e2e * 1e3, 0.0 if parity_ok else 1.0, t_e2e_replan * 1e3, t_e2e_narrow * 1e3],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_total_ms, t_compute_ms, t_e2e_ms, bad, t_e2e_replan_ms, t_e2e_narrow_ms = [float(x) for x in times.tolist()]

    if rank == 0:
        peak, peak_src = hbm_peak()
        value = N_total * args.steps / (t_total_ms * 1e-3)
        # dominant kernel of this rank's step
        dom = max(prof_acc.items(), key=lambda kv: kv[1][0]) if prof_acc else ("none", [1e-9, 1])
        dom_name, (dom_ms, dom_n) = dom[0], dom[1]
        step_ms = kernel_ms_step
        share = dom_ms / max(sum(v[0] for v in prof_acc.values()), 1e-12)
        if dom_name == "k_big_scatter":
            # 64-bit key + 32-bit payload read and written once per executed digit pass
            alg = 24 * stats["device_sort_elems"]
            alg_note = "24 B x elements per scatter launch (one launch per executed digit pass)"
        elif dom_name == "k_big_tile_hist":
            alg = 8 * stats["device_sort_elems"]
            alg_note = "8 B x elements per per-tile histogram launch"
        elif dom_name in ("k_warp_topics", "k_cta_topics"):
            alg = local_alg_bytes
            alg_note = "16 B/partition (+4 B explicit ids) + 12 B/topic, whole step in one launch"
        else:
            alg = None
            alg_note = "latency-bound kernel (sequential round fill): no bandwidth roofline, see time share"
        per_launch_ms = dom_ms / max(dom_n, 1)
        # `alg` is what ONE launch of the radix kernels moves, but what ALL launches of a step move together for the
        # on-chip kernels (N > 1 runs them chunk by chunk)
        alg_ms = dom_ms / n_prof if dom_name in ("k_warp_topics", "k_cta_topics") else per_launch_ms
        achieved = (alg / (alg_ms * 1e-3) / 1e9) if alg else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if world == 1 and os.path.exists(tp):     # measured at N = 1 (ncu is never run on a multi-rank command)
            try:
                traffic = json.load(open(tp)).get(f"config{args.config}", {}).get(dom_name)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "assignments/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": config_obj,
            "clocks": clocks,
            "e2e": {"value": N_total * e2e_steps / (t_e2e_narrow_ms * 1e-3), "unit": "assignments/s",
                    "h2d_bytes_per_step": st2_narrow["h2d_bytes"], "d2h_bytes_per_step": st2_narrow["d2h_bytes"],
                    "steps": e2e_steps,
                    "api": "lagassign_assign_batch_narrow (C ABI), pinned host buffers, per-rank shard; result form: "
                           f"partition ids in {narrow_fmt[0] if narrow_fmt else '?'} byte(s), consumers in "
                           f"{narrow_fmt[1] if narrow_fmt else '?'} byte(s) (0 = every topic has P <= C: the consumer of sorted "
                           "partition k is the k-th member of the topic's set, which the caller holds; include/lagassign.h)",
                    "note": "batch shape repeats between calls (as between rebalances), so the plan and the metadata on the "
                            "device are reused and only lags go up. value_int32 = the same call returning the two int32 arrays "
                            "of lagassign_result (lagassign_assign_batch); value_int32_replanned = that call with the plan cache "
                            "off (every call validates, plans and uploads the metadata again)",
                    "value_int32": N_total * e2e_steps / (t_e2e_ms * 1e-3),
                    "d2h_bytes_per_step_int32": st2["d2h_bytes"],
                    "value_int32_replanned": N_total * e2e_steps / (t_e2e_replan_ms * 1e-3),
                    "h2d_bytes_per_step_replanned": st2_replan["h2d_bytes"]},
            "gpu_launches": (launches_per_step + (2 * sh.n_chunks if world > 1 else 0)) * args.steps,   # + narrow / widen kernels
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "bytes_model": alg_note,
                         "launch_ms": per_launch_ms, "share_of_step": share,
                         "note": ("launch-bound step: a kernel of %.0f us is about the cost of launching and draining it on 148 "
                                  "SMs, so the HBM fraction says little here" % (per_launch_ms * 1e3))
                                 if (alg and per_launch_ms < 0.03) else None},
            "kernel_ms_per_step": {k: v[0] / n_prof for k, v in prof_acc.items()},
            "ms_per_step_unprofiled": t_unprofiled_ms,
            "ms_per_step_profiled": t_profiled_ms,
            "compute_ms_per_step": step_ms,
            "gather_ms_per_step": max(t_total_ms / args.steps - step_ms, 0.0),
            "exchange_ms": exchange_ms,
            "parity_spot_check": "ok" if bad == 0.0 else "MISMATCH",
            "next_rows": next_rows,
            "stats": stats,
        }
        if not args.no_cpu_baseline and args.config == 6:
            # sort-only variant (P = C = 2^20 per topic): the literal O(P*C) scan would need ~1e12
            # comparisons for ONE topic, so only the round formulation is timed, on one topic
            v2, sample2 = cpu_baseline(dict(batch, topic_off=batch["topic_off"][:2], topic_cset=batch["topic_cset"][:1],
                                            lag=batch["lag"][:int(batch["topic_off"][1])]), oracle.ROUNDS, 1, 1.0)
            line["cpu_baseline"] = {"value": v2, "unit": "assignments/s", "cores": 1, "kind": "port", "sample": sample2,
                                    "note": "round formulation only (the literal scan is O(P*C) = 1e12 per topic here)"}
        elif not args.no_cpu_baseline:
            v, sample = cpu_baseline(batch, oracle.LITERAL, 1, args.cpu_budget)
            v2, sample2 = cpu_baseline(batch, oracle.ROUNDS, 1, min(args.cpu_budget, 4.0))
            line["cpu_baseline"] = {"value": v, "unit": "assignments/s", "cores": 1, "kind": "port", "sample": sample,
                                    "note": "reference algorithm (O(P*C) scan per partition, Main:240-263) restated in C; "
                                            "single thread like the reference (Main:177-184)",
                                    "rounds_variant": {"value": v2, "sample": sample2,
                                                       "note": "same result with the O(P log P + P log C) round formulation"}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if bad != 0.0:
        raise SystemExit("parity spot check failed")


if __name__ == "__main__":
    main()
