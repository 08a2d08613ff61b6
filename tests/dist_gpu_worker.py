"""Worker of tests/test_gpu_parity.py::test_sharded_nccl_*: one process per GPU (torchrun), the production
ShardedAssignor path (kernels -> narrow -> NCCL all-gather -> widen); EVERY rank compares BOTH gathered arrays
of the WHOLE batch with the oracle. Exits non-zero on any mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import oracle
    import kafka_lag_based_assignor_b200 as pkg
    from kafka_lag_based_assignor_b200.distributed import ShardedAssignor, exchange_format
    from helpers import args_of, baseline_config, random_flat

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(2024)          # the same batches on every rank
    cases = []
    b4, _ = baseline_config(4, scale=0.03)                                  # P < C everywhere: (1, 0)
    cases.append(("config4 slice", b4))
    cases.append(("ragged mixed bins, shuffled ids", random_flat(rng, T=300, pmin=0, pmax=3000, C=400, lag_kind="u40",
                                                                 explicit_ids="shuffled", n_sets=4)))
    cases.append(("many rounds, small sets", random_flat(rng, T=700, pmin=1, pmax=256, C=12, lag_kind="small", n_sets=3)))
    big = random_flat(rng, T=5, pmin=70000, pmax=90000, C=100, lag_kind="full", n_sets=1, full_sets=True)   # ids need 4 bytes
    cases.append(("large topics", big))
    bad = 0
    for name, b in cases:
        f = pkg.FlatAssignor(local, stream=torch.cuda.current_stream().cuda_stream)
        sh = ShardedAssignor(f, rank, world, n_chunks=3 if name.startswith("config4") or name.startswith("many rounds") else 0)
        sh.upload(b)
        sh.execute()
        sh.execute()          # repeatable
        torch.cuda.synchronize()
        fp, fc = sh.results()
        exp = oracle.assign_batch(*args_of(b), threads=8)
        ok = np.array_equal(fp.cpu().numpy(), exp["out_partition"]) and np.array_equal(fc.cpu().numpy(), exp["out_consumer_rank"])
        print(f"rank {rank}/{world}: {name}: format {exchange_format(b)} -> {'ok' if ok else 'MISMATCH'}", flush=True)
        bad += not ok
        f.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
