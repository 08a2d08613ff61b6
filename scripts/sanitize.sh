#!/bin/bash
# compute-sanitizer over the GPU parity tests (run through gpurun). memcheck: out-of-bounds / misaligned accesses, on
# every test except the full-size BASELINE configs (minutes each under the tool); racecheck: shared-memory hazards in
# the on-chip kernels and the radix pass.
mkdir -p gpurun_out
timeout -s KILL 420 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_full.log \
  python -m pytest tests/test_gpu_parity.py -q -x --timeout=400 -k "not full and not sharded_nccl" 2>&1 | tail -2
echo "memcheck rc=$?"; tail -1 gpurun_out/memcheck_full.log
SEL2="test_mixed_bins_and_edge_topics or test_wraparound or (test_device_bin and u40 and None)"
timeout -s KILL 200 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log \
  python -m pytest tests/test_gpu_parity.py -q -x --timeout=180 -k "$SEL2" 2>&1 | tail -2
echo "racecheck rc=$?"; tail -1 gpurun_out/racecheck.log
