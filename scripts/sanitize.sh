#!/bin/bash
# compute-sanitizer over a subset of the GPU parity tests (run through gpurun). memcheck: out-of-bounds /
# misaligned accesses; racecheck: shared-memory hazards in the on-chip kernels.
mkdir -p gpurun_out
SEL="test_warp_bin_every_size or test_mixed_bins_and_edge_topics or test_plan_cache or test_single_consumer or test_wraparound or test_grouping_reference_vector or test_fused_lag"
timeout -s KILL 420 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_gpu_parity.py -q -x --timeout=400 -k "$SEL" 2>&1 | tail -3
echo "memcheck rc=$?"; grep -c "Invalid\|misaligned" gpurun_out/memcheck.log; tail -3 gpurun_out/memcheck.log
SEL2="test_mixed_bins_and_edge_topics or test_wraparound"
timeout -s KILL 300 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log \
  python -m pytest tests/test_gpu_parity.py -q -x --timeout=280 -k "$SEL2" 2>&1 | tail -3
echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/racecheck.log; tail -3 gpurun_out/racecheck.log
