#!/bin/bash
# compute-sanitizer racecheck / synccheck / memcheck over small cases of the kernels that synchronise through mbarriers,
# programmatic dependent launch, dynamic claim counters and cross-warp flags (run under gpurun; logs -> gpurun_out/).
#   bash tools/sanitize.sh
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TESTS="tests/test_gpu_chain.py::test_chain_is_bit_identical_to_hop_by_hop[default-8-5-64-4] tests/test_gpu_chain.py::test_chain_is_bit_identical_to_hop_by_hop[ragged-groups-4-7-128-6] tests/test_gpu_parity.py::test_convcheb_matches_reference_golden[mix-tcgen05-conv_cfg1] tests/test_gpu_parity.py::test_cheb_terms_match_oracle_recurrence tests/test_gpu_chain.py::test_chain_is_bit_identical_to_hop_by_hop[default-8-4-24-4] tests/test_gpu_chain.py::test_chain_is_bit_identical_to_hop_by_hop[default-8-3-96-5]"
for tool in memcheck racecheck synccheck; do
  echo "==== compute-sanitizer --tool $tool ====" | tee gpurun_out/sanitize_$tool.log
  timeout -s KILL 900 compute-sanitizer --tool $tool --target-processes all --print-limit 20 \
      python -m pytest $TESTS -x -q -p no:cacheprovider 2>&1 | grep -v -E "^$|Warning|warnings.warn|torch.sparse|return torch|rename" | tail -40 | tee -a gpurun_out/sanitize_$tool.log
done
