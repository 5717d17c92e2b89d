#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the small-N GPU tests that cover every kernel family
mkdir -p gpurun_out
SEL="wb12_matches_reference or every_tile_size or sweep_specialised or mixed_arity"
for tool in memcheck racecheck; do
  timeout 2400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_circuits.py tests/test_gpu_trajectories.py tests/test_gpu_backend.py -m gpu -q -p no:cacheprovider -x -k "$SEL or trajectories or reductions_and_elementwise or staged_transfer" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_$tool.log | tail -4
done
