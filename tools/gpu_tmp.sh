#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_circuits.py -m gpu -q --timeout 800 -p no:cacheprovider -x -k "specialised or wb24 or wb28 or round_trip" > gpurun_out/pytest_jit.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_jit.log
bash tools/gpu_bench_matrix.sh "|" "QFB_JIT_DEFER=0|" "QFB_JIT_MINB=4|" "QFB_JIT_MINB=6|" "QFB_JIT_COEF_PIN=2|" "QFB_JIT_BARB=1|"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_circuits.py -m gpu -q -p no:cacheprovider -x -k "sweep_specialised" > gpurun_out/r2_sanitizer_racecheck_jit.log 2>&1
echo "== racecheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_racecheck_jit.log | tail -4
