#!/bin/bash
# GPU pass: smoke, parity tests, short bench, ncu launch list + one full capture of the sweep kernel
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
if [ "$1" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 25 -c 3 -f -o gpurun_out/prof_sweep python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
fi
tail -n 12 gpurun_out/smoke.log gpurun_out/pytest.log gpurun_out/bench.log
