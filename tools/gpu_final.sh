#!/bin/bash
# the round's reference measurements: full default bench line (e2e + cpu baseline), the reference arm, launch list
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_full.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
tail -c 800 gpurun_out/bench_reference.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -n 3 gpurun_out/launches.csv | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
