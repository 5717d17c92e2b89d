#!/bin/bash
# round-end reference measurements on one GPU: parity tests, full default bench line (e2e + cpu baseline), reference arm,
# launch list of bench steps, ncu --set full of two sweeps, C1 / C2 / C3 lines, smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 800 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_1gpu.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
tail -c 400 gpurun_out/r2_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b.log 2>&1
tail -n 2 gpurun_out/r2_launches.csv | cut -c1-250
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qfb_sweep -s 21 -c 2 -f -o gpurun_out/prof_jit_final \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_jit_final.log 2>&1
tail -1 gpurun_out/ncu_jit_final.log
for c in c1 c2 c3; do timeout 300 python bench.py --config $c > gpurun_out/r2_bench_$c.json 2>/dev/null; echo "$c rc=$?"; cut -c1-220 gpurun_out/r2_bench_$c.json; done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
