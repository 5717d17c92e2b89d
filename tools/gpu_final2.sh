#!/bin/bash
# round-end reference measurements: GPU parity tests, full default bench line (e2e + cpu baseline), the reference arm,
# launch list of one bench step, C1 / C3 timings, smoke
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -n 3 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -c 2600 gpurun_out/bench_full.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
tail -c 600 gpurun_out/bench_reference.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -n 2 gpurun_out/launches.csv | cut -c1-300
timeout 300 python tools/configs_timing.py > gpurun_out/configs_timing.jsonl 2> gpurun_out/configs_timing.err; cat gpurun_out/configs_timing.jsonl
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
