#!/bin/bash
# full evidence run: gpu tests, bench (both arms), ncu launch list + full capture of the step kernel
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_tma -s 30 -c 2 -o gpurun_out/prof_step_tma -f python bench.py --steps 40 --warmup 10 --no-cpu-baseline > /dev/null 2>&1
tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_reference.json | cut -c1-200; cat gpurun_out/bench_b200.json
