#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2d_smoke.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest.log
python tools/time_small.py > gpurun_out/r2d_small.txt 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2d_bench_reference.json 2> gpurun_out/r2d_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_20.json 2> gpurun_out/r2d_bench_20.err
python bench.py --no-cpu-baseline > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err
bash tools/ncu_range.sh 1048576 16
bash tools/ncu_range.sh 131072 32
tail -2 gpurun_out/r2d_smoke.log; tail -3 gpurun_out/r2d_pytest.log; cat gpurun_out/r2d_small.txt; tail -c 600 gpurun_out/r2d_bench_20.err
