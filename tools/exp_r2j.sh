#!/bin/bash
mkdir -p gpurun_out
Q1PHYS_LIB=$PWD/build/libq1phys_trace.so timeout 100 python tools/trace_actor.py loop > gpurun_out/r2h_trace_loop.txt 2>&1
timeout 200 python -m pytest tests/test_sampling_gpu.py -x -q -s > gpurun_out/r2j_sampling.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_sampling.log; tail -12 gpurun_out/r2j_sampling.log
timeout 200 python -m pytest tests/test_actor_gpu.py tests/test_api_gpu.py -x -q > gpurun_out/r2j_actor.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_actor.log; tail -3 gpurun_out/r2j_actor.log
timeout 300 bash tools/ncu_range.sh 1048576 16 2>&1 | tail -3
timeout 300 bash tools/ncu_range.sh 131072 32 2>&1 | tail -3
