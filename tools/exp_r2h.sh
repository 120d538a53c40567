#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_api_gpu.py -x -q -s -k "fused_tcgen05 or shipped_policy" > gpurun_out/r2h_policy.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_policy.log
for i in 1 2; do timeout 120 python -m pytest tests/test_actor_gpu.py -x -q -s > gpurun_out/r2h_actor_$i.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_actor_$i.log; tail -4 gpurun_out/r2h_actor_$i.log; done
timeout 120 python tools/time_policy.py > gpurun_out/r2h_time_policy.txt 2>&1
