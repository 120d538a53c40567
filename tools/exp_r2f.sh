#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_api_gpu.py -x -q -s -k "fused_tcgen05 or shipped_policy" > gpurun_out/r2f_policy.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_policy.log
timeout 120 python -m pytest tests/test_actor_gpu.py -x -q -s > gpurun_out/r2f_actor.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_actor.log
timeout 120 python tools/time_policy.py > gpurun_out/r2f_time_policy.txt 2>&1
Q1PHYS_LIB=$PWD/build/libq1phys_tanhf32.so timeout 120 python tools/time_policy.py > gpurun_out/r2f_time_policy_f32.txt 2>&1
Q1PHYS_LIB=$PWD/build/libq1phys_tanhf32.so timeout 120 python -m pytest tests/test_api_gpu.py -x -q -s -k "fused_tcgen05" 2>&1 | grep "logit error" > gpurun_out/r2f_logit_f32.txt
grep "logit error\|passed\|failed\|rc=" gpurun_out/r2f_policy.log; tail -12 gpurun_out/r2f_actor.log; cat gpurun_out/r2f_time_policy.txt; echo "--- tanh f32 variant"; cat gpurun_out/r2f_time_policy_f32.txt gpurun_out/r2f_logit_f32.txt
