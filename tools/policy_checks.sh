#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_api_gpu.py -x -q -s -k "fused_tcgen05 or shipped_policy" > gpurun_out/policy_policy.log 2>&1; echo "rc=$?" >> gpurun_out/policy_policy.log
timeout 120 python -m pytest tests/test_actor_gpu.py -x -q -s > gpurun_out/policy_actor_1.log 2>&1; echo "rc=$?" >> gpurun_out/policy_actor_1.log; tail -3 gpurun_out/policy_actor_1.log
timeout 120 python tools/time_policy.py > gpurun_out/policy_time_policy.txt 2>&1
Q1PHYS_LIB=$PWD/build/libq1phys_trace.so timeout 100 python tools/trace_actor.py act > gpurun_out/policy_trace.txt 2>&1
grep "logit error\|passed\|failed\|rc=\|Error" gpurun_out/policy_policy.log | tail -5; cat gpurun_out/policy_time_policy.txt
