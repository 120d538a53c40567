#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_actor -s 2 -c 1 -f -o gpurun_out/r2_actor_act python tools/prof_policy.py act 131072 > gpurun_out/r2_actor_act.log 2>&1; tail -2 gpurun_out/r2_actor_act.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_actor -s 1 -c 1 -f -o gpurun_out/r2_actor_loop python tools/prof_policy.py loop 32768 > gpurun_out/r2_actor_loop.log 2>&1; tail -2 gpurun_out/r2_actor_loop.log
timeout 120 python -m pytest tests/test_api_gpu.py -x -q -s -k "fused_tcgen05 or shipped_policy" > gpurun_out/r2f_policy.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_policy.log
for i in 1 2 3; do timeout 120 python -m pytest tests/test_actor_gpu.py -x -q -s > gpurun_out/r2f_actor_$i.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_actor_$i.log; tail -3 gpurun_out/r2f_actor_$i.log; done
timeout 120 python tools/time_policy.py > gpurun_out/r2f_time_policy.txt 2>&1
Q1PHYS_LIB=$PWD/build/libq1phys_tanhf32.so timeout 120 python tools/time_policy.py > gpurun_out/r2f_time_policy_f32.txt 2>&1
grep "logit error\|passed\|failed\|rc=" gpurun_out/r2f_policy.log; cat gpurun_out/r2f_time_policy.txt; echo "--- tanh f32 variant"; cat gpurun_out/r2f_time_policy_f32.txt
