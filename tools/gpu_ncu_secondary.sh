#!/bin/bash
# ncu full captures of the two secondary kernels: the fused tcgen05 policy kernel (closed loop, config 5)
# and the in-register rollout kernel (config 4)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_policy_act -s 20 -c 1 -o gpurun_out/prof_policy -f python bench.py --steps 40 --warmup 10 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 1 -c 1 -o gpurun_out/prof_rollout -f python bench.py --steps 40 --warmup 10 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/prof_policy.ncu-rep gpurun_out/prof_rollout.ncu-rep
