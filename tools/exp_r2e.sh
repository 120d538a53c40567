#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_api_gpu.py -x -q -s -k "fused_tcgen05 or shipped_policy" > gpurun_out/r2e_policy.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_policy.log
timeout 600 python -m pytest tests/test_actor_gpu.py -x -q -s > gpurun_out/r2e_actor.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_actor.log
tail -25 gpurun_out/r2e_policy.log; tail -40 gpurun_out/r2e_actor.log
