#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_actor_gpu.py tests/test_api_gpu.py tests/test_sampling_gpu.py -x -q > gpurun_out/r2k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_tests.log; tail -3 gpurun_out/r2k_tests.log
timeout 120 python tools/time_policy.py > gpurun_out/r2k_time_policy.txt 2>&1; cat gpurun_out/r2k_time_policy.txt
python bench.py --scaling strong --steps 2000 --no-cpu-baseline > gpurun_out/r2_bench_strong_n1.json 2> gpurun_out/r2_bench_strong_n1.err; tail -c 300 gpurun_out/r2_bench_strong_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_strong_n1.json").read().strip().splitlines()[-1])
print(d["scaling"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["envs_per_gpu"], d["zero_start_total_reward_mean"]["env_steps_per_s"])
PY
