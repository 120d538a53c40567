#!/bin/bash
# full evidence run on a GPU box (tools/refresh_profiles.py turns gpurun_out/ into profiles/<tag>_*):
# smoke, gpu tests, both bench arms, ncu launch list, ncu --set full captures of the step kernel, the
# policy / closed-loop kernel and the in-register rollout kernel, compute-sanitizer
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b200_20steps.json 2> gpurun_out/bench_b200_20steps.err
python bench.py --scaling strong --envs 131072 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_tma -s 230 -c 2 -o gpurun_out/prof_step_tma -f python bench.py --steps 40 --warmup 10 --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_actor -s 2 -c 1 -f -o gpurun_out/prof_actor_act python tools/prof_policy.py act 1048576 > gpurun_out/prof_actor_act.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_actor -s 1 -c 1 -f -o gpurun_out/prof_actor_loop python tools/prof_policy.py loop 32768 > gpurun_out/prof_actor_loop.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 1 -c 1 -o gpurun_out/prof_rollout -f python bench.py --steps 40 --warmup 10 --no-cpu-baseline > /dev/null 2>&1
( echo "compute-sanitizer --tool memcheck python -m pytest tests/test_cuda_parity.py tests/test_api_gpu.py tests/test_recorder_gpu.py tests/test_actor_gpu.py -m gpu -k 'golden_replay or edge_cases or phys_apply or decoder or mkdemo or rollout_kernel or fused or record or eval_sim or one_launch or known_answers or ragged_k1'"
  timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_cuda_parity.py tests/test_api_gpu.py tests/test_recorder_gpu.py tests/test_actor_gpu.py -m gpu -q -k 'golden_replay or edge_cases or phys_apply or decoder or mkdemo or rollout_kernel or fused or record or eval_sim or one_launch or known_answers or ragged_k1' 2>&1 | grep -v Warning | tail -6
  echo; echo "compute-sanitizer --tool racecheck python -m pytest tests/test_cuda_parity.py tests/test_recorder_gpu.py -m gpu -k '(golden_replay and default) or many_env_record'"
  timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_cuda_parity.py tests/test_recorder_gpu.py -m gpu -q -k '(golden_replay and default) or many_env_record' 2>&1 | grep -v Warning | tail -5 ) > gpurun_out/sanitizer.txt 2>&1
tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/pytest_gpu.log; cut -c1-300 gpurun_out/bench_reference.json; cut -c1-400 gpurun_out/bench_b200.json; cat gpurun_out/sanitizer.txt
