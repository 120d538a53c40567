#!/bin/bash
# usage: tools/gpu_iter.sh <tag>  -- quick parity subset + bench + ncu full capture of k_step
tag=$1
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "golden or free_running or teacher" 2>&1 | tail -3
python tools/selftest.py 2>&1 | tail -1
python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$tag.json
python - <<PY
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print('value %.3e ms/step %.5f frac %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))
PY
ncu --set full --clock-control none --import-source on -k regex:k_step_tma -s 30 -c 1 -o gpurun_out/prof_step_$tag -f python bench.py --steps 40 --warmup 10 --no-cpu-baseline > /dev/null 2>&1
