#!/bin/bash
# usage: tools/gpu_e2e.sh <tag> -- host-path test + bench with the direct (mapped host memory) and the
# pipelined q1_step_host
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "host_pipeline or sincos or golden" 2>&1 | tail -4
for v in direct pipe2; do
  case $v in direct) export Q1PHYS_HOST_DIRECT=1;; pipe2) export Q1PHYS_HOST_DIRECT=0 Q1PHYS_HOST_CHUNKS=2;; pipe4) export Q1PHYS_HOST_DIRECT=0 Q1PHYS_HOST_CHUNKS=4;; esac
  python bench.py --no-cpu-baseline 2>gpurun_out/bench_${tag}_$v.err | tail -1 > gpurun_out/bench_${tag}_$v.json
  python - <<PY
import json; d=json.load(open('gpurun_out/bench_${tag}_$v.json')); print('$v value %.4e ms/step %.5f frac %.4f e2e %.4e' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))
PY
done
