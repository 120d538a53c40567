#!/bin/bash
# usage: tools/gpu_ab.sh <tag> -- gpu tests, bench of the in-tree library and of build/libq1phys_poly.so
# (the -DQ1_POLY_SINCOS=1 build), ncu full capture of the step kernel
tag=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
for v in libm poly; do
  if [ $v = poly ]; then export Q1PHYS_LIB=$PWD/build/libq1phys_poly.so; else unset Q1PHYS_LIB; fi
  [ $v = poly ] && [ ! -f build/libq1phys_poly.so ] && continue
  python bench.py --no-cpu-baseline 2>gpurun_out/bench_${tag}_$v.err | tail -1 > gpurun_out/bench_${tag}_$v.json
  python - <<PY
import json; d=json.load(open('gpurun_out/bench_${tag}_$v.json')); print('$v value %.4e ms/step %.5f frac %.4f e2e %.4e clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']))
PY
done
unset Q1PHYS_LIB
ncu --set full --clock-control none --import-source on -k regex:k_step_tma -s 30 -c 1 -o gpurun_out/prof_step_$tag -f python bench.py --steps 40 --warmup 10 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -5
