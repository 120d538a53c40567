#!/bin/bash
# usage: tools/build_variants.sh name:"-DFLAG=1 ..." [...]  -> build/libq1phys_<name>.so (for Q1PHYS_LIB=... A/B runs)
mkdir -p build
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared \
       $flags -o build/libq1phys_$name.so q1physrl_b200/csrc/q1phys.cu q1physrl_b200/csrc/q1_actor.cu &
done
wait
ls -la build/
