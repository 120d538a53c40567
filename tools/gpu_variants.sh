#!/bin/bash
# usage: tools/gpu_variants.sh [n ...] -- time the in-tree library and every build/libq1phys_*.so variant
mkdir -p gpurun_out
sizes=${@:-1048576}
( for n in $sizes; do
    python tools/time_step.py default $n
    for f in build/libq1phys_*.so; do Q1PHYS_LIB=$PWD/$f python tools/time_step.py $(basename $f .so | sed s/libq1phys_//) $n; done
  done ) 2>&1 | grep -v Warning | tee gpurun_out/variants.txt
