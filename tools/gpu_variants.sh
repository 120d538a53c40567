#!/bin/bash
# usage: tools/gpu_variants.sh  -- time the in-tree library and every build/libq1phys_*.so variant
mkdir -p gpurun_out
( python tools/time_step.py default
  for f in build/libq1phys_*.so; do Q1PHYS_LIB=$PWD/$f python tools/time_step.py $(basename $f .so); done ) 2>&1 | grep -v Warning | tee gpurun_out/variants.txt
