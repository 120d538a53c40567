#!/bin/bash
mkdir -p gpurun_out
( for n in 131072 262144 1048576; do
    for g in 0 1; do
      Q1_TIME_GRAPH=$g python tools/time_step.py default $n
      for v in ctas7 ctas8; do Q1_TIME_GRAPH=$g Q1PHYS_LIB=$PWD/build/libq1phys_$v.so python tools/time_step.py $v $n; done
      Q1_TIME_GRAPH=$g python tools/time_copy.py $n
    done
  done ) 2>&1 | grep -v Warning > gpurun_out/r2b_timings.txt
cat gpurun_out/r2b_timings.txt
