#!/bin/bash
mkdir -p gpurun_out
( for n in 131072 524288; do
    python tools/time_step.py small7 $n
    Q1PHYS_SMALL_TILES=0 python tools/time_step.py six $n
  done
  Q1PHYS_LIB=$PWD/build/libq1phys_passthrough.so python tools/time_step.py passthrough7 131072
  Q1PHYS_LIB=$PWD/build/libq1phys_passthrough.so Q1_TIME_GRAPH=1 python tools/time_step.py passthrough7 131072
  Q1_TIME_RING=1 python tools/time_step.py resident 131072
  Q1_TIME_RING=1 Q1_TIME_GRAPH=1 python tools/time_step.py resident 131072
  Q1_TIME_RING=1 python tools/time_step.py resident 32768
  Q1_TIME_RING=1 python tools/time_step.py resident 12800
  python tools/time_small.py
) 2>&1 | grep -v Warning > gpurun_out/r2c_timings.txt
cat gpurun_out/r2c_timings.txt
