#!/bin/bash
# round-2 first GPU pass: new recorder tests, full gpu suite, small-launch timings, cold-sincos A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_recorder_gpu.py -x -q -m gpu > gpurun_out/r2a_recorder.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_recorder.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_pytest.log
( for n in 1048576 131072; do
    python tools/time_step.py default $n
    for f in build/libq1phys_*.so; do Q1PHYS_LIB=$PWD/$f python tools/time_step.py $(basename $f .so | sed s/libq1phys_//) $n; done
    Q1_TIME_STAMPS=1 python tools/time_step.py plain_kstep $n
    python tools/time_copy.py $n
  done ) 2>&1 | grep -v Warning > gpurun_out/r2a_timings.txt
tail -3 gpurun_out/r2a_recorder.log; tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_timings.txt
