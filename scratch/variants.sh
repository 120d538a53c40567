#!/bin/bash
for f in scratch/libs/*.so ""; do
  echo "== $f"
  Q1PHYS_LIB=${f:+$PWD/$f} python bench.py --no-cpu-baseline --steps 10000 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.3e ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
