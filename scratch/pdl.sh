#!/bin/bash
for i in 1 2; do
python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('PDL    value %.3e ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
Q1PHYS_NO_PDL=1 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no PDL value %.3e ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
