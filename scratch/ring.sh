#!/bin/bash
for r in 1 2 4 8; do
  python bench.py --no-cpu-baseline --steps 10000 --ring $r 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ring $r value %.3e ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
for e in 262144 524288 2097152 4194304; do
  python bench.py --no-cpu-baseline --steps 5000 --envs $e --ring 4 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('envs $e value %.3e ms/step %.5f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
done
