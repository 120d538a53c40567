#!/bin/bash
for c in 1 2 4 8; do
Q1PHYS_HOST_CHUNKS=$c python bench.py --no-cpu-baseline --steps 2000 --e2e-steps 60 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunks $c e2e %.3e' % (d['e2e']['value']))"
done
python - <<'PY'
import torch, time
n=1<<20
h=torch.empty(n*30, dtype=torch.uint8).pin_memory(); d=torch.empty(n*30, dtype=torch.uint8, device='cuda')
h2=torch.empty(n*7, dtype=torch.uint8).pin_memory(); d2=torch.empty(n*7, dtype=torch.uint8, device='cuda')
for _ in range(3): h.copy_(d); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(20): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
print('D2H 31.5MB: %.3f ms  %.1f GB/s' % (dt*1e3, n*30/dt/1e9))
t=time.perf_counter()
for _ in range(20): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
print('H2D 7.3MB: %.3f ms  %.1f GB/s' % (dt*1e3, n*7/dt/1e9))
PY
