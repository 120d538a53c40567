"""What a plain device-to-device copy achieves under the step kernel's launch pattern: the same
traffic per launch (117 B x n envs: half read, half written), ring of 4 buffer pairs so that nothing is
L2-resident, back-to-back launches.  Gives the practical ceiling to hold k_step_tma against."""
import sys

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
half = 117 * n // 2
ring = max(4, (1 << 22) // n)
src = [torch.empty(half, dtype=torch.uint8, device="cuda").random_() for _ in range(ring)]
dst = [torch.empty(half, dtype=torch.uint8, device="cuda") for _ in range(ring)]
steps = max(500, 4000 * (1 << 20) // n)
best = 1e9
for rep in range(3):
    for i in range(100):
        dst[i % ring].copy_(src[i % ring])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        dst[i % ring].copy_(src[i % ring])
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / steps)
print(f"copy n={n}: {best * 1e3:.2f} us per launch of {2 * half / 1e6:.0f} MB traffic, {2 * half / best / 1e6:.0f} GB/s")
