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
import os
use_graph = bool(int(os.environ.get("Q1_TIME_GRAPH", "0")))
if use_graph:
    for i in range(2 * ring):
        dst[i % ring].copy_(src[i % ring])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(ring):
            dst[i].copy_(src[i])
best = 1e9
for rep in range(3):
    if use_graph:
        for i in range(10):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(1, steps // ring)
        e0.record()
        for i in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (reps * ring))
        continue
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
print(f"copy{'+graph' if use_graph else ''} n={n}: {best * 1e3:.2f} us per launch of {2 * half / 1e6:.0f} MB traffic, {2 * half / best / 1e6:.0f} GB/s")
