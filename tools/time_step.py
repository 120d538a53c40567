"""Time q1_step (k_step_tma) alone, the way bench.py's `value` leg does (ring of 4 env shards, 2^20
envs, configs[2]), for quick A/B runs of library variants: Q1PHYS_LIB=build/libq1phys_x.so
python tools/time_step.py [label]."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from q1physrl_b200 import _lib, env as benv  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(_lib.library_path())
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
ring = int(os.environ.get("Q1_TIME_RING", max(4, (1 << 22) // n)))   # 1: the same shard every tick (L2-resident when small)
steps, warmup = max(500, 8000 * (1 << 20) // n), 200
dev = torch.device("cuda", 0)
cfg = bench.workload_config(n)
lib = _lib.load()
stamps = bool(int(os.environ.get("Q1_TIME_STAMPS", "0")))   # 1: f64 key stamps -> the plain k_step kernel
envs = [benv.VectorPhysEnv(cfg, device=0, seed=r, env_index_base=r * n, f64_key_stamps=stamps) for r in range(ring)]
nk = envs[0].info.num_keys
g = torch.Generator(device=dev).manual_seed(0)
keys = [torch.randint(0, 2, (n, nk), generator=g, device=dev, dtype=torch.uint8) for _ in range(ring)]
mouse = [(torch.rand(n, generator=g, device=dev, dtype=torch.float32) * 20 - 10) for _ in range(ring)]
outs = [(torch.empty((n, 6), dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
         torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))
        for _ in range(ring)]
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
calls = [(envs[r].handle, ctypes.c_void_p(keys[r].data_ptr()), ctypes.c_void_p(mouse[r].data_ptr()),
          _lib.Q1_MOUSE_F32, *(ctypes.c_void_p(o.data_ptr()) for o in outs[r]), 1, sp) for r in range(ring)]
use_graph = bool(int(os.environ.get("Q1_TIME_GRAPH", "0")))   # 1: replay a CUDA graph of `ring` steps
if use_graph:
    for i in range(2 * ring):
        lib.q1_step(*calls[i % ring])
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    sp2 = ctypes.c_void_p(side.cuda_stream)
    gcalls = [c[:-1] + (sp2,) for c in calls]
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for i in range(ring):
            lib.q1_step(*gcalls[i])
    label += "+graph"
best = 0.0
for rep in range(3):
    if use_graph:
        for i in range(max(1, warmup // ring)):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(1, steps // ring)
        e0.record()
        for i in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = max(best, n * reps * ring / (e0.elapsed_time(e1) * 1e-3))
        continue
    for i in range(warmup):
        lib.q1_step(*calls[i % ring])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        lib.q1_step(*calls[i % ring])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = max(best, n * steps / (ms * 1e-3))
us = n / best * 1e6
us = n / best * 1e6
print(f"{label} n={n}: {best / 1e9:.2f} G env-steps/s, {us:.2f} us per tick, "
      f"{117 * best / 1e9:.0f} GB/s algorithmic")
