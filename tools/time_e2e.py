"""The e2e leg of bench.py alone (VectorPhysEnv.vector_step with page-locked NumPy arrays, 2^20 envs
per rank), for each q1_step_host mode, under torchrun: how the NumPy-facing path behaves when
several GPUs share one host.  torchrun --nproc-per-node N tools/time_e2e.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from q1physrl_b200 import env as benv, sharding  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    sharding.bind_to_device_cpus(local)
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 20
for mode, envs in (("direct", {"Q1PHYS_HOST_DIRECT": "1"}), ("pipeline x2", {"Q1PHYS_HOST_DIRECT": "0", "Q1PHYS_HOST_CHUNKS": "2"}),
                   ("pipeline x4", {"Q1PHYS_HOST_DIRECT": "0", "Q1PHYS_HOST_CHUNKS": "4"})):
    os.environ.update(envs)
    e = benv.VectorPhysEnv(bench.workload_config(n), device=local, seed=rank, env_index_base=rank * n)
    nk = e.info.num_keys
    keys, mouse = e.pinned_empty((n, nk), np.uint8), e.pinned_empty((n,), np.float32)
    rng = np.random.default_rng(rank)
    keys[...] = rng.integers(0, 2, (n, nk))
    mouse[...] = rng.uniform(-10, 10, n)
    for _ in range(5):
        e.vector_step((keys, mouse), auto_reset=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    steps = 60
    t0 = time.perf_counter()
    for _ in range(steps):
        e.vector_step((keys, mouse), auto_reset=True)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{world} GPU(s), {mode}: {world * n * steps / float(dt.item()) / 1e9:.2f} G env-steps/s, "
              f"{float(dt.item()) / steps * 1e3:.3f} ms per step", flush=True)
    e.close()
    del e
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
