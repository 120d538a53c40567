"""Multi-GPU use of the movement step: one process per GPU, each owning a contiguous block of the
env population.  The per-tick path has no exchange step -- every env reads and writes only its own
state (all reference ops are element-wise over the env axis, q1physrl_env/q1physrl_env/phys.py,
env.py) -- so the only collective is the reduction of the episode metrics
(`zero_start_total_reward_mean`, q1physrl/train.py:54-57, 67-71): one all-reduce of a few scalars.

Reset draws are a pure function of (seed, global env index, reset count), so a population sharded
over any number of ranks evolves exactly like the same population on one GPU.
"""
import math
from typing import Optional, Tuple


def shard_range(num_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block `[start, start + count)` of the global env index range owned by `rank`;
    block sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(int(num_envs), int(world_size))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def make_sharded_env(config, rank: int, world_size: int, device: int = 0, seed: int = 0, **kwargs):
    """`VectorPhysEnv` for this rank's block of `config.num_envs` global envs (same seed on every
    rank; `env_index_base` places the block in the global index space)."""
    import dataclasses
    from . import env as benv
    if isinstance(config, dict):
        config = benv.Config(**config)
    start, count = shard_range(config.num_envs, rank, world_size)
    local = dataclasses.replace(config, num_envs=count)
    return benv.VectorPhysEnv(local, device=device, seed=seed, env_index_base=start, **kwargs)


def bind_to_device_cpus(device: int = 0) -> Optional[str]:
    """Pin the calling process to the CPUs NVML reports as local to CUDA device `device` (same NUMA
    node / PCIe root), so that page-locked buffers allocated afterwards are first-touched next to
    the GPU that will read and write them.  One process per GPU calls this once at start-up.  Returns
    a description of the affinity set, or None when NVML or the topology information is missing (the
    process is then left as it is)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        pci = f"{torch.cuda.get_device_properties(device).pci_domain_id:08x}:{bus:02x}:" \
              f"{torch.cuda.get_device_properties(device).pci_device_id:02x}.0"
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(pci.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} CPUs local to GPU {pci}"
    except Exception:
        return None


_SUM_KEYS = ("zero_start_total_reward_sum", "zero_start_episodes", "episode_reward_sum", "episodes")


def reduce_metrics(local: dict, group=None, device: Optional[str] = None) -> dict:
    """All-reduce the per-rank episode statistics of `VectorPhysEnv.metrics()` into the global
    ones (sum of sums and counts, max of maxima) -- the one collective of this path.  Works with
    any initialised torch.distributed backend (NCCL on GPUs, gloo on CPU); without an initialised
    process group it returns the local statistics."""
    import torch
    import torch.distributed as dist
    sums = torch.tensor([float(local[k]) for k in _SUM_KEYS], dtype=torch.float64, device=device)
    mx = local.get("episode_reward_max", float("-inf"))
    mx = torch.tensor([mx if not math.isnan(mx) else float("-inf")], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    zs_sum, zs_n, ep_sum, ep_n = (float(x) for x in sums.tolist())
    return {
        "zero_start_total_reward_sum": zs_sum,
        "zero_start_episodes": int(round(zs_n)),
        "zero_start_total_reward_mean": zs_sum / zs_n if zs_n else float("nan"),
        "episode_reward_sum": ep_sum,
        "episodes": int(round(ep_n)),
        "episode_reward_mean": ep_sum / ep_n if ep_n else float("nan"),
        "episode_reward_max": float(mx.item()),
    }
