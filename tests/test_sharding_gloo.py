"""CPU, world_size 2 over gloo: the multi-GPU host logic -- shard ranges and the one collective of
the path (episode-metric reduction) -- runs as it does under NCCL on the GPU box."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from q1physrl_b200 import sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        start, count = sharding.shard_range(1001, rank, world)
        local = {"zero_start_total_reward_sum": 100.0 * (rank + 1), "zero_start_episodes": 2 * (rank + 1),
                 "episode_reward_sum": 10.0 + rank, "episodes": 5 + rank,
                 "episode_reward_max": 7.0 if rank == 0 else 9.5}
        red = sharding.reduce_metrics(local)
        empty = sharding.reduce_metrics({"zero_start_total_reward_sum": 0.0, "zero_start_episodes": 0,
                                         "episode_reward_sum": 0.0, "episodes": 0,
                                         "episode_reward_max": float("nan")})
        out[rank] = (start, count, red, empty)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_metric_reduction_world_size_2():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0][:2] == (0, 501) and out[1][:2] == (501, 500)
    for rank in range(world):
        red, empty = out[rank][2], out[rank][3]
        assert red["zero_start_total_reward_sum"] == 300.0 and red["zero_start_episodes"] == 6
        assert red["zero_start_total_reward_mean"] == 50.0
        assert red["episode_reward_sum"] == 21.0 and red["episodes"] == 11
        assert red["episode_reward_max"] == 9.5
        assert empty["episodes"] == 0 and empty["zero_start_total_reward_mean"] != empty["zero_start_total_reward_mean"]


def test_reduce_metrics_without_process_group():
    from q1physrl_b200 import sharding
    red = sharding.reduce_metrics({"zero_start_total_reward_sum": 30.0, "zero_start_episodes": 3,
                                   "episode_reward_sum": 50.0, "episodes": 10, "episode_reward_max": 8.0})
    assert red["zero_start_total_reward_mean"] == 10.0 and red["episode_reward_mean"] == 5.0
