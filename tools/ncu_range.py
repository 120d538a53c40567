"""The step kernel's DRAM traffic over a RANGE of consecutive ring launches (for ncu --replay-mode
application-range): warm-up, cudaProfilerStart, LAUNCHES q1_step calls round-robin over the ring,
cudaProfilerStop.  A single profiled launch under-reports writes (its dirty lines are still in the L2
when it ends); over a range every launch also pays the write-backs of its predecessors."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ring = bench.StepRing(n, max(4, int(-(-2 * bench.L2_BYTES // (117 * n)))), 0, 0, 0)
for i in range(64):
    ring.step(i)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for i in range(launches):
    ring.step(64 + i)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"range: {launches} launches of {n} envs, {ring.bytes_per_env_step} algorithmic B/env-step")
