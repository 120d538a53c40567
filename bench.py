#!/usr/bin/env python
"""bench.py -- env-steps/s of the q1physrl_env movement step on N B200s (one process per GPU).

A "step" is one lockstep tick (`VectorPhysEnv.vector_step` == one `k_step` launch) over one batch of
2^20 envs per GPU.  Workload = BASELINE.json configs[2]: num_envs 1,048,576, the 100 m Config
(reference data/params.yml:16-33) with zero_start_prob 1 and auto_jump, uniform random actions.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # CPU arm: the C port of the
                                                                   # reference's NumPy path, all cores

Prints ONE JSON line (rank 0).  `value` is device-timed with the actions already resident in HBM;
`e2e` goes through the public `VectorPhysEnv.vector_step` with host (page-locked) arrays, host<->
device copies inside the timed region.  See DESIGN.md section "Measurement".
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NUM_ENVS = 1 << 20
METRIC = "env_steps_per_s"
UNIT = "env-steps/s"
WORKLOAD = ("configs[2]: num_envs=1048576 per GPU, 100m Config (params.yml) with zero_start_prob=1 + "
            "auto_jump, uniform random actions, one lockstep tick per step")
CONFIG_100M = dict(
    num_envs=None, action_range=10, allow_jump=True, allow_yaw=True, auto_jump=False,
    discrete_yaw_steps=-1, fmove_max=800, hover=False, initial_yaw_range=(0, 360),
    key_press_delay=0.3, max_initial_speed=700, smooth_keys=True, smove_max=1060,
    speed_reward=False, time_delta=0.013888888888888, time_limit=10, zero_start_prob=0.01)


def workload_config(num_envs):
    return dict(CONFIG_100M, num_envs=num_envs, zero_start_prob=1.0, auto_jump=True)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_step_tma launch, from the committed
    `ncu --set full` capture of this workload (profiles/r1_step_tma_ncu.json); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_step_tma_ncu.json")) as f:
            d = json.load(f)
        return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:
        return None


# ------------------------------------------------------------------------------ CPU arm / baseline

def cpu_port_throughput(num_envs, ticks, threads, seed=0):
    """Time the C oracle (a port of the reference's NumPy arithmetic, oracle/q1_oracle.c) on
    `threads` host threads, each owning a contiguous slice of the envs.  -> env-steps/s."""
    from oracle import q1_oracle as qo
    cfg = workload_config(num_envs)
    bounds = np.linspace(0, num_envs, threads + 1).astype(np.int64)
    envs, acts = [], []
    rng = np.random.default_rng(seed)
    for i in range(threads):
        n = int(bounds[i + 1] - bounds[i])
        o = qo.OracleEnv(cfg, num_envs=n)
        o.reset_from_philox(seed, int(bounds[i]), 1)
        envs.append(o)
        acts.append((rng.integers(0, 2, size=(n, o.nk)).astype(np.uint8),
                     rng.uniform(-10, 10, size=n).astype(np.float32).astype(np.float64)))

    def work(i, k):
        for _ in range(k):
            envs[i].step(*acts[i])

    def run(k):
        ts = [threading.Thread(target=work, args=(i, k)) for i in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    run(1)
    dt = run(ticks)
    return num_envs * ticks / dt, dt


def run_reference_arm(args, rank):
    """--impl reference: the CPU implementation of the path (C port of the NumPy reference), all
    host threads; every step is a bounded sample (SAMPLE_ENVS envs) of the workload."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_envs = 1 << 18
    from oracle import q1_oracle as qo
    qo.build()
    cpu_port_throughput(sample_envs, max(1, args.warmup), threads)
    value, dt = cpu_port_throughput(sample_envs, args.steps, threads)
    sample = (f"{sample_envs} of {NUM_ENVS} envs per step (same Config and action distribution), "
              f"{args.steps} ticks, {threads} threads x {sample_envs // threads} envs")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ clocks

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self._nv:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------ CUDA arm

def run_cuda_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from q1physrl_b200 import _lib, env as benv, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the movement step has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if world > 1:
        # one process per GPU: run it (and first-touch its page-locked buffers) on the CPUs next to
        # that GPU, so that the e2e leg's PCIe traffic does not cross the socket interconnect
        numa = sharding.bind_to_device_cpus(local_rank)
        dist.init_process_group("nccl", device_id=dev)

    n = args.envs
    ring = args.ring
    cfg = workload_config(n)
    lib = _lib.load()
    # Ring of independent env shards: each timed step ticks the next shard, so by the time a shard
    # comes round again its state and buffers (ring x ~82 MB) have been evicted from the 126 MB L2.
    envs = [benv.VectorPhysEnv(cfg, device=local_rank, seed=args.seed + r,
                               env_index_base=(rank * ring + r) * n) for r in range(ring)]
    nk = envs[0].info.num_keys
    g = torch.Generator(device=dev).manual_seed(args.seed + rank)
    keys = [torch.randint(0, 2, (n, nk), generator=g, device=dev, dtype=torch.uint8) for _ in range(ring)]
    mouse = [(torch.rand(n, generator=g, device=dev, dtype=torch.float32) * 20 - 10) for _ in range(ring)]
    outs = [(torch.empty((n, 6), dtype=torch.float32, device=dev),
             torch.empty(n, dtype=torch.float32, device=dev),
             torch.empty(n, dtype=torch.uint8, device=dev),
             torch.empty(n, dtype=torch.uint8, device=dev)) for _ in range(ring)]
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    calls = []
    for r in range(ring):
        o = outs[r]
        calls.append((envs[r].handle, ctypes.c_void_p(keys[r].data_ptr()),
                      ctypes.c_void_p(mouse[r].data_ptr()), _lib.Q1_MOUSE_F32,
                      ctypes.c_void_p(o[0].data_ptr()), ctypes.c_void_p(o[1].data_ptr()),
                      ctypes.c_void_p(o[2].data_ptr()), ctypes.c_void_p(o[3].data_ptr()), 1, sp))
    q1_step = lib.q1_step

    def step(i):
        rc = q1_step(*calls[i % ring])
        if rc:
            _lib.check(rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record(stream)
        for i in range(args.steps):
            step(args.warmup + i)
        ev1.record(stream)
        barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * n * args.steps / (elapsed_ms * 1e-3)

    # ---- what a plain device-to-device copy achieves under the same launch pattern (same traffic per
    # launch, ring of buffers so that nothing stays in L2, back-to-back launches): the practical
    # ceiling at this launch size, reported beside the roofline (the peak itself is reached only by
    # launches that move gigabytes)
    copy_gbs = None
    if rank == 0:
        half = (2 * envs[0].info.state_bytes_per_env + nk + 34) * n // 2
        csrc = [torch.empty(half, dtype=torch.uint8, device=dev) for _ in range(ring)]
        cdst = [torch.empty(half, dtype=torch.uint8, device=dev) for _ in range(ring)]
        for i in range(100):
            cdst[i % ring].copy_(csrc[i % ring])
        torch.cuda.synchronize(dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        csteps = 2000
        c0.record(stream)
        for i in range(csteps):
            cdst[i % ring].copy_(csrc[i % ring])
        c1.record(stream)
        torch.cuda.synchronize(dev)
        copy_gbs = 2 * half * csteps / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del csrc, cdst
    barrier()

    # ---- e2e: the public API with host arrays (page-locked), copies inside the timed region
    e = envs[0]
    h_keys = e.pinned_empty((n, nk), np.uint8)
    h_mouse = e.pinned_empty((n,), np.float32)
    h_keys[...] = keys[0].cpu().numpy()
    h_mouse[...] = mouse[0].cpu().numpy()
    e2e_steps = max(3, min(args.e2e_steps, args.steps))
    for _ in range(3):
        e.vector_step((h_keys, h_mouse), auto_reset=True)
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for _ in range(e2e_steps):
        obs, rew, done, infos = e.vector_step((h_keys, h_mouse), auto_reset=True)
        checksum += float(rew[0])                                  # the result is read on the host
    t1 = time.perf_counter()
    te = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(te.item())
    h2d = n * (nk + 4)
    d2h = n * (24 + 4 + 1 + 1)

    # ---- BASELINE config 5 + the one collective of this path: zero_start_total_reward_mean of the
    # reference's shipped policy (data/checkpoints/wr, weights in tests/golden/wr_policy.npz), rolled
    # out closed-loop on the device over 32768 envs per GPU (262144 over 8), metric all-reduced.
    from q1physrl_b200 import policy as bpolicy
    policy_path = os.path.join(ROOT, "tests", "golden", "wr_policy.npz")
    n5, ticks5 = 1 << 15, 1500
    if os.path.exists(policy_path):
        pol, env_cfg = bpolicy.FusedMLPPolicy.from_npz(policy_path, device=local_rank, seed=args.seed)
        cfg5 = dict(env_cfg, initial_yaw_range=tuple(env_cfg["initial_yaw_range"]), num_envs=n5)
        tracked = benv.VectorPhysEnv(cfg5, device=local_rank, seed=args.seed,
                                     env_index_base=rank * n5, track_returns=True)
        torch.cuda.synchronize(dev)
        timing5 = {}
        bpolicy.rollout(tracked, pol, ticks5, timing=timing5)
        torch.cuda.synchronize(dev)
        rate5 = world * n5 * timing5["ticks"] / timing5["seconds"] if timing5.get("seconds") else float("nan")
        policy_desc = (f"reference checkpoint data/checkpoints/wr (stochastic Q1PhysActionDist), params.json "
                       f"env_config, {n5} envs/GPU x {ticks5} ticks closed loop on the device (fused tcgen05 "
                       f"policy kernel + step kernel, 8 ticks per CUDA graph), "
                       f"{rate5:.3e} env-steps/s incl. the policy (replays timed with CUDA events)")
    else:
        tracked = benv.VectorPhysEnv(workload_config(n5), device=local_rank, seed=args.seed,
                                     env_index_base=rank * n5, track_returns=True)
        tracked.rollout("strafe_jump", 722, policy_seed=1)
        policy_desc = f"scripted strafe_jump, 722 ticks, {n5} envs/GPU"
    red = sharding.reduce_metrics(tracked.metrics(), device=dev)
    zs_mean, zs_episodes = red["zero_start_total_reward_mean"], red["zero_start_episodes"]

    # ---- BASELINE config 4: 2^20 envs sharded 131072 per GPU, scripted strafe-jump policy, 10k ticks
    # in the multi-tick in-register rollout kernel (no per-tick HBM traffic)
    n4, ticks4 = 1 << 17, 10000
    e4 = benv.VectorPhysEnv(dict(CONFIG_100M, num_envs=n4, zero_start_prob=1.0), device=local_rank,
                            seed=args.seed, env_index_base=rank * n4)
    e4.rollout("strafe_jump", 100)
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(stream)
    e4.rollout("strafe_jump", ticks4)
    r1.record(stream)
    barrier()
    t4 = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t4, op=dist.ReduceOp.MAX)
    rollout_value = world * n4 * ticks4 / (float(t4.item()) * 1e-3)

    if rank == 0:
        info = envs[0].info
        state_b = info.state_bytes_per_env
        bytes_per_env_step = 2 * state_b + nk + 4 + 24 + 4 + 1 + 1
        peak, peak_src = measured_peak_gbs()
        per_launch_s = elapsed_ms * 1e-3 / args.steps
        achieved = bytes_per_env_step * n / per_launch_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "ring_shards": ring,
                       "l2_policy": f"inputs larger than L2: ring of {ring} env shards x "
                                    f"{(2 * state_b + nk + 34) * n / 1e6:.0f} MB touched per step",
                       "state_bytes_per_env": state_b, "key_timers": "f64" if info.f64_stamps else "u8",
                       "launch": "one q1_step call (k_step_tma launch, programmatic dependent launch) per "
                                 "step on torch's current stream"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "cpu_affinity": numa,
                    "api": "VectorPhysEnv.vector_step((keys, mouse)) with page-locked NumPy arrays: q1_step_host "
                           "launches the step kernel on the mapped host buffers (actions read and results "
                           "written over PCIe inside the launch), then synchronises"},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_bytes(), "kernel": "k_step_tma",
                         "bytes_per_env_step": bytes_per_env_step, "peak_source": peak_src,
                         "same_size_copy_gbs": copy_gbs,
                         "same_size_copy_note": "torch D2D copy_ moving the same bytes per launch, same ring, "
                                                "timed live in this run: what any kernel launched at this size "
                                                "can reach"},
            "clocks": clocks.summary(),
            "config4_rollout": {"value": rollout_value, "unit": UNIT, "envs_per_gpu": n4, "ticks": ticks4,
                                "policy": "scripted strafe_jump generated on the device",
                                "kernel": "k_rollout: one launch, state in registers for all ticks"},
            "zero_start_total_reward_mean": {"value": zs_mean, "episodes": zs_episodes,
                                             "policy": policy_desc,
                                             "collective": "all_reduce(sum) of (sum, count)" if world > 1 else "none (1 GPU)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample_envs = NUM_ENVS
            v0, dt0 = cpu_port_throughput(sample_envs, 8, threads)           # calibrate: ~10 s of work
            sample_ticks = int(min(4000, max(16, 10.0 * v0 / sample_envs)))
            v, dt = cpu_port_throughput(sample_envs, sample_ticks, threads)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{sample_envs} envs x {sample_ticks} ticks of the same Config and action "
                          f"distribution, C port of the reference NumPy path (oracle/q1_oracle.c), "
                          f"{threads} threads, {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--envs", type=int, default=NUM_ENVS, help="envs per GPU")
    ap.add_argument("--ring", type=int, default=4, help="env shards ticked round-robin (L2 eviction)")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 100
        args.warmup = args.warmup if args.warmup is not None else 3
        run_reference_arm(args, rank)
        return
    args.steps = args.steps if args.steps is not None else 20000
    args.warmup = max(3, args.warmup if args.warmup is not None else 200)
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit(f"bench.py: --gpus {args.gpus} needs torchrun (--nproc-per-node {args.gpus})")
    run_cuda_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
