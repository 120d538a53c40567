#!/usr/bin/env python
"""bench.py -- env-steps/s of the q1physrl_env movement step on N B200s (one process per GPU).

A "step" is one lockstep tick (`VectorPhysEnv.vector_step` == one `k_step_tma` launch) over one batch
of envs per GPU.  Workload = BASELINE.json configs[2]: num_envs 1,048,576, the 100 m Config (reference
data/params.yml:16-33) with zero_start_prob 1 and auto_jump, uniform random actions.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --scaling strong [...]                         # 2^20 envs TOTAL, split over the GPUs
    python bench.py --impl reference [...]                         # CPU arm: the C port of the
                                                                   # reference's NumPy path, all cores

Prints ONE JSON line (rank 0).  `value` is device-timed with the actions already resident in HBM;
`e2e` goes through the public `VectorPhysEnv.vector_step` with host (page-locked) arrays, host<->
device copies inside the timed region.  The default (weak-scaling) line also carries the strong-scaling
measurement (`strong`: 2^20 envs total, i.e. 131 072 per GPU at N = 8), the reference's own NumPy path
timed on this box's host cores (`cpu_baseline.numpy`, SURVEY.md 8(d) tiers T1 / T2 / T3), the host
ceiling of the e2e leg and the per-call latency at RLLib's production batch.  See DESIGN.md section 6.
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
WORKLOAD_STRONG = ("configs[2] split over the GPUs (north_star target read literally): num_envs=1048576 in "
                   "TOTAL, 100m Config (params.yml) with zero_start_prob=1 + auto_jump, uniform random "
                   "actions, one lockstep tick per step")
CONFIG_100M = dict(
    num_envs=None, action_range=10, allow_jump=True, allow_yaw=True, auto_jump=False,
    discrete_yaw_steps=-1, fmove_max=800, hover=False, initial_yaw_range=(0, 360),
    key_press_delay=0.3, max_initial_speed=700, smooth_keys=True, smove_max=1060,
    speed_reward=False, time_delta=0.013888888888888, time_limit=10, zero_start_prob=0.01)
L2_BYTES = 126e6


def workload_config(num_envs):
    return dict(CONFIG_100M, num_envs=num_envs, zero_start_prob=1.0, auto_jump=True)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """DRAM bytes per k_step_tma launch from the committed ncu RANGE capture over consecutive ring
    launches of this workload (profiles/r2_step_tma_range.json, written by tools/ncu_range.sh): read +
    write of the whole range / launches in it.  (A single profiled launch under-reports the writes:
    its dirty lines are still in L2 when the launch ends.)  -> (bytes per launch | None, how)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_step_tma_range.json")) as f:
            d = json.load(f)
        return (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["launches"], d.get("how", "")
    except Exception:
        return None, "no range capture committed"


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ------------------------------------------------------------------------------ CPU arm / baseline

def cpu_port_throughput(num_envs, threads, min_seconds, min_ticks=1, seed=0):
    """Time the C oracle (a port of the reference's NumPy arithmetic, oracle/q1_oracle.c) on `threads`
    host threads, each owning a contiguous slice of the envs.  The threads exist, have built their
    envs and have done one warm-up tick before the clock starts (barrier); each then runs whole ticks
    until `min_seconds` have passed and it has done `min_ticks`.  -> (env-steps/s, seconds, ticks)."""
    from oracle import q1_oracle as qo
    cfg = workload_config(num_envs)
    bounds = np.linspace(0, num_envs, threads + 1).astype(np.int64)
    start = threading.Barrier(threads + 1)
    t0 = [0.0]
    ticks_done, ends = [0] * threads, [0.0] * threads
    errors = []

    def work(i):
        try:
            n = int(bounds[i + 1] - bounds[i])
            rng = np.random.default_rng(seed + i)
            o = qo.OracleEnv(cfg, num_envs=n)
            o.reset_from_philox(seed, int(bounds[i]), 1)
            keys = rng.integers(0, 2, size=(n, o.nk)).astype(np.uint8)
            mouse = rng.uniform(-10, 10, size=n).astype(np.float32).astype(np.float64)
            o.step(keys, mouse)
        except Exception as exc:          # release the barrier so the failure is reported, not hung on
            errors.append(exc)
            start.abort()
            return
        start.wait()
        k = 0
        while True:
            o.step(keys, mouse)
            k += 1
            now = time.perf_counter()
            if k >= min_ticks and now - t0[0] >= min_seconds:
                break
        ticks_done[i], ends[i] = k, now

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    t0[0] = time.perf_counter() + 3600.0      # threads that pass the barrier first cannot stop early
    try:
        start.wait()
    except threading.BrokenBarrierError:
        raise RuntimeError(f"cpu_port_throughput worker failed: {errors[:1]}")
    t0[0] = time.perf_counter()
    for t in ts:
        t.join()
    elapsed = max(ends) - t0[0]
    env_steps = sum(int(bounds[i + 1] - bounds[i]) * ticks_done[i] for i in range(threads))
    return env_steps / elapsed, elapsed, round(env_steps / num_envs, 1)    # ticks of the whole batch


def numpy_reference_tiers(seconds=2.5):
    """The UNMODIFIED reference NumPy path (staged byte for byte in oracle/_ref by oracle/stage_ref.py,
    or the checkout where mounted) on this box's host cores: SURVEY.md 8(d) tiers at P = 1 process and
    P = all cores.  None when the reference is not available."""
    try:
        from oracle import numpy_tiers, refshim
        if not refshim.available():
            return None
        cores = os.cpu_count() or 1
        out = {"kind": "reference", "source": os.path.relpath(refshim.REFERENCE_ROOT, ROOT)
               if refshim.REFERENCE_ROOT.startswith(ROOT) else refshim.REFERENCE_ROOT,
               "numpy": np.__version__, "unit": UNIT,
               "tiers": {"T1": "VectorPhysEnv.vector_step(list of per-env action tuples): the public path",
                         "T2": "vector_step fed an array, _fix_actions bypassed: the env arithmetic",
                         "T3": "phys.apply alone"}}
        for tier, per_proc in (("T1", 16384), ("T2", 65536), ("T3", 65536)):
            one = numpy_tiers.measure(tier, per_proc, 1, seconds, workload_config(per_proc))
            per_all = max(1024, min(per_proc, NUM_ENVS // cores))
            allc = numpy_tiers.measure(tier, per_all, cores, seconds, workload_config(per_all))
            out[tier] = {"p1": one["value"], f"p{cores}": allc["value"], "procs": cores,
                         "sample": f"P=1: {per_proc} envs x {one['ticks_per_proc']} ticks; P={cores}: {per_all} envs "
                                   f"per process x >= {allc['ticks_per_proc']} ticks (same Config and action "
                                   f"distribution as the workload)"}
        return out
    except Exception as exc:                      # measurement aid only: never take the bench line down
        return {"kind": "reference", "error": f"{type(exc).__name__}: {exc}"}


def run_reference_arm(args, rank):
    """--impl reference: the CPU implementation of the path (C port of the NumPy reference) on all host
    threads over the full 2^20 envs of the workload, timed for >= 2 s whatever --steps says."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import q1_oracle as qo
    qo.build()
    value, dt, ticks = cpu_port_throughput(NUM_ENVS, threads, min_seconds=2.0, min_ticks=max(1, args.steps))
    sample = (f"all {NUM_ENVS} envs of one GPU's batch (same Config and action distribution; throughput per "
              f"env-step does not depend on how many such batches there are), {ticks} lockstep ticks in "
              f"{dt:.2f} s, {threads} threads x {NUM_ENVS // threads} envs, threads built and warmed before t0")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "timed_steps": ticks,
        "ms_per_step": 1e3 * NUM_ENVS / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": NUM_ENVS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "cpu": cpu_model(), "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_numpy_tiers:
        line["cpu_baseline"]["numpy"] = numpy_reference_tiers()
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

    def sample_once(self):
        nv = self._nv
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

    def _loop(self):
        while not self._stop.is_set():
            self.sample_once()
            time.sleep(0.002)

    def __enter__(self):
        if self._nv:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self, note=None):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        out = {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------ CUDA arm

class StepRing:
    """A ring of independent env shards ticked round-robin, so that by the time a shard comes round
    again its state and buffers have been evicted from the 126 MB L2."""

    def __init__(self, n, ring, device_index, seed, rank):
        import torch
        from q1physrl_b200 import _lib, env as benv
        dev = torch.device("cuda", device_index)
        cfg = workload_config(n)
        self.n, self.ring = n, ring
        self.envs = [benv.VectorPhysEnv(cfg, device=device_index, seed=seed + r,
                                        env_index_base=(rank * ring + r) * n) for r in range(ring)]
        self.nk = self.envs[0].info.num_keys
        g = torch.Generator(device=dev).manual_seed(seed + rank)
        self.keys = [torch.randint(0, 2, (n, self.nk), generator=g, device=dev, dtype=torch.uint8)
                     for _ in range(ring)]
        self.mouse = [(torch.rand(n, generator=g, device=dev, dtype=torch.float32) * 20 - 10)
                      for _ in range(ring)]
        self.outs = [(torch.empty((n, 6), dtype=torch.float32, device=dev),
                      torch.empty(n, dtype=torch.float32, device=dev),
                      torch.empty(n, dtype=torch.uint8, device=dev),
                      torch.empty(n, dtype=torch.uint8, device=dev)) for _ in range(ring)]
        self.stream = torch.cuda.current_stream(dev)
        sp = ctypes.c_void_p(self.stream.cuda_stream)
        self.calls = []
        for r in range(ring):
            o = self.outs[r]
            self.calls.append((self.envs[r].handle, ctypes.c_void_p(self.keys[r].data_ptr()),
                               ctypes.c_void_p(self.mouse[r].data_ptr()), _lib.Q1_MOUSE_F32,
                               ctypes.c_void_p(o[0].data_ptr()), ctypes.c_void_p(o[1].data_ptr()),
                               ctypes.c_void_p(o[2].data_ptr()), ctypes.c_void_p(o[3].data_ptr()), 1, sp))
        self._step = _lib.load().q1_step
        self._check = _lib.check
        info = self.envs[0].info
        self.state_bytes = info.state_bytes_per_env
        self.f64_stamps = bool(info.f64_stamps)
        self.bytes_per_env_step = 2 * self.state_bytes + self.nk + 4 + 24 + 4 + 1 + 1

    def step(self, i):
        rc = self._step(*self.calls[i % self.ring])
        if rc:
            self._check(rc)

    def close(self):
        for e in self.envs:
            e.close()


def time_ring(ring, steps, warmup, barrier, world, dist, dev, clocks=None, min_sample_s=0.1):
    """W warm-up steps, then exactly `steps` timed steps between barriers, CUDA events on the launching
    stream, max over ranks.  -> elapsed ms.  With `clocks`, stepping continues untimed after the timed
    region until the sampler has seen `min_sample_s` of this load (a 20-step region is 0.5 ms long)."""
    import torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()   # the ranks start their warm-up together, so that none of them sits idle at the next barrier
                # waiting for the others: the first steps behind an idle device run slower
    for i in range(warmup):
        ring.step(i)
    barrier()
    if os.environ.get("Q1_BENCH_SLEEP_US"):   # diagnostic: an idle device before the timed region
        time.sleep(float(os.environ["Q1_BENCH_SLEEP_US"]) * 1e-6)
    t_wall = time.perf_counter()
    ev0.record(ring.stream)
    marks = []
    for i in range(steps):
        ring.step(warmup + i)
        if os.environ.get("Q1_BENCH_STEP_EVENTS"):      # debugging aid: where inside a short timed region the time goes
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record(ring.stream)
    ev1.record(ring.stream)
    if marks:
        torch.cuda.synchronize(dev)
        ends = [ev0.elapsed_time(m) * 1e3 for m in marks]
        print(f"rank {os.environ.get('RANK', 0)}: first step ends {ends[0]:.1f} us after the start event; step "
              f"durations after it: {[round(b - a, 1) for a, b in zip(ends, ends[1:])][:24]}", file=sys.stderr, flush=True)
    if clocks is not None:
        i = warmup + steps
        while time.perf_counter() - t_wall < min_sample_s:
            for _ in range(64):
                ring.step(i)
                i += 1
            clocks.sample_once()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def same_size_copy_gbs(half_bytes, ring, dev, stream):
    """What a plain device-to-device copy achieves under the step kernel's launch pattern (same bytes
    per launch, ring of buffers so that nothing stays in L2, back-to-back launches)."""
    import torch
    csrc = [torch.empty(half_bytes, dtype=torch.uint8, device=dev) for _ in range(ring)]
    cdst = [torch.empty(half_bytes, dtype=torch.uint8, device=dev) for _ in range(ring)]
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
    return 2 * half_bytes * csteps / (c0.elapsed_time(c1) * 1e-3) / 1e9


def host_copy_ceiling(n, nk, dev, barrier, world, dist, reps=30):
    """What this host can move for one e2e step per GPU with the copy engines alone, all ranks at once:
    cudaMemcpyAsync of the step's results (30 B/env) device -> page-locked host and of its actions
    (nk + 4 B/env) host -> device on two streams, no kernel.  -> env-steps/s over all ranks."""
    import torch
    d2h_bytes, h2d_bytes = n * 30, n * (nk + 4)
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def one():
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_out, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    for _ in range(3):
        one()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        one()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return world * n * reps / float(t.item())


def small_batch_latency(device_index):
    """Per-call latency of the NumPy-facing API at the reference's production shape: RLLib drives
    VectorPhysEnv(num_envs=100) per worker (params.yml:28) with a list of per-env action tuples."""
    from q1physrl_b200 import env as benv
    n = 100
    e = benv.VectorPhysEnv(dict(CONFIG_100M, num_envs=n), device=device_index, seed=1)
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 2, (n, 4)).astype(np.uint8)
    mouse = rng.uniform(-10, 10, n).astype(np.float32)
    tuples = [tuple([int(k) for k in keys[i]] + [np.array([mouse[i]], np.float32)]) for i in range(n)]
    out = {"num_envs": n}
    for name, act in (("arrays_us", (keys, mouse)), ("list_of_tuples_us", tuples)):
        for _ in range(200):
            e.vector_step(act)
        best = float("inf")
        for _ in range(5):
            t = time.perf_counter()
            for _ in range(400):
                e.vector_step(act)
            best = min(best, (time.perf_counter() - t) / 400)
        out[name] = best * 1e6
    e.close()
    return out


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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    strong_headline = args.scaling == "strong"
    peak, peak_src = measured_peak_gbs()

    def ring_depth(n, bytes_per_env):
        # deep enough that a shard's working set is evicted before its turn comes again: the ring's
        # bytes are at least twice the L2
        return max(4, int(np.ceil(2 * L2_BYTES / (bytes_per_env * n))))

    # ---- weak scaling: 2^20 envs per GPU (the line's `value` unless --scaling strong)
    n = args.envs
    ring_n = args.ring or ring_depth(n, 117)
    ring = StepRing(n, ring_n, local_rank, args.seed, rank)
    nk = ring.nk
    # the contract's W warm-up steps are a minimum: a launch is 23 us, and W = 5 of them leave the clocks
    # and the instruction cache cold; run at least 200 (untimed either way, and reported)
    warmup_run = max(args.warmup, 200)
    with ClockSampler(local_rank) as clocks:
        elapsed_ms = time_ring(ring, args.steps, warmup_run, barrier, world, dist, dev, clocks)
    value = world * n * args.steps / (elapsed_ms * 1e-3)
    per_launch_s = elapsed_ms * 1e-3 / args.steps
    achieved = ring.bytes_per_env_step * n / per_launch_s / 1e9
    copy_gbs = None
    if rank == 0:
        copy_gbs = same_size_copy_gbs(ring.bytes_per_env_step * n // 2, ring_n, dev, ring.stream)
    barrier()

    # ---- strong scaling: 2^20 envs in TOTAL (north_star's target read literally): NUM_ENVS / world per
    # GPU through the same lockstep step, ring deep enough to evict the L2 between visits, and the same
    # shard stepped again and again (L2-resident: what a caller that keeps 131 072 envs per GPU sees)
    ns = NUM_ENVS // world
    strong = None
    if ns % 128 == 0:
        ring_s = ring if ns == n else StepRing(ns, ring_depth(ns, 117), local_rank, args.seed + 100, rank)
        ssteps = max(args.steps, 200) if not strong_headline else args.steps
        if ns == n and not strong_headline:
            s_ms, s_steps = elapsed_ms, args.steps
        else:
            s_ms, s_steps = time_ring(ring_s, ssteps, warmup_run, barrier, world, dist, dev), ssteps
        s_launch = s_ms * 1e-3 / s_steps
        s_copy = None
        if rank == 0:
            s_copy = copy_gbs if ns == n else same_size_copy_gbs(ring_s.bytes_per_env_step * ns // 2,
                                                                 ring_s.ring, dev, ring_s.stream)
        barrier()
        res = StepRing(ns, 1, local_rank, args.seed + 200, rank) if ns != n else None
        r_launch = None
        if res is not None:
            r_ms = time_ring(res, max(ssteps, 200), 20, barrier, world, dist, dev)
            r_launch = r_ms * 1e-3 / max(ssteps, 200)
            res.close()
        strong = {
            "workload": WORKLOAD_STRONG, "envs_total": ns * world, "envs_per_gpu": ns,
            "value": world * ns / s_launch, "unit": UNIT, "us_per_tick": s_launch * 1e6, "steps": s_steps,
            "ring_shards": ring_s.ring,
            "l2_policy": f"ring of {ring_s.ring} shards x {ring_s.bytes_per_env_step * ns / 1e6:.1f} MB per step "
                         f"(>= 2 x 126 MB L2): every step streams from HBM",
            "roofline": {"bound": "hbm", "achieved": ring_s.bytes_per_env_step * ns / s_launch / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": ring_s.bytes_per_env_step * ns / s_launch / 1e9 / peak,
                         "same_size_copy_gbs": s_copy,
                         "note": "a launch this small cannot reach the copy peak: same_size_copy_gbs is a plain "
                                 "torch D2D copy of the same bytes per launch under the same launch pattern"},
            "l2_resident": None if r_launch is None else {
                "value": world * ns / r_launch, "us_per_tick": r_launch * 1e6,
                "note": f"the same shard every step ({ring_s.bytes_per_env_step * ns / 1e6:.0f} MB working set, "
                        f"L2 is 126 MB): not an HBM figure unless the set exceeds the L2"},
        }
        if ring_s is not ring:
            ring_s.close()

    # ---- e2e: the public API with host arrays (page-locked), copies inside the timed region
    e = ring.envs[0]
    h_keys = e.pinned_empty((n, nk), np.uint8)
    h_mouse = e.pinned_empty((n,), np.float32)
    h_keys[...] = ring.keys[0].cpu().numpy()
    h_mouse[...] = ring.mouse[0].cpu().numpy()
    e2e_steps = max(3, min(args.e2e_steps, max(args.steps, 20)))
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
    ceiling = host_copy_ceiling(n, nk, dev, barrier, world, dist)
    barrier()

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
                       f"env_config, {n5} envs/GPU x {ticks5} ticks closed loop on the device "
                       f"({timing5.get('how', 'policy kernel + step kernel per tick')}), "
                       f"{rate5:.3e} env-steps/s incl. the policy (timed with CUDA events)")
    else:
        tracked = benv.VectorPhysEnv(workload_config(n5), device=local_rank, seed=args.seed,
                                     env_index_base=rank * n5, track_returns=True)
        tracked.rollout("strafe_jump", 722, policy_seed=1)
        rate5 = None
        policy_desc = f"scripted strafe_jump, 722 ticks, {n5} envs/GPU"
    red = sharding.reduce_metrics(tracked.metrics(), device=dev)
    zs_mean, zs_episodes = red["zero_start_total_reward_mean"], red["zero_start_episodes"]

    # ---- the policy kernel alone (k_actor, one launch per call): forward + sampling for 2^20 observations,
    # against the measured bf16 tensor peak (the MMAs it issues incl. the padding of layers 1 and 3)
    policy_step = None
    if os.path.exists(policy_path) and rank == 0:
        obs_p = torch.rand((NUM_ENVS, 6), device=dev) * 2
        out_p = (torch.empty((NUM_ENVS, pol.num_keys), dtype=torch.uint8, device=dev),
                 torch.empty(NUM_ENVS, device=dev))
        for _ in range(5):
            pol.act(obs_p, out=out_p)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        p0.record()
        for _ in range(reps):
            pol.act(obs_p, out=out_p)
        p1.record()
        torch.cuda.synchronize(dev)
        us = p0.elapsed_time(p1) / reps * 1e3
        issued = 2 * (32 + 256 + 16) * 256           # flop per env the tensor cores execute (K = 32 / 256 / 256)
        useful = 2 * (6 * 256 + 256 * 256 + 256 * 10)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sustained = peaks.get("bf16_tflops_sustained")
        policy_step = {"us_per_call": us, "envs": NUM_ENVS, "kernel": "k_actor<ACT>: tcgen05 MLP 6-256-256-10 + sampling",
                       "roofline": {"bound": "tensor", "achieved": NUM_ENVS * issued / us / 1e6,
                                    "peak": sustained, "unit": "TFLOP/s",
                                    "frac": (NUM_ENVS * issued / us / 1e6 / sustained) if sustained else None,
                                    "useful_tflops": NUM_ENVS * useful / us / 1e6,
                                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if sustained else "none",
                                    "note": "the tanh epilogues (XU), not the MMAs, bound this kernel: DESIGN.md 8.1"}}
        del obs_p, out_p
    barrier()

    # ---- BASELINE config 4: 2^20 envs sharded 131072 per GPU, scripted strafe-jump policy, 10k ticks
    # in the multi-tick in-register rollout kernel (no per-tick HBM traffic)
    n4, ticks4 = 1 << 17, 10000
    e4 = benv.VectorPhysEnv(dict(CONFIG_100M, num_envs=n4, zero_start_prob=1.0), device=local_rank,
                            seed=args.seed, env_index_base=rank * n4)
    e4.rollout("strafe_jump", 100)
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(ring.stream)
    e4.rollout("strafe_jump", ticks4)
    r1.record(ring.stream)
    barrier()
    t4 = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t4, op=dist.ReduceOp.MAX)
    rollout_value = world * n4 * ticks4 / (float(t4.item()) * 1e-3)

    if rank == 0:
        traffic, traffic_how = ncu_traffic()
        weak_roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_how": traffic_how,
                         "kernel": "k_step_tma", "bytes_per_env_step": ring.bytes_per_env_step,
                         "peak_source": peak_src, "same_size_copy_gbs": copy_gbs,
                         "same_size_copy_note": "torch D2D copy_ moving the same bytes per launch, same ring, "
                                                "timed live in this run: what any kernel launched at this size "
                                                "can reach"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "warmup_steps_run": warmup_run, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n, "ring_shards": ring_n,
                       "l2_policy": f"inputs larger than L2: ring of {ring_n} env shards x "
                                    f"{ring.bytes_per_env_step * n / 1e6:.0f} MB touched per step",
                       "state_bytes_per_env": ring.state_bytes, "key_timers": "f64" if ring.f64_stamps else "u8",
                       "launch": "one q1_step call (k_step_tma launch, programmatic dependent launch) per "
                                 "step on torch's current stream"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "cpu_affinity": numa,
                    "host_ceiling": ceiling, "frac_of_host_ceiling": e2e_value / ceiling,
                    "host_ceiling_how": "cudaMemcpyAsync of one step's results D2H and actions H2D between "
                                        "page-locked host memory and HBM on two streams, all ranks at once, no kernel",
                    "api": "VectorPhysEnv.vector_step((keys, mouse)) with page-locked NumPy arrays: q1_step_host "
                           "launches the step kernel on the mapped host buffers (actions read and results "
                           "written over PCIe inside the launch), then synchronises"},
            "gpu_launches": args.steps,
            "roofline": weak_roofline,
            "clocks": clocks.summary("sampled over the timed region and the same load continued for >= 100 ms"),
            "strong": strong,
            "config4_rollout": {"value": rollout_value, "unit": UNIT, "envs_per_gpu": n4, "ticks": ticks4,
                                "policy": "scripted strafe_jump generated on the device",
                                "kernel": "k_rollout: one launch, state in registers for all ticks"},
            "zero_start_total_reward_mean": {"value": zs_mean, "episodes": zs_episodes,
                                             "policy": policy_desc, "env_steps_per_s": rate5,
                                             "collective": "all_reduce(sum) of (sum, count)" if world > 1 else "none (1 GPU)"},
        }
        if policy_step is not None:
            line["policy_step"] = policy_step
        if strong_headline and strong is not None:
            line.update({"value": strong["value"], "ms_per_step": strong["us_per_tick"] * 1e-3,
                         "scaling": "strong", "steps": strong["steps"], "roofline": dict(
                             strong["roofline"], traffic=None, kernel="k_step_tma",
                             bytes_per_env_step=ring.bytes_per_env_step, peak_source=peak_src),
                         "weak": {"value": value, "ms_per_step": elapsed_ms / args.steps, "roofline": weak_roofline}})
            line["config"] = {"workload": WORKLOAD_STRONG, "envs_per_gpu": ns, "ring_shards": strong["ring_shards"],
                              "l2_policy": strong["l2_policy"], "state_bytes_per_env": ring.state_bytes,
                              "key_timers": "f64" if ring.f64_stamps else "u8"}
            line["gpu_launches"] = strong["steps"]
        if world == 1:
            line["small_batch"] = small_batch_latency(local_rank)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, dt, ticks = cpu_port_throughput(NUM_ENVS, threads, min_seconds=10.0)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": "port", "cpu": cpu_model(),
                "sample": f"{NUM_ENVS} envs x {ticks} ticks of the same Config and action "
                          f"distribution, C port of the reference NumPy path (oracle/q1_oracle.c), "
                          f"{threads} threads, {dt:.1f} s"}
            if not args.no_numpy_tiers:
                line["cpu_baseline"]["numpy"] = numpy_reference_tiers()
        print(json.dumps(line), flush=True)
    ring.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: 2^20 envs per GPU (default); strong: 2^20 envs in total, split over the GPUs")
    ap.add_argument("--envs", type=int, default=NUM_ENVS, help="envs per GPU (weak scaling)")
    ap.add_argument("--ring", type=int, default=0, help="env shards ticked round-robin (0: enough to evict L2)")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numpy-tiers", action="store_true")
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
