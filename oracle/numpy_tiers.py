"""Time the UNMODIFIED reference NumPy path on host cores (tiers T1 / T2 / T3).

MEASUREMENT INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg).  The reference is imported through
oracle/refshim.py from the checkout or from the byte-for-byte staging `oracle/_ref/`:

  T1  VectorPhysEnv.vector_step fed RLLib's list of per-env action tuples -- the true public path
      (env.py:482-510 incl. ActionDecoder._fix_actions, env.py:221-223)
  T2  the same call fed a pre-built (N, nk+1) array with _fix_actions made the identity: the env
      arithmetic alone
  T3  phys.apply alone (phys.py:184-197)

NumPy runs these ufuncs on one thread, so P worker PROCESSES each own a slice of the envs and the
aggregate is reported: measure(tier, envs_per_proc, procs, seconds).  Workers start their timed loops
at a common wall-clock instant and run whole ticks until `seconds` have passed.
"""
import argparse
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(tier, n, seconds, start_at, cfg_json):
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import refshim
    ref_env, ref_phys = refshim.load()
    cfg = json.loads(cfg_json)
    cfg["initial_yaw_range"] = tuple(cfg["initial_yaw_range"])
    cfg["num_envs"] = n
    np.random.seed(os.getpid() & 0xFFFF)
    rng = np.random.default_rng(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = ref_env.VectorPhysEnv(ref_env.Config(**cfg))
        nk = e._action_decoder._num_keys
        keys = rng.integers(0, 2, (n, nk))
        mouse = rng.uniform(-10, 10, n).astype(np.float32)
        if tier == "T1":
            actions = [tuple([int(k) for k in keys[i]] + [np.array([mouse[i]], np.float32)]) for i in range(n)]
            step = lambda: e.vector_step(actions)
        elif tier == "T2":
            actions = np.concatenate([keys.astype(np.float64), mouse[:, None].astype(np.float64)], axis=1)
            e._action_decoder._fix_actions = lambda a: a
            step = lambda: e.vector_step(actions)
        else:
            inputs = ref_phys.Inputs(yaw=rng.uniform(0, 360, n), pitch=np.zeros(n), roll=np.zeros(n),
                                     fmove=np.full(n, 800), smove=rng.choice([-1060, 0, 1060], n),
                                     button2=rng.random(n) < 0.5, time_delta=np.full(n, cfg["time_delta"]))
            state = [e.player_state]

            def step():
                state[0] = ref_phys.apply(inputs, state[0])
        step()                                                  # warm-up tick
        while time.time() < start_at:
            time.sleep(0.001)
        t0 = time.perf_counter()
        ticks = 0
        while True:
            step()
            ticks += 1
            dt = time.perf_counter() - t0
            if dt >= seconds:
                break
    print(json.dumps({"env_steps": n * ticks, "seconds": dt, "ticks": ticks}), flush=True)


def measure(tier, envs_per_proc, procs, seconds, config):
    """-> dict(value env-steps/s aggregate, procs, envs_per_proc, ticks_per_proc, seconds)."""
    lead = 2.0 + 0.0025 * envs_per_proc / 1000 * (8 if tier == "T1" else 1)   # imports + warm-up tick
    start_at = time.time() + lead
    cmd = [sys.executable, "-m", "oracle.numpy_tiers", "--worker", "--tier", tier, "--envs", str(envs_per_proc),
           "--seconds", str(seconds), "--start-at", repr(start_at), "--config", json.dumps(config, default=float)]
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    ps = [subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
          for _ in range(procs)]
    outs = []
    for p in ps:
        out, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"numpy_tiers worker failed: {err[-400:]}")
        outs.append(json.loads(out.strip().splitlines()[-1]))
    total = sum(o["env_steps"] for o in outs)
    longest = max(o["seconds"] for o in outs)
    return {"value": total / longest, "procs": procs, "envs_per_proc": envs_per_proc,
            "ticks_per_proc": min(o["ticks"] for o in outs), "seconds": longest}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--tier", default="T2")
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--start-at", type=float, default=0.0)
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--config", default="")
    a = ap.parse_args()
    if a.worker:
        _worker(a.tier, a.envs, a.seconds, a.start_at, a.config)
    else:
        sys.path.insert(0, ROOT)
        import bench
        print(measure(a.tier, a.envs, a.procs, a.seconds, bench.workload_config(a.envs)))
