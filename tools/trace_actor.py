"""Timeline of one CTA of k_actor from a -DQ1_ACTOR_TRACE=1 build (Q1PHYS_LIB=build/libq1phys_trace.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from q1physrl_b200 import _lib, env as benv, policy as bpolicy
mode = sys.argv[1] if len(sys.argv) > 1 else "act"
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "wr_policy.npz")
pol, env_config = bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
if mode == "act":
    n = 148 * 128 * 10
    obs = torch.rand((n, 6), device="cuda") * 2
    out = (torch.empty((n, 4), dtype=torch.uint8, device="cuda"), torch.empty(n, device="cuda"))
    for _ in range(3):
        pol.act(obs, out=out)
else:
    cfg = dict(env_config, initial_yaw_range=tuple(env_config["initial_yaw_range"]), num_envs=148 * 2 * 128)
    e = benv.VectorPhysEnv(cfg, seed=2, track_returns=True)
    for _ in range(2):
        pol.rollout_fused(e, 8, want_outputs=False)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.library_path())
t = np.zeros((16, 64), np.int64)
assert lib.q1_actor_trace(ctypes.c_void_p(t.ctypes.data)) == 0
names = {0: "mma:top", 1: "mma:X", 2: "mma:L1 issued", 11: "", 12: "mma:L2q1 issued", 13: "mma:L2q2", 14: "mma:L2q3",
         15: "mma:M3q0", 16: "mma:M3q1", 17: "mma:M3q2", 18: "mma:M3q3", 52: "env:top", 53: "env:D3", 54: "env:E arrived",
         55: "env:acted", 56: "env:prepared"}
for i in range(2):
    names[3 + i] = f"mma:L2q0 step{i}"
for p in range(4):
    for k, nm in enumerate(("top", "L1 seen", "H1 step0", "H1 step1", "L2 seen", "H2 done")):
        names[20 + 8 * p + k] = f"epi{p}:{nm}"
t0 = t[2, 0]
for s in range(2, 6):
    ev = sorted((int(t[s, e]), e) for e in range(64) if t[s, e] and names.get(e))
    print(f"--- sequence {s} (cycles since sequence 2 top)")
    for c, e in ev:
        print(f"  {c - t0:8d}  {names[e]}")
