"""Timeline of one CTA of k_actor from a -DQ1_ACTOR_TRACE=1 build (Q1PHYS_LIB=build/libq1phys_trace.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from q1physrl_b200 import _lib, env as benv, policy as bpolicy
mode = sys.argv[1] if len(sys.argv) > 1 else "act"
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "wr_policy.npz")
pol, env_config = bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
if mode == "act":
    n = 148 * 128 * (int(sys.argv[2]) if len(sys.argv) > 2 else 10)   # tiles per SM
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
if mode == "act":
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        pol.act(obs, out=out)
    e1.record()
    torch.cuda.synchronize()
    print(f"{n} envs: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (this build stamps clocks: slower than the shipped one)")
lib = ctypes.CDLL(_lib.library_path())
t = np.zeros((16, 64), np.int64)
assert lib.q1_actor_trace(ctypes.c_void_p(t.ctypes.data)) == 0
names = {0: "mma:top", 1: "mma:X", 2: "mma:L1 issued", 11: "mma:L2q0 issued", 12: "mma:L2q1 issued", 13: "mma:L2q2 issued", 14: "mma:L2q3 issued",
         15: "mma:M3q0", 16: "mma:M3q1", 17: "mma:M3q2", 18: "mma:M3q3", 52: "env:top", 53: "env:D3", 54: "env:E arrived",
         55: "env:acted", 56: "env:prepared", 57: "env:sampled", 58: "env:ticked"}
for q in range(4):
    names[5 + q] = f"mma:H2 quarter {q} seen"
names[9], names[10] = "mma:quarters 0, 1 read (or last tile)", "mma:quarter 3 read"
for i in range(2):
    names[3 + i] = f"mma:H1 half {i} seen"
for p in range(4):
    for k, nm in enumerate(("top", "L1 seen", "H1 step0", "H1 step1", "L2 seen", "H2 done")):
        names[20 + 8 * p + k] = f"epi{p}:{nm}"
t0 = t[2, 0]
first = 2 if mode == 'act' else 6
t0 = t[first, 0]
for s in range(first, first + 2):
    ev = sorted((int(t[s, e]), e) for e in range(64) if t[s, e] and names.get(e))
    print(f"--- sequence {s} (cycles since sequence 2 top)")
    for c, e in ev:
        print(f"  {c - t0:8d}  {names[e]}")
tops = [int(t[s, 0]) for s in range(16) if t[s, 0]]
if tops:
    print("tile periods (mma:top to mma:top):", [b - a for a, b in zip(tops, tops[1:])])
if t[0, 62] and t[0, 63]:
    print(f"kernel entry -> first recorded mma:top {tops[0] - int(t[0, 62]) if tops else None} cycles; entry -> exit {int(t[0, 63] - t[0, 62])} cycles"
          f" = {int(t[0, 61] - t[0, 60])} ns: the SM clock ran at {(t[0, 63] - t[0, 62]) / max(1, t[0, 61] - t[0, 60]) * 1e3:.0f} MHz")
if hasattr(lib, "q1_actor_block_cycles"):
    b = np.zeros(256, np.int64)
    assert lib.q1_actor_block_cycles(ctypes.c_void_p(b.ctypes.data)) == 0
    b = b[b != 0]
    cyc, sm = b >> 8, b & 255
    order = np.argsort(cyc)
    print(f"entry -> exit per CTA: min {cyc.min()} median {int(np.median(cyc))} max {cyc.max()} cycles over {len(b)} CTAs; "
          f"slowest on SMs {sm[order[-6:]].tolist()}, fastest on SMs {sm[order[:6]].tolist()}")
    print("sorted cycles/1000:", np.round(np.sort(cyc) / 1000).astype(int).tolist())
