"""The shipped PPO policy on the GPU: the RLLib fully-connected model
`obs(6) -> tanh 256 -> tanh 256 -> 2 * num_keys + 2` (q1physrl checkpoints: default_policy/fc_1, fc_2,
fc_out) and sampling from `Q1PhysActionDist` (q1physrl/action_dist.py), so that rollouts of
`VectorPhysEnv` run closed-loop on the device with no host traffic.

`MLPPolicy`: three plain library GEMMs (torch / cuBLAS, the fp32 parity reference) + the
`k_sample_actions` kernel.  `FusedMLPPolicy`: the hand-written tcgen05 kernel `k_actor`
(csrc/q1_actor.cu), either per call (`act`) or as the whole closed loop in one launch
(`rollout_fused`, `record_episode`).
"""
import ctypes
import json

import numpy as np

from . import _lib


class MLPPolicy:
    def __init__(self, weights, num_keys=4, action_range=10.0, device=0, seed=0, dtype=None):
        import torch
        self._torch = torch
        self.device = torch.device("cuda", device)
        self.dtype = dtype or torch.float32
        self.num_keys = int(num_keys)
        self.low, self.high = -float(action_range), float(action_range)
        self.seed = int(seed)
        self.step_count = 0
        self._step_dev = None  # device-side step counter, used while a CUDA graph is replayed
        t = lambda a, dt: torch.as_tensor(np.asarray(a, np.float32)).to(self.device, dt).contiguous()
        self.w1, self.b1 = t(weights["fc_1_kernel"], self.dtype), t(weights["fc_1_bias"], self.dtype)
        self.w2, self.b2 = t(weights["fc_2_kernel"], self.dtype), t(weights["fc_2_bias"], self.dtype)
        self.w3, self.b3 = t(weights["fc_out_kernel"], self.dtype), t(weights["fc_out_bias"], self.dtype)
        assert self.w3.shape[1] == 2 * self.num_keys + 2, "policy head does not match the action space"

    @classmethod
    def from_npz(cls, path, **kwargs):
        """-> (policy, env_config dict stored beside the weights)."""
        with np.load(path) as z:
            weights = {k: z[k] for k in z.files if k.startswith("fc_")}
            cfg = json.loads(str(z["env_config"])) if "env_config" in z.files else {}
        kwargs.setdefault("action_range", cfg.get("action_range", 10.0))
        has_jump = not cfg.get("auto_jump", False) and cfg.get("allow_jump", True)
        kwargs.setdefault("num_keys", 4 if has_jump else 3)
        return cls(weights, **kwargs), cfg

    def logits(self, obs):
        """obs (N, 6) CUDA tensor -> (N, 2 * num_keys + 2) float32 policy outputs."""
        torch = self._torch
        x = obs.to(self.dtype)
        h = torch.addmm(self.b1, x, self.w1).tanh_()
        h = torch.addmm(self.b2, h, self.w2).tanh_()
        return torch.addmm(self.b3, h, self.w3).float().contiguous()

    def act(self, obs, deterministic=False, env_index_base=0, out=None):
        """obs (N, 6) CUDA tensor -> (keys uint8 (N, num_keys), mouse float32 (N,)) CUDA tensors."""
        torch = self._torch
        lg = self.logits(obs)
        n = lg.shape[0]
        if out is None:
            out = (torch.empty((n, self.num_keys), dtype=torch.uint8, device=self.device),
                   torch.empty(n, dtype=torch.float32, device=self.device))
        keys, mouse = out
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        step_ptr = ctypes.c_void_p(self._step_dev.data_ptr()) if self._step_dev is not None else None
        _lib.check(_lib.load().q1_sample_actions(
            self.device.index, n, self.num_keys, ctypes.c_void_p(lg.data_ptr()), self.low, self.high,
            int(bool(deterministic)), self.seed, self.step_count, step_ptr, int(env_index_base),
            ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(mouse.data_ptr()), stream))
        if self._step_dev is not None:
            self._step_dev += 1
        self.step_count += 1
        return keys, mouse

    def compute_action(self, obs):
        """`trainer.compute_action(obs)` stand-in for analyse.eval_sim: one observation in, the
        deterministic action tuple out."""
        torch = self._torch
        o = torch.as_tensor(np.asarray(obs, np.float32)[None]).to(self.device)
        keys, mouse = self.act(o, deterministic=True)
        k = keys.cpu().numpy()[0]
        return tuple(int(x) for x in k) + (mouse.cpu().numpy().astype(np.float32),)


class FusedMLPPolicy(MLPPolicy):
    """The same policy as ONE warp-specialised sm_100a kernel (`k_actor`, csrc/q1_actor.cu): all three
    layers on the tensor cores (tcgen05.mma, bf16 operands -- layer 1 with bf16-split operands, exact
    to ~2^-17 -- fp32 accumulators in TMEM), tanh epilogues overlapped with the MMAs, sampling and
    (closed loop) the env tick on their own warps; the hidden activations never leave tensor memory."""

    def __init__(self, weights, num_keys=4, action_range=10.0, device=0, seed=0):
        super().__init__(weights, num_keys=num_keys, action_range=action_range, device=device, seed=seed)
        f32 = lambda k: np.ascontiguousarray(weights[k], dtype=np.float32)
        arrs = [f32(k) for k in ("fc_1_kernel", "fc_1_bias", "fc_2_kernel", "fc_2_bias",
                                 "fc_out_kernel", "fc_out_bias")]
        assert arrs[0].shape == (6, 256) and arrs[2].shape == (256, 256)
        self._handle = ctypes.c_void_p()
        _lib.check(_lib.load().q1_policy_create(
            int(device), self.num_keys, *[ctypes.c_void_p(a.ctypes.data) for a in arrs],
            ctypes.byref(self._handle)))

    def close(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            _lib.load().q1_policy_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: the library module may already be gone
            pass

    def _launch(self, obs, deterministic, env_index_base, out, logits_out):
        torch = self._torch
        obs = obs.to(torch.float32).contiguous()
        n = obs.shape[0]
        if out is None:
            out = (torch.empty((n, self.num_keys), dtype=torch.uint8, device=self.device),
                   torch.empty(n, dtype=torch.float32, device=self.device))
        keys, mouse = out
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        step_ptr = ctypes.c_void_p(self._step_dev.data_ptr()) if self._step_dev is not None else None
        _lib.check(_lib.load().q1_policy_act(
            self._handle, n, ctypes.c_void_p(obs.data_ptr()), self.low, self.high,
            int(bool(deterministic)), self.seed, self.step_count, step_ptr, int(env_index_base),
            ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(mouse.data_ptr()),
            ctypes.c_void_p(logits_out.data_ptr()) if logits_out is not None else None, stream))
        return out

    def logits(self, obs):
        torch = self._torch
        lg = torch.empty((obs.shape[0], 2 * self.num_keys + 2), dtype=torch.float32, device=self.device)
        self._launch(obs, True, 0, None, lg)
        return lg

    def act(self, obs, deterministic=False, env_index_base=0, out=None):
        out = self._launch(obs, deterministic, env_index_base, out, None)
        if self._step_dev is not None:
            self._step_dev += 1
        self.step_count += 1
        return out

    def check(self):
        """Synchronise and raise if a policy kernel's watchdog fired (`q1_policy_check`)."""
        _lib.check(_lib.load().q1_policy_check(self._handle))

    # ------------------------------------------------------------------ the closed loop, fused
    @property
    def device_policy(self):
        """`analyse.eval_sim` runs the episode closed loop on the device for trainers that have one."""
        return self

    def can_fuse(self, env):
        """The fused closed-loop kernel needs the counter form of the key timers."""
        return not env.info.f64_stamps and env._num_keys == self.num_keys

    def rollout_fused(self, env, ticks, deterministic=False, auto_reset=True, want_outputs=True):
        """`ticks` ticks of policy -> sample -> env tick -> observation in ONE launch
        (`q1_policy_rollout`: env state in shared memory, activations in tensor memory, no launch and
        no HBM traffic per tick), asynchronous on torch's current stream.  The sampling noise is keyed
        by (policy seed, global env index, env tick counter).  -> (final obs (N, 6), per-env reward
        sum (N,)) CUDA tensors, or None."""
        torch = self._torch
        n = env.num_envs
        obs = rsum = None
        if want_outputs:
            obs = torch.empty((n, 6), dtype=torch.float32, device=self.device)
            rsum = torch.empty(n, dtype=torch.float32, device=self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.load().q1_policy_rollout(
            self._handle, env.handle, int(ticks), int(bool(auto_reset)), int(bool(deterministic)), self.seed,
            self.low, self.high, 0, None,
            ctypes.c_void_p(obs.data_ptr()) if obs is not None else None,
            ctypes.c_void_p(rsum.data_ptr()) if rsum is not None else None, stream))
        env._step_num += int(ticks)
        return (obs, rsum) if want_outputs else None

    def record_episode(self, env, ticks, deterministic=True, auto_reset=False, shadow_jump=True):
        """The fused closed loop with the per-tick record of `VectorPhysEnv.record`
        (`q1_policy_rollout_host`) -> dict of arrays (ticks, num_envs, ...) + "final_obs"."""
        tape = env.new_record(int(ticks))
        final_obs = np.empty((env.num_envs, 6), np.float32)
        view = env._record_view(tape, 0)
        _lib.check(_lib.load().q1_policy_rollout_host(
            self._handle, env.handle, int(ticks), int(bool(auto_reset)), int(bool(deterministic)), self.seed,
            self.low, self.high, _lib.Q1_RECORD_SHADOW_JUMP if shadow_jump else 0, ctypes.byref(view),
            ctypes.c_void_p(final_obs.ctypes.data)))
        env._step_num += int(ticks)
        tape["final_obs"] = final_obs
        return tape


def rollout(env, policy, ticks, deterministic=False, graph=True, ticks_per_graph=8, timing=None, fused=None):
    """Closed loop on the device for `ticks` ticks with fused auto-reset; with `track_returns` the
    episode statistics accumulate in `env.metrics()`.

    fused (default: whenever possible): a `FusedMLPPolicy` on an env with counter key timers runs the
    whole loop as ONE kernel launch (`FusedMLPPolicy.rollout_fused`) and returns (final obs, per-env
    reward sum).  Otherwise: policy kernel -> `step_tensors` per tick, returning the last (obs, reward,
    done, zero_start) tensors:

    graph=True captures `ticks_per_graph` ticks (policy kernel, noise counter, step kernel each) in
    one CUDA graph and replays it: the loop is launch-bound otherwise.  The sampling noise advances
    through a device counter, so replays draw fresh actions.  `timing`: a dict that receives
    `ticks` and `seconds` of the replayed part, timed with CUDA events (capture excluded)."""
    torch = policy._torch
    ticks = int(ticks)
    dev = policy.device
    if fused is None:
        fused = hasattr(policy, "rollout_fused") and policy.can_fuse(env)
    if fused:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = policy.rollout_fused(env, ticks, deterministic=deterministic)
        e1.record()
        if timing is not None:
            torch.cuda.current_stream(dev).synchronize()
            timing.update(ticks=ticks, seconds=e0.elapsed_time(e1) * 1e-3,
                          how="ONE launch of the fused policy + env kernel k_actor: env state in shared "
                              "memory, activations in tensor memory")
        return out
    base = env.info.env_index_base
    obs = torch.as_tensor(env._get_obs()).to(dev)
    n = obs.shape[0]
    act_buf = (torch.empty((n, policy.num_keys), dtype=torch.uint8, device=dev),
               torch.empty(n, dtype=torch.float32, device=dev))
    step_buf = (obs, torch.empty(n, dtype=torch.float32, device=dev),
                torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))

    def tick():
        policy.act(step_buf[0], deterministic=deterministic, env_index_base=base, out=act_buf)
        env.step_tensors(act_buf[0], act_buf[1], auto_reset=True, out=step_buf)

    if not graph or ticks < 8:
        for _ in range(ticks):
            tick()
        return step_buf
    policy._step_dev = torch.full((1,), policy.step_count, dtype=torch.int64, device=dev)
    try:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):          # warm up allocators / cuBLAS workspaces outside the capture
                tick()
        torch.cuda.current_stream(dev).wait_stream(side)
        per = max(1, min(int(ticks_per_graph), ticks - 3))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(per):
                tick()
        replays = (ticks - 3) // per      # the capture itself executes nothing
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(replays):
            g.replay()
        e1.record()
        for _ in range(ticks - 3 - replays * per):
            tick()
        torch.cuda.current_stream(dev).synchronize()
        # the library (and this wrapper) counted the captured ticks once, when they were captured;
        # the capture executed nothing and the replays executed replays * per ticks
        _lib.check(_lib.load().q1_advance_ticks(env.handle, (replays - 1) * per))
        env._step_num += (replays - 1) * per
        if timing is not None:
            timing.update(ticks=replays * per, seconds=e0.elapsed_time(e1) * 1e-3,
                          how=f"policy kernel + step kernel per tick, {per} ticks per CUDA graph")
    finally:
        policy.step_count = int(policy._step_dev.item())
        policy._step_dev = None
    return step_buf
