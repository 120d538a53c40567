"""The shipped PPO policy on the GPU: the RLLib fully-connected model
`obs(6) -> tanh 256 -> tanh 256 -> 2 * num_keys + 2` (q1physrl checkpoints: default_policy/fc_1, fc_2,
fc_out) and sampling from `Q1PhysActionDist` (q1physrl/action_dist.py), so that rollouts of
`VectorPhysEnv` run closed-loop on the device with no host traffic.

The three GEMMs are plain library GEMMs (torch / cuBLAS); the distribution sampling is the
`k_sample_actions` kernel of libq1phys (`q1_sample_actions`), which writes the action arrays in the
layout `VectorPhysEnv.step_tensors` consumes.
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
        t = lambda a, dt: torch.as_tensor(np.asarray(a, np.float32)).to(self.device, dt).contiguous()
        self.w1, self.b1 = t(weights["fc_1_kernel"], self.dtype), t(weights["fc_1_bias"], self.dtype)
        self.w2, self.b2 = t(weights["fc_2_kernel"], self.dtype), t(weights["fc_2_bias"], self.dtype)
        self.w3, self.b3 = t(weights["fc_out_kernel"], self.dtype), t(weights["fc_out_bias"], self.dtype)
        assert self.w3.shape[1] == 2 * self.num_keys + 2, "policy head does not match the action space"

    @classmethod
    def from_npz(cls, path, **kwargs):
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
        _lib.check(_lib.load().q1_sample_actions(
            self.device.index, n, self.num_keys, ctypes.c_void_p(lg.data_ptr()), self.low, self.high,
            int(bool(deterministic)), self.seed, self.step_count, int(env_index_base),
            ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(mouse.data_ptr()), stream))
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


def rollout(env, policy, ticks, deterministic=False):
    """Closed loop on the device: policy -> `step_tensors` (fused auto-reset) for `ticks` ticks.
    Returns the last (obs, reward, done, zero_start) tensors; with `track_returns` the episode
    statistics accumulate in `env.metrics()`."""
    torch = policy._torch
    obs = torch.as_tensor(env._get_obs()).to(policy.device)
    base = env.info.env_index_base
    act_buf = step_buf = None
    out = None
    for _ in range(int(ticks)):
        act_buf = policy.act(obs, deterministic=deterministic, env_index_base=base, out=act_buf)
        out = env.step_tensors(act_buf[0], act_buf[1], auto_reset=True, out=step_buf)
        step_buf = out
        obs = out[0]
    return out
