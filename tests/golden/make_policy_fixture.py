"""Extract the policy MLP of the reference's shipped checkpoint (data/checkpoints/wr/checkpoint, an
RLLib 0.8.4 pickle) into `wr_policy.npz`, and record known answers of a NumPy evaluation of it on the
UNMODIFIED reference env: the deterministic zero-start episode.

    python tests/golden/make_policy_fixture.py

ray / tensorflow are not installed, so their classes are stubbed while unpickling; only the
`default_policy/{fc_1,fc_2,fc_out}/{kernel,bias}` arrays are kept (the value branch is not needed
for rollouts).  Sampling rule restated from q1physrl/action_dist.py:67-76, 84-101, 151, 186-192.
"""
import io
import json
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402

CKPT = os.path.join(refshim.REFERENCE_ROOT, "data", "checkpoints", "wr")


class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__["state"] = state


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("ray", "tensorflow"):
            return type(name, (_Stub,), {})
        return super().find_class(module, name)


def load_weights():
    with open(os.path.join(CKPT, "checkpoint"), "rb") as f:
        d = _Unpickler(f).load()
    w = d["worker"]
    w = _Unpickler(io.BytesIO(w)).load() if isinstance(w, bytes) else w
    st = w["state"]["default_policy"]
    return {k.split("/", 1)[1].replace("/", "_"): np.asarray(v) for k, v in st.items()
            if k.split("/")[1] in ("fc_1", "fc_2", "fc_out")}


def mlp(w, obs):
    h = np.tanh(obs.astype(np.float32) @ w["fc_1_kernel"] + w["fc_1_bias"])
    h = np.tanh(h @ w["fc_2_kernel"] + w["fc_2_bias"])
    return h @ w["fc_out_kernel"] + w["fc_out_bias"]


def deterministic_action(logits, action_range):
    """Categorical argmax for the four keys; squash(mean) for the mouse (action_dist.py:84-88)."""
    from math import erf, sqrt
    keys = [int(logits[2 * k + 1] > logits[2 * k]) for k in range(4)]
    mean = float(np.clip(logits[8], -3, 3))
    cdf = 0.5 * (1 + erf(mean / (0.5 * 1.8137) / sqrt(2)))
    val = float(np.clip(cdf, 1e-6, 1 - 1e-6)) * (2 * action_range) - action_range
    return keys, val


def main():
    ref_env, _ = refshim.load()
    w = load_weights()
    with open(os.path.join(CKPT, "params.json")) as f:
        env_config = json.load(f)["env_config"]
    env_config["initial_yaw_range"] = tuple(env_config["initial_yaw_range"])
    cfg = ref_env.Config(**dict(env_config, num_envs=1, zero_start_prob=1.0))
    np.random.seed(0)
    e = ref_env.VectorPhysEnv(cfg)
    (o,) = e.vector_reset()
    total, ticks, max_speed, done = 0.0, 0, 0.0, False
    obs_log, logit_log, act_log = [], [], []
    with np.errstate(invalid="ignore", divide="ignore"):
        while not done:
            logits = mlp(w, o[None])[0]
            keys, val = deterministic_action(logits, float(cfg.action_range))
            obs_log.append(o.astype(np.float32)); logit_log.append(logits); act_log.append(keys + [val])
            (o,), (r,), (done,), _ = e.vector_step([tuple(keys + [np.array([val], np.float32)])])
            total += float(r)
            ticks += 1
            v = e.player_state.vel[0]
            max_speed = max(max_speed, float(np.hypot(v[0], v[1])))
    print(f"deterministic zero-start episode: {ticks} ticks, sum reward {total:.4f}, "
          f"max ground speed {max_speed:.2f}")
    np.savez_compressed(os.path.join(HERE, "wr_policy.npz"), env_config=json.dumps(env_config),
                        det_ticks=ticks, det_return=total, det_max_speed=max_speed,
                        det_obs=np.stack(obs_log), det_logits=np.stack(logit_log),
                        det_actions=np.array(act_log, np.float64), **w)
    print("wrote wr_policy.npz", os.path.getsize(os.path.join(HERE, "wr_policy.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
