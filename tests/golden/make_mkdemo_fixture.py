"""Golden fixture for the real-game adapters: run the reference's OWN `_make_observation` /
`_apply_action` (q1physrl/mkdemo.py:39-55, taken from the mounted checkout at generation time by
parsing the file -- the module itself imports ray / pyquake / cv2, none of which exist here)
against a scripted fake client, and record what they return / send.

    python tests/golden/make_mkdemo_fixture.py
"""
import ast
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402

ref_env, _ = refshim.load()
src = open(os.path.join(refshim.REFERENCE_ROOT, "q1physrl", "mkdemo.py")).read()
tree = ast.parse(src)
wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("_make_observation", "_apply_action")]
assert len(wanted) == 2
ns = {"np": np, "env": ref_env}
exec(compile(ast.Module(body=wanted, type_ignores=[]), "mkdemo.py", "exec"), ns)


class FakeClient:
    def __init__(self):
        self.angles, self.velocity, self.player_origin = (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)
        self.moves = []

    def move(self, **kw):
        self.moves.append(kw)


def main():
    cfg_dict = dict(num_envs=1, action_range=10, allow_jump=True, allow_yaw=True, auto_jump=False,
                    discrete_yaw_steps=-1, fmove_max=800, hover=False, initial_yaw_range=[0, 360],
                    key_press_delay=0.3, max_initial_speed=700, smooth_keys=True, smove_max=1060,
                    speed_reward=False, time_delta=0.013888888888888, time_limit=10, zero_start_prob=0.01)
    out = {"config": json.dumps(cfg_dict)}
    for tag, auto_jump in (("keys", False), ("autojump", True)):
        cfg = ref_env.Config(**dict(cfg_dict, auto_jump=auto_jump))
        nk = 3 if auto_jump else 4
        dec = ref_env.ActionDecoder(cfg)
        dec.vector_reset(np.array([ref_env.INITIAL_YAW_ZERO]))
        rng = np.random.default_rng(11 + auto_jump)
        client = FakeClient()
        T = 800
        t = 0.0
        rec = dict(time_remaining=[], angles=[], velocity=[], origin=[], keys=[], mouse=[], obs=[])
        for k in range(T):
            t += 1 / 72 + rng.uniform(-1e-3, 1e-3)          # the server's frame times jitter
            time_remaining = cfg.time_limit - t
            client.angles = (0.0, float(rng.uniform(-np.pi, np.pi)), 0.0)
            client.velocity = tuple(rng.normal(0, 200, 3).astype(np.float32).astype(np.float64))
            client.player_origin = (float(rng.uniform(-100, 100)), float(rng.uniform(0, 5000)), float(rng.uniform(24, 80)))
            obs = ns["_make_observation"](client, time_remaining, cfg)
            keys = rng.integers(0, 2, nk)
            mouse = np.float32(rng.uniform(-10, 10))
            action = tuple([np.array([k_]) for k_ in keys] + [np.array([mouse])])
            ns["_apply_action"](client, dec, action, time_remaining)
            rec["time_remaining"].append(time_remaining)
            rec["angles"].append(client.angles)
            rec["velocity"].append(client.velocity)
            rec["origin"].append(client.player_origin)
            rec["keys"].append(keys)
            rec["mouse"].append(mouse)
            rec["obs"].append(obs)
        for name, v in rec.items():
            out[f"{tag}_{name}"] = np.array(v)
        out[f"{tag}_move_yaw"] = np.array([float(m["yaw"]) for m in client.moves])
        out[f"{tag}_move_forward"] = np.array([int(m["forward"]) for m in client.moves])
        out[f"{tag}_move_side"] = np.array([int(m["side"]) for m in client.moves])
        out[f"{tag}_move_buttons"] = np.array([int(m["buttons"]) for m in client.moves])
        out[f"{tag}_final_last_press"] = np.array(dec._last_key_press_time, np.float64)
        assert all(m["pitch"] == 0 and m["roll"] == 0 and m["up"] == 0 and m["impulse"] == 0 for m in client.moves)
    np.savez_compressed(os.path.join(HERE, "mkdemo_adapters.npz"), **out)
    print("wrote mkdemo_adapters.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
