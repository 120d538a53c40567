"""GPU: Q1PhysActionDist sampling on the device (csrc/q1_sample.cuh, q1physrl/action_dist.py:67-76, 84-101,
151, 186-243) against the analytic distributions: squashed-Gaussian CDF vs scipy.stats.norm on a grid,
Categorical key frequencies down to p = 1e-5 (the key draw has 24 bits), determinism of the stream."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCALE = 0.5 * 1.8137          # action_dist.py:151 (the reference's std of the squashing CDF's argument)


def _sample(logits, num_keys=4, low=-10.0, high=10.0, deterministic=False, seed=7, step=3, base=0):
    import torch
    from q1physrl_b200 import _lib
    lg = torch.as_tensor(np.ascontiguousarray(logits, np.float32)).cuda()
    n = lg.shape[0]
    keys = torch.empty((n, num_keys), dtype=torch.uint8, device="cuda")
    mouse = torch.empty(n, dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().q1_sample_actions(0, n, num_keys, ctypes.c_void_p(lg.data_ptr()), low, high,
                                             int(deterministic), seed, step, None, base,
                                             ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(mouse.data_ptr()),
                                             None))
    torch.cuda.synchronize()
    return keys.cpu().numpy(), mouse.cpu().numpy()


def test_deterministic_squash_equals_the_normal_cdf_on_a_grid():
    """deterministic action = clip(Phi(clip(mean, -3, 3) / 0.90685), 1e-6, 1 - 1e-6) * (high - low) + low
    (action_dist.py:84-88, 186-192).  Budget: 4 float32 ulp of the +-10 range."""
    from scipy.stats import norm
    mean = np.linspace(-4, 4, 20001)
    lg = np.zeros((mean.size, 10), np.float32)
    lg[:, 8] = mean
    lg[:, 1::2][:, :4] = np.array([1, -1, 0.5, -0.5])           # key logits: argmax -> 1, 0, 1, 0
    keys, mouse = _sample(lg, deterministic=True)
    assert (keys == np.array([1, 0, 1, 0])).all()
    m32 = lg[:, 8].astype(np.float64)
    want = np.clip(norm.cdf(np.clip(m32, -3, 3) / np.float64(np.float32(SCALE))), 1e-6, 1 - 1e-6) * 20 - 10
    err = np.abs(mouse - want)
    print("squash vs scipy: max abs err", err.max())
    assert err.max() <= 4 * np.spacing(np.float32(10))          # 4 ulp at |x| ~ 10 = 3.8e-6
    assert mouse.min() >= -10 and mouse.max() <= 10


def test_stochastic_mouse_action_follows_the_squashed_gaussian():
    """raw ~ N(mean, exp(log_std)); action = squash(raw): Kolmogorov-Smirnov against the analytic CDF
    F(a) = Phi((0.90685 * Phi^-1((a - low) / (high - low)) - mean) / std), at three (mean, log_std)."""
    from scipy.stats import kstest, norm
    n = 400000
    for mean, log_std in ((0.3, -0.5), (-1.2, 0.0), (2.5, -2.0)):
        lg = np.zeros((n, 10), np.float32)
        lg[:, 8], lg[:, 9] = mean, log_std
        _, mouse = _sample(lg, seed=11, step=5)
        std = np.exp(np.float32(log_std))

        def cdf(a):
            u = np.clip((a + 10) / 20, 1e-9, 1 - 1e-9)
            return norm.cdf((SCALE * norm.ppf(u) - mean) / std)
        stat, pvalue = kstest(mouse.astype(np.float64), cdf)
        print(f"mean {mean} log_std {log_std}: KS {stat:.5f} p {pvalue:.3f}")
        assert stat < 0.004                                  # n = 4e5: 1.36 / sqrt(n) = 0.0022 at 5 %
    # a wide Gaussian piles probability onto the clip (action_dist.py:188: cdf clipped to [1e-6, 1 - 1e-6]):
    # the atom at the low end must carry P(raw < 0.90685 * Phi^-1(1e-6))
    lg = np.zeros((n, 10), np.float32)
    lg[:, 8], lg[:, 9] = -1.2, 0.4
    _, mouse = _sample(lg, seed=12, step=1)
    low_atom = np.float32(np.float32(1e-6) * np.float32(20) + np.float32(-10))
    frac = np.mean(mouse == low_atom)
    want = norm.cdf((SCALE * norm.ppf(1e-6) + 1.2) / np.exp(np.float32(0.4)))
    print(f"clip atom: {frac:.5f} of the samples, expected {want:.5f}")
    assert mouse.min() == low_atom and abs(frac - want) < 5 * np.sqrt(want / n)


def test_key_frequencies_down_to_rare_events():
    """Categorical(2) per key: P(key) = sigmoid(l1 - l0).  The draw is a 24-bit uniform, so p = 1e-5 fires
    at its rate (round 1's 16-bit draw could not go below 1.5e-5 at all)."""
    n = 1 << 23
    probs = np.array([1e-5, 1e-3, 0.3, 0.999])
    lg = np.zeros((1, 10), np.float32)
    lg[0, 1:8:2] = np.log(probs / (1 - probs))
    lg = np.broadcast_to(lg, (n, 10))
    hits = np.zeros(4)
    for step in range(4):
        keys, _ = _sample(lg, seed=5, step=step)
        hits += keys.sum(axis=0)
    total = 4 * n
    for p, h in zip(probs, hits):
        sigma = np.sqrt(total * p * (1 - p))
        print(f"p = {p}: {int(h)} hits, expected {total * p:.1f} +- {sigma:.1f}")
        assert abs(h - total * p) < 5 * sigma + 1


def test_noise_stream_is_a_function_of_seed_env_and_step():
    lg = np.random.default_rng(0).normal(0, 1, (4096, 10)).astype(np.float32)
    k0, m0 = _sample(lg, seed=1, step=2, base=100)
    k1, m1 = _sample(lg, seed=1, step=2, base=100)
    assert np.array_equal(k0, k1) and np.array_equal(m0, m1)
    k2, m2 = _sample(lg[50:], seed=1, step=2, base=150)          # sharding: same global index, same draw
    assert np.array_equal(k0[50:], k2) and np.array_equal(m0[50:], m2)
    for kw in (dict(seed=2, step=2, base=100), dict(seed=1, step=3, base=100)):
        k3, m3 = _sample(lg, **kw)
        assert not np.array_equal(m0, m3)
    k4, m4 = _sample(lg[:, :8], num_keys=3, seed=1, step=2, base=100)
    assert k4.shape == (4096, 3) and np.isfinite(m4).all()
