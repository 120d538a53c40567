"""Coefficients of the odd polynomial tanh(x) ~ x * P(x^2) on |x| <= C that the fused policy kernel evaluates on
the FMA pipe for a share of the hidden units (q1_actor.cu, tanh_poly2), next to MUFU.TANH for the rest.

Minimax in RELATIVE error by linear programming on a dense grid (the error of an odd polynomial times x is
relative by construction near 0), with the end point constrained from below so that the clamped tail
x >= C rounds to 1.0 in bfloat16 exactly as tanh itself does there.  Prints the C array and the achieved
errors of the float32 Horner evaluation, before and after the bfloat16 rounding the kernel applies."""
import numpy as np
from scipy.optimize import linprog

C, K = 3.5, 8


def fit(c=C, k=K, n=6000):
    x = np.linspace(1e-4, c, n)
    t = x * x
    y = np.tanh(x)
    V = np.stack([x * t ** j for j in range(k + 1)], 1)
    A = np.vstack([np.hstack([V, -y[:, None]]), np.hstack([-V, -y[:, None]])])
    b = np.hstack([y, -y])
    # end point: c * P(c^2) >= tanh(c) (never below the value whose bfloat16 rounding is 1.0)
    A = np.vstack([A, np.hstack([-V[-1], [0.0]])])
    b = np.hstack([b, -y[-1]])
    cost = np.zeros(k + 2)
    cost[-1] = 1
    r = linprog(cost, A_ub=A, b_ub=b, bounds=[(None, None)] * (k + 2), method="highs")
    assert r.status == 0
    return r.x[:-1], r.x[-1]


def bf16(v):
    u = np.asarray(v, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16 << 16
    return u.astype(np.uint32).view(np.float32)


def evaluate(a, x):
    a = a.astype(np.float32)
    xc = np.clip(x.astype(np.float32), np.float32(-C), np.float32(C))
    t = xc * xc
    p = np.full_like(t, a[-1])
    for j in range(len(a) - 2, -1, -1):
        p = (p * t).astype(np.float32) + a[j]     # two roundings >= the fused one: an upper bound
    return (p * xc).astype(np.float32)


if __name__ == "__main__":
    a, e = fit()
    print(f"/* tools/make_tanh_poly.py: C = {C}, degree {2 * K + 1}, minimax relative error {e:.3e} */")
    print("constexpr float kTanhPoly[%d] = {%s};" % (K + 1, ", ".join(f"{float(np.float32(v))!r}f" for v in a)))
    x = np.concatenate([np.linspace(-8, 8, 2_000_001), np.linspace(-0.01, 0.01, 200_001)])
    got, ref = evaluate(a, x), np.tanh(x.astype(np.float64))
    err = np.abs(got - ref)
    print(f"float32 Horner: max abs error {err.max():.3e} (at x = {x[err.argmax()]:.3f}); "
          f"max relative error on |x| <= {C}: {(err / np.maximum(np.abs(ref), 1e-30))[np.abs(x) <= C].max():.3e}")
    rb, gb = bf16(ref.astype(np.float32)), bf16(got)
    ulp = np.abs(gb.view(np.int32).astype(np.int64) - rb.view(np.int32).astype(np.int64)) >> 16
    print(f"after bfloat16 rounding: {100 * (ulp == 0).mean():.2f} % equal to bf16(tanh), max {ulp.max()} ulp apart; "
          f"|x| >= {C}: {int((gb[np.abs(x) >= C] != np.sign(x[np.abs(x) >= C])).sum())} values not +-1")
