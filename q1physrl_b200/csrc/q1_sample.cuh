/*
 * q1_sample.cuh -- sampling from the policy's output distribution for one env (device code shared by
 * k_sample_actions and the fused policy kernel).  Reference: q1physrl/action_dist.py.
 */
#pragma once

#include "q1_tick.cuh"

namespace q1 {

/* Q1PhysActionDist (action_dist.py:199-243) for one row of policy outputs `row` = per key (logit
 * of 0, logit of 1), then (mean, log_std) of the mouse action: one Categorical(2) per key, then
 * GaussianSquashedGaussian (mean, log_std clipped as in action_dist.py:67-76; squash =
 * clip(NormalCDF(raw / 0.90685), 1e-6, 1 - 1e-6) * (high - low) + low, action_dist.py:151, 186-192).
 * deterministic: argmax / squash(mean) (action_dist.py:84-88).  Noise: Philox4x32-10 keyed by
 * `seed`, counter (global env index, step).  Returns the key bit mask; *mouse receives the action. */
__device__ __forceinline__ uint32_t sample_action_row(const float *row, int num_keys, float low,
                                                      float high, bool deterministic, uint64_t seed,
                                                      uint64_t step, uint64_t gidx, float *mouse)
{
    /* one Philox block per (env, step): 16 bits per key draw, 2 x 32 bits for the Gaussian */
    uint32_t w[4] = {0, 0, 0, 0};
    if (!deterministic)
        philox4x32((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)step,
                   (uint32_t)(step >> 32) ^ 0x504F4C00u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    uint32_t keybits = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k < num_keys) {
            const float l0 = row[2 * k], l1 = row[2 * k + 1];
            bool key;
            if (deterministic) {
                key = l1 > l0;
            } else {
                const float p1 = __fdividef(1.0f, 1.0f + __expf(l0 - l1));  /* softmax over two logits */
                const uint32_t bits = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
                const float u = ((float)bits + 0.5f) * (1.0f / 65536.0f);
                key = u < p1;
            }
            keybits |= (key ? 1u : 0u) << k;
        }
    }
    float raw = fminf(fmaxf(row[2 * num_keys], -3.0f), 3.0f);             /* clipped mean */
    if (!deterministic) {
        const float log_std = fminf(fmaxf(row[2 * num_keys + 1], -20.0f), 2.0f);
        const float u1 = ((float)(w[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u2 = ((float)(w[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float eps = sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530717958647692f * u2); /* Box-Muller */
        raw = raw + __expf(log_std) * eps;
    }
    const float scale = 0.5f * 1.8137f;
    float cdf = 0.5f * erfcf(-(raw / scale) * 0.70710678118654752440f);   /* NormalCDF */
    cdf = fminf(fmaxf(cdf, 1e-6f), 1.0f - 1e-6f);
    *mouse = cdf * (high - low) + low;
    return keybits;
}

} // namespace q1
