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
 * deterministic: argmax / squash(mean) (action_dist.py:84-88).  Returns the key bit mask; *mouse
 * receives the action.
 *
 * Noise: two Philox4x32-10 blocks keyed by `seed`, counter (global env index, step, block tag).  Block 0
 * gives each key its own 32-bit word, of which the top 24 bits make a uniform in (0, 1) (the resolution of
 * a float significand: a key with probability >= 2^-24 can fire); block 1 gives the two uniforms of the
 * Box-Muller normal.  The arithmetic is float32 with the accurate library functions (expf, logf, cospif,
 * erfcf: <= 2 ulp each), i.e. the sampled action is a function of the float32 logits to within a few ulp of
 * float32 -- orders of magnitude inside what bf16 tensor-core logits differ from fp32 ones by.
 * tests/test_sampling_gpu.py checks the key frequencies down to p = 1e-5 and the mouse action's
 * distribution against scipy.stats.norm. */
/* The part that does not need the logits: callers with something to wait for (the fused policy kernel's env
 * rows) draw the noise first.  sample_action_row = draw_action_noise + apply_action_noise, same arithmetic. */
struct ActionNoise {
    float u[4]; /* one uniform in (0, 1) per key */
    float eps;  /* standard normal */
};
__device__ __forceinline__ ActionNoise draw_action_noise(bool deterministic, uint64_t seed, uint64_t step, uint64_t gidx)
{
    ActionNoise z = {{0.0f, 0.0f, 0.0f, 0.0f}, 0.0f};
    if (!deterministic) {
        uint32_t w[4], g[4];
        philox4x32((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)step,
                   (uint32_t)(step >> 32) ^ 0x504F4C00u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
        philox4x32((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)step,
                   (uint32_t)(step >> 32) ^ 0x504F4C01u, (uint32_t)seed, (uint32_t)(seed >> 32), g);
#pragma unroll
        for (int k = 0; k < 4; k++)
            z.u[k] = ((float)(w[k] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u1 = ((float)(g[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u2 = ((float)(g[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        z.eps = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);      /* Box-Muller */
    }
    return z;
}
__device__ __forceinline__ uint32_t apply_action_noise(const float *row, int num_keys, float low, float high,
                                                       bool deterministic, const ActionNoise &z, float *mouse)
{
    uint32_t keybits = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k < num_keys) {
            const float l0 = row[2 * k], l1 = row[2 * k + 1];
            bool key;
            if (deterministic) {
                key = l1 > l0;
            } else {
                /* u < softmax(l0, l1)[1] = 1 / (1 + e^(l0 - l1)), without the division: u (1 + e) < 1 */
                key = z.u[k] + z.u[k] * expf(l0 - l1) < 1.0f;
            }
            keybits |= (key ? 1u : 0u) << k;
        }
    }
    float raw = fminf(fmaxf(row[2 * num_keys], -3.0f), 3.0f);             /* clipped mean */
    if (!deterministic) {
        const float log_std = fminf(fmaxf(row[2 * num_keys + 1], -20.0f), 2.0f);
        raw = raw + expf(log_std) * z.eps;
    }
    const float scale = 0.5f * 1.8137f;
    float cdf = 0.5f * erfcf(-(raw / scale) * 0.70710678118654752440f);   /* NormalCDF */
    cdf = fminf(fmaxf(cdf, 1e-6f), 1.0f - 1e-6f);
    *mouse = cdf * (high - low) + low;
    return keybits;
}
__device__ __forceinline__ uint32_t sample_action_row(const float *row, int num_keys, float low,
                                                      float high, bool deterministic, uint64_t seed,
                                                      uint64_t step, uint64_t gidx, float *mouse)
{
    const ActionNoise z = draw_action_noise(deterministic, seed, step, gidx);
    return apply_action_noise(row, num_keys, low, high, deterministic, z, mouse);
}

} // namespace q1
