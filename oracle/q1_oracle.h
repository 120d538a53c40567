/*
 * q1_oracle.h -- CPU restatement of q1physrl_env's per-tick movement step.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA path is compared with; it is never the
 * thing shipped or measured as the product.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * Parity pin: the reference holds no golden vectors for this path, so this
 * restatement is pinned against the unmodified reference source executed in the build container
 * (tests/golden/make_golden.py -> tests/golden/ fixtures, and tests/test_oracle_vs_reference.py when
 * /root/reference is mounted).
 *
 * All citations are relative to the reference checkout:
 *   phys = q1physrl_env/q1physrl_env/phys.py,  env = q1physrl_env/q1physrl_env/env.py
 */
#ifndef Q1_ORACLE_H
#define Q1_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* POD mirror of env.Config (env:132-148), widths as NumPy 2 executes them. */
typedef struct q1o_config {
    double time_delta;
    double time_limit;
    double action_range;       /* continuous mouse action bound */
    double key_press_delay;
    double fmove_max;          /* rounded to f32 before use (env:260-261) */
    double smove_max;
    double zero_start_prob;
    double initial_yaw_lo;
    double initial_yaw_hi;
    double max_initial_speed;
    int32_t discrete_yaw_steps; /* -1 = continuous */
    int32_t allow_yaw;
    int32_t speed_reward;
    int32_t hover;
    int32_t smooth_keys;
    int32_t auto_jump;
    int32_t allow_jump;
    int32_t reserved;          /* bit 0: evaluate env:230 in float64, as NumPy < 2 does */
} q1o_config;

/* Per-env state, struct of arrays, the reference's own widths (SURVEY.md 8(a) "canonical state"). */
typedef struct q1o_state {
    float   *vel;            /* (n,3) f32   phys:159 */
    double  *z_pos;          /* (n,)  f64   phys:158 */
    double  *yaw;            /* (n,)  f64   env:376 == decoder _yaw env:202 */
    double  *time_remaining; /* (n,)  f64   env:377 */
    uint8_t *on_ground;      /* (n,)        phys:160 */
    uint8_t *jump_released;  /* (n,)        phys:161 */
    uint8_t *zero_start;     /* (n,)        env:379 */
    uint8_t *last_keys;      /* (n,nk)      env:201 */
    double  *last_press;     /* (n,nk) f64  env:200 */
} q1o_state;

int q1o_num_keys(const q1o_config *cfg);

/* One lockstep tick of env.VectorPhysEnv.vector_step (env:482-510) for n envs.
 * keys: (n,nk) key actions already reduced to bit 0; mouse: (n,) f64 raw mouse action
 * (continuous value, or the integer index for discrete yaw).  obs: (n,6) f64. */
void q1o_step(const q1o_config *cfg, int64_t n, q1o_state *st,
              const uint8_t *keys, const double *mouse,
              double *obs, float *reward, uint8_t *done);

/* env.ActionDecoder.map (env:225-269) alone, for the standalone decoder. */
void q1o_decode(const q1o_config *cfg, int64_t n,
                uint8_t *last_keys, double *last_press, double *yaw,
                const uint8_t *keys, const double *mouse,
                const float *z_vel, const double *time_remaining,
                double *yaw_out, int64_t *smove, int64_t *fmove, uint8_t *jump);

/* phys.apply (phys:184-197) with arbitrary per-row inputs.  dt_f32: the time_delta array was
 * float32 (the f32 widths NumPy then uses for friction / gravity, see q1_oracle.c). */
void q1o_phys_apply(int64_t n,
                    const double *yaw, const double *pitch, const double *roll,
                    const double *fmove, const double *smove, const uint8_t *button2,
                    const double *time_delta, int dt_f32,
                    const double *z_in, const float *vel_in,
                    const uint8_t *og_in, const uint8_t *jr_in,
                    double *z_out, float *vel_out, uint8_t *og_out, uint8_t *jr_out);

/* The same for a FLOAT64 velocity array (what PlayerState.from_df builds): every intermediate f64. */
void q1o_phys_apply_vel64(int64_t n,
                          const double *yaw, const double *pitch, const double *roll,
                          const double *fmove, const double *smove, const uint8_t *button2,
                          const double *time_delta, int dt_f32,
                          const double *z_in, const double *vel_in,
                          const uint8_t *og_in, const uint8_t *jr_in,
                          double *z_out, double *vel_out, uint8_t *og_out, uint8_t *jr_out);

/* Observation of the current state (env:392-408). */
void q1o_observe(const q1o_config *cfg, int64_t n, const q1o_state *st, double *obs);

/* Episode (re-)initialisation of env i from five explicit uniform draws in [0,1)
 * (env:428-480 restated as a pure function of the draws). */
void q1o_reset_env(const q1o_config *cfg, q1o_state *st, int64_t i, const double u[5]);

/* Philox4x32-10 counter-based generator shared with the CUDA path (integer work, bit-exact). */
void q1o_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t k0, uint32_t k1, uint32_t out[4]);

/* The five reset draws for (seed, global env index, epoch), as the CUDA reset kernels make them. */
void q1o_reset_draws(uint64_t seed, uint64_t env_index, uint32_t epoch, double u[5]);

/* Device-side synthetic policies of the rollout kernel, restated: fills keys (nk bytes) and mouse. */
void q1o_policy_action(const q1o_config *cfg, int32_t policy, uint64_t seed, uint64_t env_index,
                       uint32_t tick, uint8_t *keys, double *mouse);

/* Batched forms of the two above over envs [env_index_base, env_index_base + n). */
void q1o_policy_actions(const q1o_config *cfg, int32_t policy, uint64_t seed,
                        uint64_t env_index_base, int64_t n, uint32_t tick,
                        uint8_t *keys, double *mouse);
void q1o_reset_philox(const q1o_config *cfg, q1o_state *st, int64_t n, uint64_t seed,
                      uint64_t env_index_base, uint32_t epoch, const uint8_t *mask);

#ifdef __cplusplus
}
#endif
/* sin / cos of n doubles through the C library, as NumPy does for phys.py:58-59 */
void q1o_sincos(int64_t n, const double *x, double *s, double *c);

#endif
