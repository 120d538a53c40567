/*
 * q1phys.h -- C ABI of the B200-native q1physrl_env movement step (libq1phys.so).
 *
 * The reference (matthewearl/q1physrl) has no FFI layer: its boundary for this path is a set of
 * Python classes.  Each entry point below replaces one of those methods; the Python mirror in
 * q1physrl_b200/ (and the stub a reference maintainer would add) binds them
 * with ctypes.  Citations: env = q1physrl_env/q1physrl_env/env.py, phys = .../phys.py.
 *
 * Conventions
 *   - every function returns 0 on success or a negative Q1_E* code; q1_last_error() gives the text
 *     (thread-local).  Nothing throws across the ABI.
 *   - buffers are caller-owned.  Unless the name ends in _host every pointer is a DEVICE pointer on
 *     the handle's GPU; `stream` is a cudaStream_t passed as void* (NULL = the legacy default
 *     stream).  The library owns only the persistent per-env state.
 *   - a handle is bound to one device and is not thread-safe; use one handle per GPU.
 *   - there is no CPU implementation behind this ABI: q1_create fails when no CUDA device exists.
 */
#ifndef Q1PHYS_H
#define Q1PHYS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Q1_ABI_VERSION 2

enum {
    Q1_OK = 0,
    Q1_EINVAL = -1,   /* bad argument / unsupported configuration */
    Q1_ECUDA = -2,    /* a CUDA runtime call failed */
    Q1_ENODEV = -3,   /* no usable CUDA device */
    Q1_ENOMEM = -4
};

/* q1_create flags */
enum {
    Q1_F_TRACK_RETURNS = 1u << 0,   /* keep a per-env f64 episode return + on-device episode metrics */
    Q1_F_FORCE_F64_STAMPS = 1u << 1, /* keep env:200 key time stamps in f64 even when u8 tick counters are exact */
    Q1_F_IEEE_DIVISION = 1u << 2,   /* divide with the CUDA IEEE intrinsics instead of the (equally exact,
                                       branch-free) reciprocal-multiply sequences: slower, for self-checks */
    Q1_F_NUMPY1_PROMOTION = 1u << 3 /* env:230 `np.float32(720) * time_delta` as NumPy < 2 evaluates it (in
                                       float64) -- the reference pins numpy 1.18.2; the default is NumPy 2's
                                       float32 product, which the oracle and all fixtures were recorded with */
};

/* Element type of the `mouse` array handed to q1_step / q1_step_host. */
enum {
    Q1_MOUSE_F32 = 0,  /* the action space's own dtype (env:214-215); exact for discrete indices too */
    Q1_MOUSE_I32 = 1,  /* discrete_yaw_steps index */
    Q1_MOUSE_F64 = 2   /* the f64 column ActionDecoder._fix_actions produces (env:221-223) */
};

/* Built-in device-side policies of q1_rollout. */
enum {
    Q1_POLICY_RANDOM = 0,      /* uniform keys, uniform mouse in the action range */
    Q1_POLICY_STRAFE_JUMP = 1  /* scripted: forward + alternating strafe/turn, jump taps */
};

/* POD mirror of env.Config (env:132-148). */
typedef struct q1_config {
    int64_t num_envs;
    double zero_start_prob;
    double initial_yaw_lo;      /* initial_yaw_range[0] */
    double initial_yaw_hi;      /* initial_yaw_range[1] */
    double max_initial_speed;
    double time_delta;
    double time_limit;
    double action_range;
    double fmove_max;
    double smove_max;
    double key_press_delay;
    int32_t allow_yaw;
    int32_t discrete_yaw_steps; /* -1 = continuous mouse action */
    int32_t speed_reward;
    int32_t hover;
    int32_t smooth_keys;
    int32_t auto_jump;
    int32_t allow_jump;
    int32_t reserved;           /* 0; q1_decode_host (which has no flags argument) reads bit 0 as
                                   Q1_F_NUMPY1_PROMOTION */
} q1_config;

typedef struct q1_env q1_env; /* opaque */

typedef struct q1_env_info {
    int64_t num_envs;
    int32_t num_keys;           /* env:206-207 */
    int32_t device;
    int32_t f64_stamps;         /* 1: f64 key time stamps, 0: u8 tick counters */
    int32_t track_returns;
    int32_t key_delay_ticks;    /* ceil(key_press_delay / time_delta) in counter mode */
    int32_t state_bytes_per_env;
    uint64_t env_index_base;
    uint64_t seed;
    uint64_t ticks;             /* q1_step / q1_rollout ticks executed so far */
} q1_env_info;

/* Host-side view of the full per-env state in the reference's own layout and widths
 * (phys:156-161, env:200-202, 375-379).  Any pointer may be NULL to skip that field. */
typedef struct q1_state_view {
    float   *vel;             /* (n,3) */
    double  *z_pos;           /* (n,)  */
    double  *yaw;             /* (n,)  */
    double  *time_remaining;  /* (n,)  */
    uint8_t *on_ground;       /* (n,)  */
    uint8_t *jump_released;   /* (n,)  */
    uint8_t *zero_start;      /* (n,)  */
    uint8_t *last_keys;       /* (n,num_keys) */
    double  *last_press;      /* (n,num_keys) */
    double  *episode_return;  /* (n,)  only with Q1_F_TRACK_RETURNS */
} q1_state_view;

/* On-device episode statistics (q1physrl/train.py:54-57 `zero_start_total_reward`, and the
 * episode_reward_mean / max RLLib tracks, train.py:67-71). */
typedef struct q1_metrics {
    double  zero_start_return_sum;
    int64_t zero_start_episodes;
    double  return_sum;
    int64_t episodes;
    double  return_max;
} q1_metrics;

const char *q1_last_error(void);
int q1_abi_version(void);
int q1_device_count(int *count);

/* env:206-207 -- number of key actions for this config (3 or 4). */
int q1_num_keys(const q1_config *cfg);

/* env.VectorPhysEnv.__init__ (env:410-426) minus the implicit vector_reset.  `env_index_base` is
 * the global index of this shard's env 0: reset draws are a pure function of
 * (seed, global env index, reset epoch), so a sharded run equals the single-GPU run. */
int q1_create(const q1_config *cfg, int device, uint64_t seed, uint64_t env_index_base,
              uint32_t flags, q1_env **out);
int q1_destroy(q1_env *env);
int q1_info(const q1_env *env, q1_env_info *out);
int q1_sync(q1_env *env, void *stream);

/* env.VectorPhysEnv.vector_reset (env:428-455).  obs (n,6) f32, may be NULL. */
int q1_reset_all(q1_env *env, float *obs, void *stream);
/* Batched env.VectorPhysEnv.reset_at (env:457-480): resets envs with mask[i] != 0 and overwrites
 * their obs rows; other rows are left untouched.  obs may be NULL. */
int q1_reset_masked(q1_env *env, const uint8_t *mask, float *obs, void *stream);
/* The two calls above with HOST buffers (mask (n,) u8, obs (n,6) f32; obs may be NULL). */
int q1_reset_all_host(q1_env *env, float *obs_host);
int q1_reset_masked_host(q1_env *env, const uint8_t *mask_host, float *obs_host);
/* env.VectorPhysEnv.reset_at(index) (env:457-480).  obs6_host: 6 floats in HOST memory. */
int q1_reset_at_host(q1_env *env, int64_t index, float *obs6_host);

/* env.VectorPhysEnv.vector_step (env:482-510), one lockstep tick for all envs.
 *   keys        (n,num_keys) u8, bit 0 of each byte is the key action (env:228)
 *   mouse       (n,) raw mouse action (continuous value, or the index for discrete_yaw_steps
 *               != -1) of element type mouse_kind (Q1_MOUSE_*); ignored (may be NULL) when
 *               allow_yaw is 0
 *   obs         (n,6) f32    reward (n,) f32    done (n,) u8    zero_start (n,) u8 or NULL
 *   auto_reset  0: reference behaviour, the caller resets finished envs;
 *               1: envs whose episode ended are re-initialised in the same launch and their obs
 *                  row holds the first observation of the new episode. */
int q1_step(q1_env *env, const uint8_t *keys, const void *mouse, int mouse_kind, float *obs,
            float *reward, uint8_t *done, uint8_t *zero_start, int auto_reset, void *stream);
/* Same call with HOST buffers: copies the actions in, steps, copies the results out and
 * synchronises.  This is what PhysEnv.step / VectorPhysEnv.vector_step bind for NumPy callers.
 * With buffers from q1_host_alloc (or any page-locked, device-mapped memory) the step kernel itself
 * bulk-loads the actions from and bulk-stores the results to host memory, so both PCIe directions
 * run for the whole launch and a step costs max(H2D, D2H) of the wire (Q1PHYS_HOST_DIRECT=0 selects
 * the older chunked copy / tick / copy pipeline instead).  Pageable memory works too: up to 65 536 envs
 * bounce through one page-locked, device-mapped buffer owned by the handle (two memcpy in, one launch,
 * four memcpy out); larger batches through staging copies at the driver's speed. */
int q1_step_host(q1_env *env, const uint8_t *keys, const void *mouse, int mouse_kind, float *obs,
                 float *reward, uint8_t *done, uint8_t *zero_start, int auto_reset);

/* Page-locked host memory for the *_host entry points. */
int q1_host_alloc(uint64_t bytes, void **out);
int q1_host_free(void *ptr);

/* `ticks` consecutive vector_step calls in one launch with a built-in policy generating the actions
 * on the device (state stays in registers; finished envs auto-reset).  reward_sum (n,) f32 receives
 * the per-env sum of rewards over the call, obs (n,6) the final observation; either may be NULL. */
int q1_rollout(q1_env *env, int policy, int ticks, uint64_t policy_seed, float *obs,
               float *reward_sum, void *stream);

/* -- trajectory recorder: q1physrl/analyse.py:197-240 `eval_sim`, N envs at once, on the device ----- */

/* Where a recorded rollout takes its actions from. */
enum {
    Q1_ACTIONS_ARRAYS = 0,   /* caller arrays laid out [tick][env] (a scripted or pre-computed stream) */
    Q1_ACTIONS_BUILTIN = 1   /* a built-in device-side policy (Q1_POLICY_*) */
};
typedef struct q1_action_source {
    int32_t kind;             /* Q1_ACTIONS_* */
    int32_t builtin_policy;   /* Q1_POLICY_*, kind BUILTIN */
    uint64_t policy_seed;     /* kind BUILTIN */
    const uint8_t *keys;      /* kind ARRAYS: (ticks, n, num_keys) u8, bit 0 = the key action */
    const void *mouse;        /* kind ARRAYS: (ticks, n) of mouse_kind; may be NULL without allow_yaw */
    int32_t mouse_kind;       /* Q1_MOUSE_* */
    int32_t reserved;
} q1_action_source;

/* Per-tick record, every array [ticks][n] (tick-major, C order); any pointer may be NULL to skip that
 * field.  Row t holds what analyse.py:214-229 appends in iteration t: the movement state
 * (`e.player_state`, analyse.py:218), its time_remaining and the observation the policy saw
 * (analyse.py:219) BEFORE the tick; the action (analyse.py:220); the move command ActionDecoder.map
 * makes of it (analyse.py:215-216, 221-224); reward and done of the tick (analyse.py:226-228).
 * Field widths are the reference's (EvalSimResult, analyse.py:71-82) except obs, which is f32 like
 * every observation of this library. */
typedef struct q1_record_view {
    float   *vel;             /* (T,n,3) */
    double  *z_pos;           /* (T,n)   */
    uint8_t *on_ground;       /* (T,n)   */
    uint8_t *jump_released;   /* (T,n)   */
    double  *time_remaining;  /* (T,n)   */
    float   *obs;             /* (T,n,6) */
    uint8_t *keys;            /* (T,n,num_keys) */
    float   *mouse;           /* (T,n)   the mouse action in the action space's dtype (env:214-215) */
    double  *yaw;             /* (T,n)   */
    int64_t *smove;           /* (T,n)   */
    int64_t *fmove;           /* (T,n)   */
    uint8_t *jump;            /* (T,n)   */
    float   *reward;          /* (T,n)   */
    uint8_t *done;            /* (T,n)   */
} q1_record_view;

/* record_flags */
enum {
    /* analyse.py:215-216 feeds its shadow ActionDecoder the OBSERVATION's z velocity (quantised and
     * divided by 200, env:399-400) where the env's own decoder sees the raw one (env:487), so with
     * auto_jump the `jump` it records is `obs[Z_VEL] <= 16`.  Set: record that value (EvalSimResult
     * parity); clear: record the jump the env actually executed. */
    Q1_RECORD_SHADOW_JUMP = 1u << 0
};

/* `ticks` consecutive vector_step calls in ONE launch (state in registers throughout), each tick's
 * row written from the same tick routine that advances the env.  auto_reset = 0 is the reference's
 * behaviour: an env whose episode ended keeps stepping with negative time_remaining (env:505-506)
 * and the caller cuts its rows at the first done.  final_obs (n,6) or NULL receives the observation
 * after the last tick.  DEVICE pointers in `actions` and `record`. */
int q1_rollout_record(q1_env *env, const q1_action_source *actions, int ticks, int auto_reset,
                      uint32_t record_flags, const q1_record_view *record, float *final_obs,
                      void *stream);
/* Same with HOST pointers in `actions`, `record` and final_obs_host; synchronises. */
int q1_rollout_record_host(q1_env *env, const q1_action_source *actions, int ticks, int auto_reset,
                           uint32_t record_flags, const q1_record_view *record, float *final_obs_host);

/* Adds `delta` to the handle's tick counter (q1_env_info.ticks; the position of q1_rollout's built-in
 * action streams).  For callers that replay a captured CUDA graph: the library counts a q1_step only
 * when it is called, i.e. during capture, not when the graph is replayed. */
int q1_advance_ticks(q1_env *env, int64_t delta);

/* Observation of the current state without stepping (env:392-400). */
int q1_observe(q1_env *env, float *obs, void *stream);
int q1_observe_host(q1_env *env, float *obs_host);

/* Full state copy-out / copy-in through HOST arrays in the reference layout. */
int q1_get_state_host(q1_env *env, const q1_state_view *view);
int q1_set_state_host(q1_env *env, const q1_state_view *view);

/* Episode statistics accumulated on the device since creation or the last clear (HOST struct).
 * Requires Q1_F_TRACK_RETURNS. */
int q1_get_metrics_host(q1_env *env, int clear, q1_metrics *out);

/* phys.apply (phys:184-197) on explicit arrays: a pure function, no handle.  All (n,) except
 * vel (n,3); pitch / roll may be NULL (= 0).  Output arrays must not alias the inputs.
 * time_delta_f32 != 0: the caller's time_delta array was float32 (values passed widened); NumPy
 * then keeps friction, gravity and dt * z_vel in f32 (q1physrl/analyse.py:110 does this), and
 * the library reproduces those widths. */
int q1_phys_apply(int device, int64_t n,
                  const double *yaw, const double *pitch, const double *roll,
                  const double *fmove, const double *smove, const uint8_t *button2,
                  const double *time_delta, int time_delta_f32,
                  const double *z_pos, const float *vel, const uint8_t *on_ground,
                  const uint8_t *jump_released,
                  double *z_pos_out, float *vel_out, uint8_t *on_ground_out,
                  uint8_t *jump_released_out, void *stream);
/* Same with HOST arrays. */
int q1_phys_apply_host(int device, int64_t n,
                       const double *yaw, const double *pitch, const double *roll,
                       const double *fmove, const double *smove, const uint8_t *button2,
                       const double *time_delta, int time_delta_f32,
                       const double *z_pos, const float *vel, const uint8_t *on_ground,
                       const uint8_t *jump_released,
                       double *z_pos_out, float *vel_out, uint8_t *on_ground_out,
                       uint8_t *jump_released_out);

/* phys.apply for a PlayerState whose velocity array is FLOAT64 -- what PlayerState.from_df (phys:163-170)
 * builds.  NumPy then never rounds the velocity to f32 (friction speed phys:85, store phys:190, z velocity
 * phys:119-122 all in f64); vel / vel_out are (n,3) f64 HOST arrays, the rest as q1_phys_apply_host. */
int q1_phys_apply_vel64_host(int device, int64_t n,
                             const double *yaw, const double *pitch, const double *roll,
                             const double *fmove, const double *smove, const uint8_t *button2,
                             const double *time_delta, int time_delta_f32,
                             const double *z_pos, const double *vel, const uint8_t *on_ground,
                             const uint8_t *jump_released,
                             double *z_pos_out, double *vel_out, uint8_t *on_ground_out,
                             uint8_t *jump_released_out);

/* EvalSimResult.hypothetical_delta_speeds (q1physrl/analyse.py:92-118) in one launch instead of
 * num_angles phys.apply calls: delta_speed[a * n + t] = |v'_xy| - |v_xy| (f32) of one phys.apply
 * tick on row t with yaw = base_yaw[t] + rel_angles[a] and constant fmove / smove / time_delta.
 * HOST arrays; state arrays as for q1_phys_apply. */
int q1_delta_speed_sweep_host(int device, int64_t n, int64_t num_angles, const double *base_yaw,
                              const double *rel_angles, double fmove, double smove,
                              const uint8_t *button2, double time_delta, int time_delta_f32,
                              const double *z_pos, const float *vel, const uint8_t *on_ground,
                              const uint8_t *jump_released, float *delta_speed);

/* env.ActionDecoder.map (env:225-269) on explicit decoder state, HOST arrays: last_keys
 * (n,num_keys) u8, last_press (n,num_keys) f64 and yaw (n,) f64 are updated in place; mouse is f64
 * here (the standalone decoder is driven by arbitrary callers, mkdemo.py:47-50). */
int q1_decode_host(const q1_config *cfg, int device, int64_t n,
                   uint8_t *last_keys, double *last_press, double *yaw,
                   const uint8_t *keys, const double *mouse,
                   const float *z_vel, const double *time_remaining,
                   int64_t *smove, int64_t *fmove, uint8_t *jump);

/* Q1PhysActionDist (q1physrl/action_dist.py:199-243) on a batch of policy outputs, DEVICE arrays:
 * logits (n, 2 * num_keys + 2) f32 = per key (logit of 0, logit of 1), then (mean, log_std) of the
 * mouse action.  Writes keys (n, num_keys) u8 and mouse (n,) f32 in the layout q1_step consumes.
 * deterministic != 0: argmax keys and squash(mean) (action_dist.py:84-88); otherwise Categorical and
 * Gaussian draws from a counter-based generator keyed by (seed, env_index_base + i, step).
 * step_device (DEVICE pointer, may be NULL) overrides `step` with the value it points to, so a
 * captured CUDA graph can advance the noise stream between replays. */
int q1_sample_actions(int device, int64_t n, int num_keys, const float *logits, double action_low,
                      double action_high, int deterministic, uint64_t seed, uint64_t step,
                      const uint64_t *step_device, uint64_t env_index_base, uint8_t *keys,
                      float *mouse, void *stream);

/* The shipped policy network fused into one kernel: RLLib fcnet obs(6) -> tanh 256
 * -> tanh 256 -> 2 * num_keys + 2 outputs (checkpoint arrays default_policy/fc_1, fc_2, fc_out), then
 * q1_sample_actions' sampling.  All three layers on the tensor cores (tcgen05, bf16 operands, fp32
 * accumulation in TMEM; layer 1 with the observation and its weights split into bf16 pieces, so that
 * it is as exact as an fp32 product).  Weights are HOST arrays in the checkpoint's (in, out) layout:
 * w1 (6, 256), w2 (256, 256), w3 (256, 2 * num_keys + 2). */
typedef struct q1_policy q1_policy; /* opaque */
int q1_policy_create(int device, int num_keys, const float *w1, const float *b1, const float *w2,
                     const float *b2, const float *w3, const float *b3, q1_policy **out);
int q1_policy_destroy(q1_policy *policy);
/* Waits for the device and reports whether a policy kernel's internal watchdog fired (a role of the
 * warp-specialised kernel gave up waiting for another: a bug, never a property of the input).  The
 * launches are asynchronous, so this is where such a fault surfaces: Q1_OK, or Q1_ECUDA with the
 * record in q1_last_error() and the results of the launches since the last check invalid. */
int q1_policy_check(q1_policy *policy);
/* obs (n, 6) f32, keys (n, num_keys) u8, mouse (n,) f32, logits_out (n, 2 * num_keys + 2) f32 or
 * NULL: DEVICE arrays.  Other arguments as q1_sample_actions. */
int q1_policy_act(q1_policy *policy, int64_t n, const float *obs, double action_low,
                  double action_high, int deterministic, uint64_t seed, uint64_t step,
                  const uint64_t *step_device, uint64_t env_index_base, uint8_t *keys, float *mouse,
                  float *logits_out, void *stream);

/* The closed loop policy -> sample -> env tick -> observation -> policy ... for `ticks` ticks in ONE
 * launch (q1physrl/action_dist.py:84-101, 186-243 + env:482-510; BASELINE config 5): every CTA keeps the
 * state of its envs in shared memory and the activations in tensor memory, so a tick costs no launch and
 * no HBM traffic.  The sampling noise is keyed by (seed, global env index, the handle's tick counter).
 * auto_reset as q1_step; deterministic as q1_sample_actions.  record / record_flags (may be NULL / 0):
 * the per-tick record of q1_rollout_record, DEVICE pointers; final_obs (n,6) and reward_sum (n,) f32
 * DEVICE arrays or NULL.  Needs the counter form of the key timers (q1_env_info.f64_stamps == 0). */
int q1_policy_rollout(q1_policy *policy, q1_env *env, int ticks, int auto_reset, int deterministic,
                      uint64_t seed, double action_low, double action_high, uint32_t record_flags,
                      const q1_record_view *record, float *final_obs, float *reward_sum, void *stream);
/* Same with HOST pointers in `record` and final_obs_host; synchronises.  This is analyse.eval_sim
 * (q1physrl/analyse.py:197-240) for a policy that lives on the device. */
int q1_policy_rollout_host(q1_policy *policy, q1_env *env, int ticks, int auto_reset, int deterministic,
                           uint64_t seed, double action_low, double action_high, uint32_t record_flags,
                           const q1_record_view *record, float *final_obs_host);

/* Self-check of the branch-free reciprocal-multiply division sequences the kernels use against the
 * CUDA IEEE intrinsics, on ~`samples` random operand pairs per class (bit comparison):
 *   [0] reciprocal  [1] a / variable b  [2] a / constant  [3] wish_vel / wish_speed range
 *   [4] new_speed / speed range  [5] the f32 observation quotients (v / 200, z / 100), exhaustive
 *   [6], [7] the same two f32 quotient sets with a second correction step.
 * Every count must be 0. */
int q1_selftest_division(int device, uint64_t samples, uint64_t seed, uint64_t mismatches[8]);

/* Exact checkpoint / resume (the reference env has none; RLLib's trainer.save() loses env state,
 * train.py:119-133): the raw device image of the handle -- state blocks, key stamps, reset epochs,
 * episode returns, metric accumulators -- plus its tick count and RNG key.  A handle created with
 * the same config and flags that loads the image continues bit for bit like the one that saved it,
 * resets included.  `bytes` from q1_snapshot_bytes. */
int q1_snapshot_bytes(const q1_env *env, uint64_t *bytes);
int q1_snapshot_save_host(q1_env *env, void *buffer, uint64_t bytes);
int q1_snapshot_load_host(q1_env *env, const void *buffer, uint64_t bytes);

/* sin and cos of n HOST doubles (radians) as the kernels compute them: glibc 2.39's __sin / __cos
 * (sysdeps/ieee754/dbl-64/s_sin.c, FMA build) restated on the device, i.e. the bits np.sin /
 * np.cos return in the reference (phys.py:58-59, env.py:475-476).  For parity tests. */
int q1_sincos_host(int device, int64_t n, const double *x, double *sin_out, double *cos_out);

#ifdef __cplusplus
}
#endif
#endif /* Q1PHYS_H */
