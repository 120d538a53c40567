/*
 * q1_oracle.c -- scalar CPU restatement of the q1physrl_env movement tick.
 *
 * TEST INFRASTRUCTURE ONLY (see q1_oracle.h).  One env at a time, every intermediate in the
 * width NumPy 2 (NEP 50 promotion) gives it in the reference.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math      (no FMA contraction: the reference never fuses)
 *
 * phys = q1physrl_env/q1physrl_env/phys.py,  env = q1physrl_env/q1physrl_env/env.py
 */
#include "q1_oracle.h"

#include <math.h>
#include <string.h>

/* phys:47-53 -- all np.float32 scalars; every one is exactly representable. */
static const float Q_MAX_SPEED = 320.0f;
static const float Q_ACCELERATE = 10.0f;
static const float Q_FRICTION = 4.0f;
static const float Q_STOP_SPEED = 100.0f;
static const float Q_JUMP_SPEED = 270.0f;
static const float Q_GRAVITY = 800.0f;
static const float Q_FLOOR_HEIGHT = 24.03125f;

/* env:54-58 */
static const float Q_INITIAL_Z = 32.843201f;
static const float Q_INITIAL_VZ = -12.0f;
static const float Q_INITIAL_YAW_ZERO = 90.0f;
/* env:91 */
static const float Q_MAX_YAW_SPEED = 720.0f;

enum { KEY_LEFT = 0, KEY_RIGHT = 1, KEY_FORWARD = 2, KEY_JUMP = 3 }; /* env:61-73 */

int q1o_num_keys(const q1o_config *cfg)
{
    /* env:206-207 */
    return (!cfg->auto_jump && cfg->allow_jump) ? 4 : 3;
}

/* ------------------------------------------------------------------ movement (phys.py) ------ */

typedef struct {
    double z;
    float vx, vy, vz;
    int on_ground, jump_released;
} body_t;

/* phys.apply for one row.  (fwd_x, fwd_y, right_x, right_y) is the 2x2 block _angle_vectors
 * returns (phys:56-66). */
/* dt32: the reference's time_delta array was float32 (analyse.py:110); NumPy then keeps friction
 * (phys:87-90), gravity (phys:122) and dt * z_vel (phys:127) in f32 and forms 10 * dt in f32. */
static void move_body(body_t *b, double fwd_x, double right_x, double fwd_y, double right_y,
                      double fmove, double smove, int jump, double dt, int dt32)
{
    const float dtf = (float)dt;
    const int was_on_ground = b->on_ground; /* phys:191 passes the OLD flag */

    /* phys:95-101: einsum('ijk,ik->ij') is mul, mul, add per component; norm is sqrt(x*x+y*y). */
    double wx = fwd_x * fmove + right_x * smove;
    double wy = fwd_y * fmove + right_y * smove;
    double ws = sqrt(wx * wx + wy * wy);
    double wdx = wx, wdy = wy;
    if (ws > 0) {
        wdx = wx / ws;
        wdy = wy / ws;
    }
    /* phys:103; phys:104-106 scale wish_vel afterwards but nothing reads it again. */
    double wish_speed = ws < (double)Q_MAX_SPEED ? ws : (double)Q_MAX_SPEED;
    if (ws != ws)
        wish_speed = ws;

    /* phys:108 with phys:83-90: friction only for envs standing on the floor. */
    double hx = (double)b->vx, hy = (double)b->vy;
    if (was_on_ground) {
        float speed = sqrtf(b->vx * b->vx + b->vy * b->vy);            /* f32 norm, phys:85 */
        float control = speed > Q_STOP_SPEED ? speed : Q_STOP_SPEED;   /* phys:86 */
        if (dt32) {
            float ns = speed - dtf * control * Q_FRICTION;             /* all f32 */
            if (!(ns > 0))
                ns = 0;
            if (speed > 0) {
                float ratio = ns / speed;
                hx = (double)(b->vx * ratio);
                hy = (double)(b->vy * ratio);
            }
        } else {
            double new_speed = (double)speed - dt * (double)control * (double)Q_FRICTION; /* phys:87 */
            if (!(new_speed > 0))
                new_speed = 0;                                         /* phys:88 */
            if (speed > 0) {
                double ratio = new_speed / (double)speed;              /* phys:90 */
                hx = (double)b->vx * ratio;
                hy = (double)b->vy * ratio;
            }
        }
    }

    /* phys:69-80 */
    double current = hx * wdx + hy * wdy;
    double clipped = (wish_speed > 30 && !was_on_ground) ? 30.0 : wish_speed;
    double add = clipped - current;
    if (!(add > 0))
        add = 0;
    double accel = (dt32 ? (double)(Q_ACCELERATE * dtf) : (double)Q_ACCELERATE * dt) * wish_speed;
    if (add < accel)
        accel = add;
    hx = hx + accel * wdx;
    hy = hy + accel * wdy;

    /* phys:190: the one f64 -> f32 rounding of the horizontal velocity. */
    b->vx = (float)hx;
    b->vy = (float)hy;

    /* phys:112-132 */
    b->jump_released = b->jump_released | !jump;
    int do_jump = was_on_ground && jump && b->jump_released;
    float vz = b->vz + (do_jump ? Q_JUMP_SPEED : 0.0f);                /* phys:119, f32 */
    double z;
    if (dt32) {
        vz = vz - Q_GRAVITY * dtf;                                     /* f32 -= f32 */
        z = b->z + (double)(dtf * vz);
    } else {
        vz = (float)((double)vz - (double)Q_GRAVITY * dt);             /* phys:122, f64 then store */
        z = b->z + dt * (double)vz;                                    /* phys:127 */
    }
    int og = z < (double)Q_FLOOR_HEIGHT;                               /* phys:128 */
    b->z = og ? (double)Q_FLOOR_HEIGHT : z;                            /* phys:129 */
    b->vz = og ? 0.0f : vz;                                            /* phys:130 */
    b->on_ground = og;
}

void q1o_phys_apply(int64_t n,
                    const double *yaw, const double *pitch, const double *roll,
                    const double *fmove, const double *smove, const uint8_t *button2,
                    const double *time_delta, int dt_f32,
                    const double *z_in, const float *vel_in,
                    const uint8_t *og_in, const uint8_t *jr_in,
                    double *z_out, float *vel_out, uint8_t *og_out, uint8_t *jr_out)
{
    for (int64_t i = 0; i < n; i++) {
        /* phys:58-66 (left to right: multiply by pi, then divide by 180) */
        double ay = yaw[i] * M_PI / 180.;
        double ap = pitch ? pitch[i] * M_PI / 180. : 0.0;
        double ar = roll ? roll[i] * M_PI / 180. : 0.0;
        double sy = sin(ay), cy = cos(ay);
        double sp = sin(ap), cp = cos(ap);
        double sr = sin(ar), cr = cos(ar);
        double fwd_x = cp * cy;
        double right_x = (-1 * sr * sp * cy + -1 * cr * -sy);
        double fwd_y = cp * sy;
        double right_y = (-1 * sr * sp * sy + -1 * cr * cy);

        body_t b = { z_in[i], vel_in[3 * i], vel_in[3 * i + 1], vel_in[3 * i + 2],
                     og_in[i] != 0, jr_in[i] != 0 };
        move_body(&b, fwd_x, right_x, fwd_y, right_y, fmove[i], smove[i], button2[i] != 0,
                  time_delta[i], dt_f32);
        z_out[i] = b.z;
        vel_out[3 * i] = b.vx;
        vel_out[3 * i + 1] = b.vy;
        vel_out[3 * i + 2] = b.vz;
        og_out[i] = (uint8_t)b.on_ground;
        jr_out[i] = (uint8_t)b.jump_released;
    }
}

/* phys.apply when PlayerState.vel is FLOAT64 (PlayerState.from_df, phys:163-170, yields that: the
 * notebook / demo comparison path).  NumPy then keeps every intermediate in f64: the friction speed
 * (phys:85 norm of an f64 array), the stored velocity (phys:190 assigns into an f64 array: no rounding)
 * and the z velocity (phys:119-122).  dt_f32 as above: 10 * dt and 800 * dt are then f32 products. */
void q1o_phys_apply_vel64(int64_t n,
                          const double *yaw, const double *pitch, const double *roll,
                          const double *fmove, const double *smove, const uint8_t *button2,
                          const double *time_delta, int dt_f32,
                          const double *z_in, const double *vel_in,
                          const uint8_t *og_in, const uint8_t *jr_in,
                          double *z_out, double *vel_out, uint8_t *og_out, uint8_t *jr_out)
{
    for (int64_t i = 0; i < n; i++) {
        double ay = yaw[i] * M_PI / 180.;
        double ap = pitch ? pitch[i] * M_PI / 180. : 0.0;
        double ar = roll ? roll[i] * M_PI / 180. : 0.0;
        double sy = sin(ay), cy = cos(ay);
        double sp = sin(ap), cp = cos(ap);
        double sr = sin(ar), cr = cos(ar);
        double fwd_x = cp * cy;
        double right_x = (-1 * sr * sp * cy + -1 * cr * -sy);
        double fwd_y = cp * sy;
        double right_y = (-1 * sr * sp * sy + -1 * cr * cy);
        const double dt = time_delta[i];
        const float dtf = (float)dt;
        const int was_on_ground = og_in[i] != 0, jump = button2[i] != 0;
        double vx = vel_in[3 * i], vy = vel_in[3 * i + 1], vz = vel_in[3 * i + 2];

        double wx = fwd_x * fmove[i] + right_x * smove[i];             /* phys:95-101 */
        double wy = fwd_y * fmove[i] + right_y * smove[i];
        double ws = sqrt(wx * wx + wy * wy);
        double wdx = wx, wdy = wy;
        if (ws > 0) {
            wdx = wx / ws;
            wdy = wy / ws;
        }
        double wish_speed = ws < (double)Q_MAX_SPEED ? ws : (double)Q_MAX_SPEED;
        if (ws != ws)
            wish_speed = ws;
        if (was_on_ground) {                                           /* phys:83-90, all f64 */
            double speed = sqrt(vx * vx + vy * vy);
            double control = speed > (double)Q_STOP_SPEED ? speed : (double)Q_STOP_SPEED;
            double new_speed = speed - dt * control * (double)Q_FRICTION;
            if (!(new_speed > 0))
                new_speed = 0;
            if (speed > 0) {
                double ratio = new_speed / speed;
                vx = vx * ratio;
                vy = vy * ratio;
            }
        }
        double current = vx * wdx + vy * wdy;                          /* phys:69-80 */
        double clipped = (wish_speed > 30 && !was_on_ground) ? 30.0 : wish_speed;
        double add = clipped - current;
        if (!(add > 0))
            add = 0;
        double accel = (dt_f32 ? (double)(Q_ACCELERATE * dtf) : (double)Q_ACCELERATE * dt) * wish_speed;
        if (add < accel)
            accel = add;
        vx = vx + accel * wdx;
        vy = vy + accel * wdy;

        int jr = (jr_in[i] != 0) | !jump;                              /* phys:112-132 */
        int do_jump = was_on_ground && jump && jr;
        vz = vz + (do_jump ? (double)Q_JUMP_SPEED : 0.0);
        vz = vz - (dt_f32 ? (double)(Q_GRAVITY * dtf) : (double)Q_GRAVITY * dt);
        double z = z_in[i] + dt * vz;
        int og = z < (double)Q_FLOOR_HEIGHT;
        z_out[i] = og ? (double)Q_FLOOR_HEIGHT : z;
        vel_out[3 * i] = vx;
        vel_out[3 * i + 1] = vy;
        vel_out[3 * i + 2] = og ? 0.0 : vz;
        og_out[i] = (uint8_t)og;
        jr_out[i] = (uint8_t)jr;
    }
}

/* ------------------------------------------------------------------ action decode (env.py) -- */

typedef struct {
    double yaw;
    int64_t smove, fmove;
    int jump;
} command_t;

/* env:225-269 for one env.  key[] holds bit 0 of the int-truncated key actions: with last_keys in
 * {0,1} (it starts False, env:279) `key_actions & (elapsed | last_keys)` only ever sees bit 0. */
static void decode_one(const q1o_config *cfg, int nk, uint8_t *last_keys, double *last_press,
                       double *yaw_state, const uint8_t *key, double mouse, float z_vel,
                       double time_remaining, command_t *out)
{
    /* env:230: np.float32(720) * python float -> f32 under NEP 50 (NumPy 2, the default here); with
     * cfg->reserved & 1 the float64 product NumPy < 2 forms (the reference pins numpy 1.18.2) */
    double max_yaw_delta = (cfg->reserved & 1) ? 720.0 * cfg->time_delta
                                               : (double)(Q_MAX_YAW_SPEED * (float)cfg->time_delta);
    double mouse_x;
    if (!cfg->allow_yaw)
        mouse_x = 0.;                                                            /* env:234 */
    else if (cfg->discrete_yaw_steps == -1)
        mouse_x = mouse * max_yaw_delta / cfg->action_range;                     /* env:236 */
    else
        mouse_x = (mouse - cfg->discrete_yaw_steps) * max_yaw_delta
                  / cfg->discrete_yaw_steps;                                     /* env:238 */

    double now = cfg->time_limit - time_remaining;
    double smoothed[4] = { 0, 0, 0, 0 };
    int pressed[4] = { 0, 0, 0, 0 };
    for (int k = 0; k < nk; k++) {
        int elapsed = now >= last_press[k] + cfg->key_press_delay;               /* env:241-242 */
        int down = (key[k] & 1) & (elapsed | last_keys[k]);                      /* env:243 */
        if (down & ~last_keys[k] & 1)
            last_press[k] = now;                                                 /* env:244-248 */
        smoothed[k] = cfg->smooth_keys ? (down + last_keys[k]) * 0.5 : (double)down; /* env:251-254 */
        last_keys[k] = (uint8_t)down;                                            /* env:256 */
        pressed[k] = down;
    }

    *yaw_state = *yaw_state + mouse_x;                                           /* env:258 */
    double strafe = smoothed[KEY_RIGHT] - smoothed[KEY_LEFT];                    /* env:259 */
    out->yaw = *yaw_state;
    out->smove = (int64_t)((double)(float)cfg->smove_max * strafe);              /* env:260, 269 */
    out->fmove = (int64_t)((double)(float)cfg->fmove_max * smoothed[KEY_FORWARD]); /* env:261, 269 */
    if (cfg->auto_jump)
        out->jump = z_vel <= 16;                                                 /* env:263 */
    else if (cfg->allow_jump)
        out->jump = pressed[KEY_JUMP];                                           /* env:265 */
    else
        out->jump = 0;                                                           /* env:267 */
}

void q1o_decode(const q1o_config *cfg, int64_t n,
                uint8_t *last_keys, double *last_press, double *yaw,
                const uint8_t *keys, const double *mouse,
                const float *z_vel, const double *time_remaining,
                double *yaw_out, int64_t *smove, int64_t *fmove, uint8_t *jump)
{
    int nk = q1o_num_keys(cfg);
    for (int64_t i = 0; i < n; i++) {
        command_t c;
        decode_one(cfg, nk, last_keys + i * nk, last_press + i * nk, yaw + i, keys + i * nk,
                   mouse ? mouse[i] : 0.0, z_vel[i], time_remaining[i], &c);
        yaw_out[i] = c.yaw;
        smove[i] = c.smove;
        fmove[i] = c.fmove;
        jump[i] = (uint8_t)c.jump;
    }
}

/* ------------------------------------------------------------------ observation ------------- */

static void observe_one(const q1o_config *cfg, const q1o_state *st, int64_t i, double *o)
{
    /* env:381-400: velocity truncated to multiples of 16 (f32 divide, int cast), origin rounded
     * half-to-even to eighths, then divided by get_obs_scale (env:294-296). */
    o[0] = st->time_remaining[i] / cfg->time_limit;
    o[1] = st->yaw[i] / 90.;
    o[2] = (rint(st->z_pos[i] * 8) / 8) / 100.;
    for (int k = 0; k < 3; k++) {
        int64_t q = (int64_t)(st->vel[3 * i + k] / 16.0f) * 16;
        o[3 + k] = (double)q / 200.;
    }
}

void q1o_observe(const q1o_config *cfg, int64_t n, const q1o_state *st, double *obs)
{
    for (int64_t i = 0; i < n; i++)
        observe_one(cfg, st, i, obs + 6 * i);
}

/* ------------------------------------------------------------------ the env tick ------------ */

void q1o_step(const q1o_config *cfg, int64_t n, q1o_state *st,
              const uint8_t *keys, const double *mouse,
              double *obs, float *reward, uint8_t *done)
{
    int nk = q1o_num_keys(cfg);
    for (int64_t i = 0; i < n; i++) {
        if (cfg->hover) {                                                       /* env:483-485 */
            st->vel[3 * i + 2] = 0.0f;
            st->z_pos[i] = 100.0;
        }

        command_t cmd;                                                          /* env:487-488 */
        decode_one(cfg, nk, st->last_keys + i * nk, st->last_press + i * nk, st->yaw + i,
                   keys + i * nk, mouse ? mouse[i] : 0.0, st->vel[3 * i + 2],
                   st->time_remaining[i], &cmd);

        /* env:490-498; pitch = roll = 0 so _angle_vectors is [[cy, sy], [sy, -cy]] (phys:65-66) */
        double a = cmd.yaw * M_PI / 180.;
        double sy = sin(a), cy = cos(a);
        body_t b = { st->z_pos[i], st->vel[3 * i], st->vel[3 * i + 1], st->vel[3 * i + 2],
                     st->on_ground[i] != 0, st->jump_released[i] != 0 };
        move_body(&b, cy, sy, sy, -cy, (double)cmd.fmove, (double)cmd.smove, cmd.jump,
                  cfg->time_delta, 0);
        st->z_pos[i] = b.z;
        st->vel[3 * i] = b.vx;
        st->vel[3 * i + 1] = b.vy;
        st->vel[3 * i + 2] = b.vz;
        st->on_ground[i] = (uint8_t)b.on_ground;
        st->jump_released[i] = (uint8_t)b.jump_released;

        /* env:500-503: python float * f32 array stays f32 */
        float dtf = (float)cfg->time_delta;
        if (cfg->speed_reward)
            reward[i] = dtf * sqrtf(b.vx * b.vx + b.vy * b.vy);
        else
            reward[i] = dtf * b.vy;

        st->time_remaining[i] -= cfg->time_delta;                               /* env:505 */
        done[i] = st->time_remaining[i] < 0;                                    /* env:506 */
        observe_one(cfg, st, i, obs + 6 * i);                                   /* env:510 */
    }
}

/* ------------------------------------------------------------------ reset ------------------- */

void q1o_reset_env(const q1o_config *cfg, q1o_state *st, int64_t i, const double u[5])
{
    int nk = q1o_num_keys(cfg);
    /* env:458-459 / env:429-430 */
    st->z_pos[i] = (double)Q_INITIAL_Z;
    st->vel[3 * i + 2] = Q_INITIAL_VZ;
    st->on_ground[i] = 0;
    st->jump_released[i] = 1;

    /* env:461-471.  np.random.uniform(x) is uniform(low=x, high=1.0) = x + (1 - x) * u. */
    int zs = u[0] < cfg->zero_start_prob;
    st->zero_start[i] = (uint8_t)zs;
    st->yaw[i] = zs ? (double)Q_INITIAL_YAW_ZERO
                    : cfg->initial_yaw_lo + (cfg->initial_yaw_hi - cfg->initial_yaw_lo) * u[1];
    st->time_remaining[i] = zs ? cfg->time_limit : cfg->time_limit + (1.0 - cfg->time_limit) * u[2];
    double speed = zs ? 0.0 : cfg->max_initial_speed + (1.0 - cfg->max_initial_speed) * u[3];
    double two_pi = 2 * M_PI;
    double angle = two_pi + (1.0 - two_pi) * u[4];
    if (cfg->hover) {
        speed = 320;
        angle = M_PI / 2;
    }
    st->vel[3 * i] = (float)(speed * cos(angle));                               /* env:475 */
    st->vel[3 * i + 1] = (float)(speed * sin(angle));                           /* env:476 */

    /* env:283-291 */
    for (int k = 0; k < nk; k++) {
        st->last_press[i * nk + k] = -cfg->key_press_delay;
        st->last_keys[i * nk + k] = 0;
    }
}

/* ------------------------------------------------------------------ shared RNG / policies --- */

void q1o_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double unit53(uint32_t a, uint32_t b)
{
    /* the legacy NumPy random_sample construction: 27 + 26 bits */
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

void q1o_reset_draws(uint64_t seed, uint64_t env_index, uint32_t epoch, double u[5])
{
    uint32_t w[12];
    for (uint32_t j = 0; j < 3; j++)
        q1o_philox4x32((uint32_t)env_index, (uint32_t)(env_index >> 32), epoch, 0x52455300u + j,
                       (uint32_t)seed, (uint32_t)(seed >> 32), w + 4 * j);
    for (int j = 0; j < 5; j++)
        u[j] = unit53(w[2 * j], w[2 * j + 1]);
}

void q1o_policy_action(const q1o_config *cfg, int32_t policy, uint64_t seed, uint64_t env_index,
                       uint32_t tick, uint8_t *keys, double *mouse)
{
    int nk = q1o_num_keys(cfg);
    float max_yaw_delta = Q_MAX_YAW_SPEED * (float)cfg->time_delta;
    if (policy == 0) {
        /* uniform random keys + mouse */
        uint32_t w[4];
        q1o_philox4x32((uint32_t)env_index, (uint32_t)(env_index >> 32), tick, 0x41435400u,
                       (uint32_t)seed, (uint32_t)(seed >> 32), w);
        for (int k = 0; k < nk; k++)
            keys[k] = (uint8_t)((w[0] >> k) & 1u);
        if (cfg->discrete_yaw_steps == -1) {
            double unit = (double)w[1] * (1.0 / 4294967296.0);
            *mouse = (double)(float)(-cfg->action_range + 2.0 * cfg->action_range * unit);
        } else {
            *mouse = (double)(w[1] % (uint32_t)(2 * cfg->discrete_yaw_steps + 1));
        }
    } else {
        /* scripted strafe-jump: hold forward, swap strafe side every 36 ticks while turning
         * 1.5 degrees per tick into the strafe, tap jump on odd ticks. */
        uint32_t phase = ((tick + (uint32_t)(env_index % 72u)) / 36u) & 1u;
        keys[KEY_LEFT] = (uint8_t)(phase == 0);
        keys[KEY_RIGHT] = (uint8_t)(phase == 1);
        keys[KEY_FORWARD] = 1;
        if (nk == 4)
            keys[KEY_JUMP] = (uint8_t)(tick & 1u);
        double turn = phase == 0 ? 1.5 : -1.5;
        if (cfg->discrete_yaw_steps == -1) {
            *mouse = (double)(float)(turn * cfg->action_range / (double)max_yaw_delta);
        } else {
            int steps = cfg->discrete_yaw_steps;
            *mouse = (double)(steps + (phase == 0 ? 1 : -1) * ((steps + 3) / 4));
        }
    }
}

void q1o_policy_actions(const q1o_config *cfg, int32_t policy, uint64_t seed,
                        uint64_t env_index_base, int64_t n, uint32_t tick,
                        uint8_t *keys, double *mouse)
{
    int nk = q1o_num_keys(cfg);
    for (int64_t i = 0; i < n; i++)
        q1o_policy_action(cfg, policy, seed, env_index_base + (uint64_t)i, tick, keys + i * nk,
                          mouse + i);
}

void q1o_reset_philox(const q1o_config *cfg, q1o_state *st, int64_t n, uint64_t seed,
                      uint64_t env_index_base, uint32_t epoch, const uint8_t *mask)
{
    for (int64_t i = 0; i < n; i++) {
        if (mask && !mask[i])
            continue;
        double u[5];
        q1o_reset_draws(seed, env_index_base + (uint64_t)i, epoch, u);
        q1o_reset_env(cfg, st, i, u);
    }
}

/* np.sin / np.cos on float64 (phys.py:58-59): NumPy forwards to the C library.  Checker for the
 * device build of q1_libm_sincos.cuh. */
void q1o_sincos(int64_t n, const double *x, double *s, double *c)
{
    for (int64_t i = 0; i < n; i++) {
        s[i] = sin(x[i]);
        c[i] = cos(x[i]);
    }
}
