/*
 * q1_tick.cuh -- device code of one q1physrl_env movement tick for ONE env held in registers.
 *
 * Everything here is written for sm_100a and mirrors the widths the reference executes under
 * NumPy 2: f32 velocity, f64 z / yaw / time_remaining, f64 intermediates
 * with a single rounding to f32 at the store, f32 friction speed.  All f64/f32 arithmetic goes
 * through the *_rn intrinsics so that no multiply-add is ever contracted into an FMA, whatever the
 * compiler flags are (the reference never fuses: einsum and ufuncs round after every operation).
 *
 * Citations: phys = q1physrl_env/q1physrl_env/phys.py, env = q1physrl_env/q1physrl_env/env.py.
 */
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "q1_libm_sincos.cuh"

/* 1: the branch-free polynomial sincos_rad() (< 1 ulp, NOT bit-identical to the reference's libm)
 * in the lean kernels, for A/B timing only; 0 (default): glibc's algorithm, bit for bit. */
#ifndef Q1_POLY_SINCOS
#define Q1_POLY_SINCOS 0
#endif
#ifndef Q1_TICK_LIBDEVICE_HUGE
#define Q1_TICK_LIBDEVICE_HUGE 0
#endif

namespace q1 {

/* ------------------------------------------------------------------ constants ---------------- */

/* phys:47-53 (np.float32 scalars, all exactly representable). */
constexpr float kMaxSpeed = 320.0f;
constexpr float kStopSpeed = 100.0f;
constexpr float kFriction = 4.0f;
constexpr float kJumpSpeed = 270.0f;
constexpr float kFloorHeight = 24.03125f;
/* env:54-58 */
constexpr float kInitialZ = 32.843201f;
constexpr float kInitialVz = -12.0f;
constexpr float kInitialYawZero = 90.0f;

constexpr double kPi = 3.14159265358979323846; /* np.pi */

/* The per-env "bits" word of the persistent state:
 *   bits  0..19  four 5-bit "ticks until the key may be pressed again" counters (counter mode)
 *   bits 20..27  flags */
enum : uint32_t {
    TIMER_BITS = 5u,
    TIMER_MAX = 31u,
    TIMER_FIELD_MASK = 0xFFFFFu,
    TIMER_LOW = 0x08421u,       /* bit 0 of each 5-bit field */
    TIMER_HIGH = 0x84210u,      /* bit 4 of each 5-bit field */
    F_SHIFT = 20u,
    F_ON_GROUND = 1u << 20,
    F_JUMP_RELEASED = 1u << 21,
    F_ZERO_START = 1u << 22,
    F_LAST_KEY_SHIFT = 23u,     /* bits 23..26 = last_keys[0..3] (env:201) */
    F_DONE_SEEN = 1u << 27      /* episode already reported to the metrics (time_remaining crossed 0) */
};

enum : int { KEY_LEFT = 0, KEY_RIGHT = 1, KEY_FORWARD = 2, KEY_JUMP = 3 }; /* env:61-73 */

/* Per-config constants.  Passed to every kernel as a __grid_constant__ parameter: they live in
 * the constant bank and reach the FP64 pipe as uniform operands (no shared-memory round trip). */
struct Params {
    int64_t n;
    uint64_t env_index_base;
    uint64_t seed;
    /* env.Config numbers in the width they are used */
    double dt;              /* time_delta */
    double time_limit;
    double key_delay;       /* key_press_delay */
    double max_yaw_delta;   /* f64(f32(720) * f32(dt))           env:230 */
    double action_range;    /* env:236 */
    double yaw_steps;       /* f64(discrete_yaw_steps)           env:238 */
    double accel_dt;        /* f64(f32(10)) * dt                 phys:78 */
    double gravity_dt;      /* f64(f32(800)) * dt                phys:122 */
    double fmove_half, fmove_full; /* trunc(f64(f32(fmove_max)) * {0.5, 1})   env:261,269 */
    double smove_half, smove_full; /* trunc(f64(f32(smove_max)) * {0.5, 1})   env:260,269 */
    double zero_start_prob, yaw_lo, yaw_hi, max_initial_speed;
    /* RN(1/b) of the constant divisors, for div_const() */
    double rcp_action_range, rcp_yaw_steps, rcp_time_limit;
    double fmove_tab[3];    /* fmove for twice-the-smoothed-forward-key = 0, 1, 2 */
    double smove_tab[5];    /* smove for twice-(right - left) + 2 = 0 .. 4 */
    float dt_f32;           /* f32(time_delta) for the reward      env:500-503 */
    int32_t num_keys;       /* env:206-207 */
    int32_t delay_ticks;    /* D = ceil(key_delay / dt) (counter mode) */
    int32_t press_ticks;    /* max(D - 1, 0): counter value stored on a key press */
    int32_t jump_mode;      /* 0: never (env:267)  1: jump key (env:265)  2: auto jump (env:263) */
    int32_t allow_yaw, discrete_yaw, speed_reward, hover, smooth_keys, auto_jump, allow_jump;
    int32_t ieee_div;       /* use the IEEE division intrinsics instead of the reciprocal sequences */
    /* persistent state, tile-contiguous: env i lives in block i / 128 of kTileBytes bytes,
     * [float4 rec_a[128] {vx, vy, vz, bits} | double2 rec_b[128] {z_pos, yaw} | double trem[128]],
     * so the state of a tile of 128 envs is one contiguous 5 KB bulk copy */
    unsigned char *state;
    double *stamps;         /* stamp mode: (num_keys, n) f64 last key press time (env:200) */
    uint32_t *epoch;        /* reset count per env: RNG stream position, touched by resets only */
    double *ep_return;      /* TRACK only: running f64 episode return */
    double *metrics;        /* TRACK only: [zs_sum, zs_count, sum, count, max-as-ordered-bits] */
};

constexpr int kTile = 128;                          /* envs per state block / per CTA tile */
constexpr int kTileRecA = 0;                        /* byte offsets inside a state block */
constexpr int kTileRecB = 16 * kTile;
constexpr int kTileTrem = 32 * kTile;
constexpr int kTileBytes = 40 * kTile;

/* One env in registers. */
struct Env {
    float vx, vy, vz;
    uint32_t bits;
    double z, yaw, trem;
    double stamp[4];
};

/* ------------------------------------------------------------------ rounding-exact helpers --- */

__device__ __forceinline__ double mul64(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add64(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub64(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div64(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float mul32(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add32(float a, float b) { return __fadd_rn(a, b); }

__device__ __forceinline__ double fma64(double a, double b, double c) { return __fma_rn(a, b, c); }

/* Correctly rounded a / b for a constant divisor b with y = RN(1 / b) precomputed on the host.
 * Markstein's theorem: with y the correctly rounded reciprocal and q a faithful quotient,
 * RN(q + RN(a - b q) y) = RN(a / b).  The first correction makes q faithful (RN(a y) alone can be
 * 2 ulp off), the second makes it exact.  Five FP64-pipe operations, no branches, no special-case
 * path: valid for finite a with |a / b| comfortably inside the normal range (|a| in [2^-900, 2^900]
 * or zero), which covers everything this path divides. */
__device__ __forceinline__ double div_const(double a, double b, double y)
{
    double q = __dmul_rn(a, y);
    double r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-b, q, a);
    return __fma_rn(r, y, q);
}

/* The same quotient in three operations, for divisors whose rounded reciprocal is good enough:
 * if |y b - 1| <= 2^-54 then q = RN(a y) is already faithful (|q - a/b| < |a/b| 2^-54 + ulp/2
 * < 1 ulp, the first term being below half an ulp because |a/b| < 2^(e+1) strictly), and one
 * Markstein correction rounds correctly.  180 and 90 qualify (|y b - 1| = 0.34 * 2^-53), so do
 * 10, 5 and f32(10.08); a handle whose action_range / yaw steps / time_limit do not is given the
 * IEEE-division kernels instead (short_division_ok() in q1phys.cu).  q1_selftest_division checks
 * this form against __ddiv_rn for every qualifying constant it draws. */
__device__ __forceinline__ double div_const3(double a, double b, double y)
{
    double q = __dmul_rn(a, y);
    double r = __fma_rn(-b, q, a);
    return __fma_rn(r, y, q);
}
/* device-side evaluation of the criterion (y b - 1 is exact in one fused multiply-add) */
__device__ __forceinline__ bool short_division_ok_dev(double b, double y)
{
    return fabs(__fma_rn(y, b, -1.0)) <= 0x1p-54;
}

/* RN(1 / b) for a normal b well inside the exponent range: hardware seed (2^-23), one cubic and
 * one Markstein step.  q1_selftest_division checks this bit for bit against __drcp_rn / __ddiv_rn
 * on 10^10 random operands, one in 64 of them with a significand of all ones. */
__device__ __forceinline__ double rcp_rn(double b)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    double e = __fma_rn(-b, y, 1.0);
    double t = __fma_rn(e, e, e);
    y = __fma_rn(y, t, y);            /* relative error ~2^-69 before rounding */
    e = __fma_rn(-b, y, 1.0);
    y = __fma_rn(y, e, y);            /* RN(1 / b) ... */
    /* ... except for a significand of all ones, where the last step lands exactly on a tie
     * (Markstein's one exception, probability 2^-52): take the IEEE intrinsic there. */
    const bool all_ones = (__double2loint(b) == -1) & ((__double2hiint(b) & 0xFFFFF) == 0xFFFFF);
    if (__builtin_expect(all_ones, 0))
        y = __drcp_rn(b);
    return y;
}

/* a / b given y = RN(1 / b) (same theorem as div_const). */
__device__ __forceinline__ double div_rcp(double a, double b, double y) { return div_const(a, b, y); }

/* f32 quotient a / b for the two observation divisors (b = 200 with a a multiple of 16, b = 100
 * with a a multiple of 1/8; y = RN32(1 / b)).  One residual correction is exact on those whole
 * domains (|a / 16| <= 2^20, |8 a| <= 2^24): q1_selftest_division checks every value against the
 * reference's f64 division rounded to f32. */
__device__ __forceinline__ float div_const32(float a, float b, float y)
{
    float q = __fmul_rn(a, y);
    float r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, y, q);
}
/* two corrections (Markstein, as div_const): the self-test reports it beside the short form */
__device__ __forceinline__ float div_const32_long(float a, float b, float y)
{
    float q = __fmul_rn(a, y);
    float r = __fmaf_rn(-b, q, a);
    q = __fmaf_rn(r, y, q);
    r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, y, q);
}

/* ------------------------------------------------------------------ Philox4x32-10 ------------ */

__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 53-bit uniform in [0,1) from two words: the 27+26 bit construction of NumPy's random_sample. */
__device__ __forceinline__ double unit53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

/* ------------------------------------------------------------------ sin / cos ---------------- */

/* Minimax coefficients of sin(r) = r + r^3 S(r^2) and cos(r) = 1 - r^2/2 + r^4 C(r^2) on
 * |r| <= pi/4 (the fdlibm k_sin / k_cos sets), kept in the constant bank so the FP64 pipe reads
 * them as operands instead of building each 64-bit immediate from two moves. */
/* Other 64-bit constants of the tick, read from the constant bank as well. */
static __constant__ double kTickConst[10] = {
    3.14159265358979323846,       /* 0: np.pi */
    1.0 / 180.0,                  /* 1 */
    1.0 / 90.0,                   /* 2 */
    6.36619772367581382433e-01,   /* 3: 2 / pi */
    6755399441055744.0,           /* 4: 1.5 * 2^52 */
    -1.57079632679489655800e+00,  /* 5: -pi/2 high */
    -6.12323399573676603587e-17,  /* 6: -pi/2 middle */
    1.4973849048591698e-33,       /* 7: -pi/2 low (the third term is negative) */
    1.0e5,                        /* 8: sincos fast-path bound */
    0.0};
static __constant__ double kSinCosCoef[12] = {
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};

/* sin and cos of a (radians), < 1 ulp like libdevice's sincos but branch-free for |a| < 1e5 (yaw
 * below 5.7 million degrees): quadrant by the round-to-nearest magic add, three-term Cody-Waite
 * reduction with fused multiply-adds, two Horner chains.  Larger arguments take libdevice's path. */
__device__ __forceinline__ void sincos_rad(double a, double &s, double &c)
{
    if (__builtin_expect(!(fabs(a) < kTickConst[8]), 0)) {
        sincos(a, &s, &c);
        return;
    }
    const double magic = kTickConst[4]; /* 1.5 * 2^52 */
    double t = __fma_rn(a, kTickConst[3], magic);
    const int q = __double2loint(t);
    const double j = __dsub_rn(t, magic);
    double r = __fma_rn(j, kTickConst[5], a);
    r = __fma_rn(j, kTickConst[6], r);
    r = __fma_rn(j, kTickConst[7], r); /* pi/2 = 1.5707963267948966 + 6.123233995736766e-17 - 1.4973849048591698e-33 */
    const double z = __dmul_rn(r, r);
    double ps = kSinCosCoef[5];
    ps = __fma_rn(ps, z, kSinCosCoef[4]);
    ps = __fma_rn(ps, z, kSinCosCoef[3]);
    ps = __fma_rn(ps, z, kSinCosCoef[2]);
    ps = __fma_rn(ps, z, kSinCosCoef[1]);
    ps = __fma_rn(ps, z, kSinCosCoef[0]);
    double pc = kSinCosCoef[11];
    pc = __fma_rn(pc, z, kSinCosCoef[10]);
    pc = __fma_rn(pc, z, kSinCosCoef[9]);
    pc = __fma_rn(pc, z, kSinCosCoef[8]);
    pc = __fma_rn(pc, z, kSinCosCoef[7]);
    pc = __fma_rn(pc, z, kSinCosCoef[6]);
    const double sr = __fma_rn(__dmul_rn(r, z), ps, r);
    const double cr = __fma_rn(z, __fma_rn(z, pc, -0.5), 1.0);
    /* quadrant: q odd swaps, bit 1 negates sin (after the swap), bits 0^1 negate cos */
    double ss = (q & 1) ? cr : sr;
    double cc = (q & 1) ? sr : cr;
    const int sflip = (q & 2) << 30, cflip = ((q + 1) & 2) << 30;
    s = __hiloint2double(__double2hiint(ss) ^ sflip, __double2loint(ss));
    c = __hiloint2double(__double2hiint(cc) ^ cflip, __double2loint(cc));
}

/* sin and cos of a (radians) with the bits np.sin / np.cos return in the reference (phys:58-59,
 * env:475-476): glibc's __sin / __cos restated in q1_libm_sincos.cuh, exact for every finite
 * argument everywhere.  HUGE = true (phys.apply, the sweep, q1_sincos_host) carries __branred
 * inline; HUGE = false (the per-tick kernels, resets) keeps the hot path to |a| < 105414350 (yaw
 * below 6e9 degrees) and hands anything beyond to sincos_cold(), one out-of-line copy of the
 * general routine: a call that a lockstep episode never makes, but the one that keeps "bit-identical
 * for every finite input" free of footnotes.  Infinities and NaN give NaN like libm. */
static __device__ __noinline__ void sincos_cold(double a, double *s, double *c)
{
    double sv, cv;
    if (!q1libm::sincos<true>(a, sv, cv))
        sincos(a, &sv, &cv);   /* inf / NaN only */
    *s = sv;
    *c = cv;
}
template <bool HUGE = true>
__device__ __forceinline__ void sincos_ref(double a, double &s, double &c)
{
    if (__builtin_expect(!q1libm::sincos<HUGE>(a, s, c), 0)) {
#if Q1_TICK_LIBDEVICE_HUGE
        sincos(a, &s, &c);         /* A/B timing only: libdevice beyond 1.05e8 (not bit-exact there) */
#else
        if (HUGE)
            sincos(a, &s, &c);     /* inf / NaN only */
        else
            sincos_cold(a, &s, &c);
#endif
    }
}

/* ------------------------------------------------------------------ observation -------------- */

/* env:381-400 with get_obs_scale (env:294-296), result cast to f32 (the dtype the env declares,
 * env:416-417).  The reference divides in f64; time and yaw do exactly that.  For the three
 * velocity entries and z the numerator is a small dyadic rational (multiple of 16, of 1/8) exactly
 * representable in f32 and the divisor is 200 or 100, so the quotient is either exact or has a
 * binary expansion of period <= 20: it can never sit within 2^-53 of an f32 rounding boundary, and
 * rounding the exact quotient once to f32 equals rounding the f64 quotient to f32. */
template <bool LEAN>
__device__ __forceinline__ void observe(const Params &P, const Env &e, float o[6])
{
    if (LEAN) {
        o[0] = __double2float_rn(div_const3(e.trem, P.time_limit, P.rcp_time_limit));
        o[1] = __double2float_rn(div_const3(e.yaw, 90.0, kTickConst[2]));
    } else {
        o[0] = __double2float_rn(div64(e.trem, P.time_limit));
        o[1] = __double2float_rn(div64(e.yaw, 90.0));
    }
    const double r = rint(mul64(e.z, 8.0));          /* np.round: half to even (env:390) */
    const float qx = truncf(mul32(e.vx, 0.0625f));   /* (v / 16).astype(int): exact scaling, */
    const float qy = truncf(mul32(e.vy, 0.0625f));   /* toward zero (env:383)                */
    const float qz = truncf(mul32(e.vz, 0.0625f));
    const bool small = (fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) < 1048576.0f) &
                       (fabs(r) < 16777216.0);
    if (__builtin_expect(small, 1)) {
        /* numerators exactly representable in f32: one rounding of the exact quotient */
        const float zq = mul32((float)r, 0.125f);
        if (LEAN) {
            o[2] = div_const32(zq, 100.0f, 1.0f / 100.0f);
            o[3] = div_const32(mul32(qx, 16.0f), 200.0f, 1.0f / 200.0f);
            o[4] = div_const32(mul32(qy, 16.0f), 200.0f, 1.0f / 200.0f);
            o[5] = div_const32(mul32(qz, 16.0f), 200.0f, 1.0f / 200.0f);
        } else {
            o[2] = __fdiv_rn(zq, 100.0f);
            o[3] = __fdiv_rn(mul32(qx, 16.0f), 200.0f);
            o[4] = __fdiv_rn(mul32(qy, 16.0f), 200.0f);
            o[5] = __fdiv_rn(mul32(qz, 16.0f), 200.0f);
        }
    } else {
        /* the reference's own arithmetic: int64 * 16 -> f64, divided in f64 */
        o[2] = __double2float_rn(div64(mul64(r, 0.125), 100.0));
        o[3] = __double2float_rn(div64(mul64((double)(long long)qx, 16.0), 200.0));
        o[4] = __double2float_rn(div64(mul64((double)(long long)qy, 16.0), 200.0));
        o[5] = __double2float_rn(div64(mul64((double)(long long)qz, 16.0), 200.0));
    }
}

/* ------------------------------------------------------------------ phys.apply, one row ------ */

/* phys:184-197 for one env.  (fx, rx, fy, ry) is the 2x2 block of _angle_vectors (phys:56-66). */
/* DT32: `time_delta` arrived as a float32 array (analyse.py:110 builds it with np.full_like of an
 * f32 array).  NumPy then keeps friction, gravity and dt * z_vel in f32 (phys:87-90, 122, 127)
 * and forms 10 * dt in f32 (phys:78); with the f64 array env.vector_step passes (env:493) they
 * are f64.  Both width sets are reproduced. */
template <bool LEAN, bool DT32 = false>
__device__ __forceinline__ void move_body(float &vx, float &vy, float &vz, double &z,
                                          bool &on_ground, bool &jump_released,
                                          double fx, double rx, double fy, double ry,
                                          double fmove, double smove, bool jump, double dt,
                                          double accel_dt, double gravity_dt)
{
    const bool was_on_ground = on_ground; /* phys:191 hands _air_move the OLD flag */

    /* phys:95-103.  einsum = mul, mul, add; norm = sqrt(x*x + y*y). */
    double wx = add64(mul64(fx, fmove), mul64(rx, smove));
    double wy = add64(mul64(fy, fmove), mul64(ry, smove));
    double ws2 = add64(mul64(wx, wx), mul64(wy, wy));
    double ws, wdx = wx, wdy = wy;
    if (LEAN) {
        /* keep sqrt and the reciprocal on their branch-free fast paths: a zero (no key held) is
         * replaced by 1 for the arithmetic and selected back afterwards */
        const bool moving = ws2 > 0.0;
        double arg = moving ? ws2 : 1.0;
        asm("" : "+d"(arg)); /* keep the select in front of sqrt (else sqrt(0) takes the slow path) */
        ws = __dsqrt_rn(arg);
        double y = rcp_rn(ws);
        double qx = div_rcp(wx, ws, y), qy = div_rcp(wy, ws, y);
        ws = moving ? ws : ws2;
        wdx = moving ? qx : wx;
        wdy = moving ? qy : wy;
    } else {
        ws = __dsqrt_rn(ws2);
        if (ws > 0.0) {
            wdx = div64(wx, ws);
            wdy = div64(wy, ws);
        }
    }
    double wish_speed = ws < (double)kMaxSpeed ? ws : (double)kMaxSpeed;
    if (ws != ws)
        wish_speed = ws;
    /* phys:104-106 rescales wish_vel, which nothing reads afterwards. */

    /* phys:108, phys:83-90 */
    double hx = (double)vx, hy = (double)vy;
    if (was_on_ground) {
        float speed = __fsqrt_rn(add32(mul32(vx, vx), mul32(vy, vy)));
        float control = speed > kStopSpeed ? speed : kStopSpeed;
        if (DT32) {
            float ns = __fsub_rn(speed, mul32(mul32((float)dt, control), kFriction));
            if (!(ns > 0.0f))
                ns = 0.0f;
            if (speed > 0.0f) {
                float ratio = __fdiv_rn(ns, speed);
                hx = (double)mul32(vx, ratio);
                hy = (double)mul32(vy, ratio);
            }
        } else {
            double new_speed = sub64((double)speed, mul64(mul64(dt, (double)control), (double)kFriction));
            if (!(new_speed > 0.0))
                new_speed = 0.0;
            if (speed > 0.0f) {
                double sd = (double)speed;
                double ratio = (LEAN && speed > 1e-30f) ? div_rcp(new_speed, sd, rcp_rn(sd))
                                                        : div64(new_speed, sd);
                hx = mul64(hx, ratio);
                hy = mul64(hy, ratio);
            }
        }
    }

    /* phys:69-80 */
    double current = add64(mul64(hx, wdx), mul64(hy, wdy));
    double clipped = (wish_speed > 30.0 && !was_on_ground) ? 30.0 : wish_speed;
    double add = sub64(clipped, current);
    if (!(add > 0.0))
        add = 0.0;
    double accel = mul64(accel_dt, wish_speed);
    if (add < accel)
        accel = add;
    hx = add64(hx, mul64(accel, wdx));
    hy = add64(hy, mul64(accel, wdy));

    /* phys:190: the single f64 -> f32 rounding */
    vx = __double2float_rn(hx);
    vy = __double2float_rn(hy);

    /* phys:112-132 */
    jump_released = jump_released | !jump;
    bool do_jump = was_on_ground && jump && jump_released;
    float v = add32(vz, do_jump ? kJumpSpeed : 0.0f);
    double zn;
    if (DT32) {
        v = __fsub_rn(v, (float)gravity_dt);                 /* f32 -= f32(800) * f32(dt) */
        zn = add64(z, (double)mul32((float)dt, v));
    } else {
        v = __double2float_rn(sub64((double)v, gravity_dt));
        zn = add64(z, mul64(dt, (double)v));
    }
    bool og = zn < (double)kFloorHeight;
    z = og ? (double)kFloorHeight : zn;
    vz = og ? 0.0f : v;
    on_ground = og;
}

/* ------------------------------------------------------------------ env tick ----------------- */

/* env.VectorPhysEnv.vector_step (env:482-510) for one env: hover override, ActionDecoder.map
 * (env:225-269), phys.apply, reward, time, done.  keybits: bit k = key action k; mouse: the raw
 * mouse action as f64 (continuous value or the discrete index). */
/* COMMON: the configuration every shipped setup uses (mouse action present and continuous, no
 * hover, reward = y velocity) is compiled in; otherwise those switches are read from Params. */
/* The move command ActionDecoder.map returns (env:269), for callers that record it. */
struct Move {
    double yaw;      /* decoder yaw after this tick's mouse movement (env:258) */
    double smove, fmove;
    bool jump;
};

template <bool STAMPS, bool LEAN, bool COMMON>
__device__ __forceinline__ void tick(const Params &P, Env &e, uint32_t keybits, double mouse,
                                     float &reward, bool &done, Move *move = nullptr)
{
    const bool hover = COMMON ? false : (bool)P.hover;
    const bool allow_yaw = COMMON ? true : (bool)P.allow_yaw;
    const bool discrete_yaw = COMMON ? false : (bool)P.discrete_yaw;
    const bool speed_reward = COMMON ? false : (bool)P.speed_reward;

    if (hover) { /* env:483-485 */
        e.vz = 0.0f;
        e.z = 100.0;
    }

    /* ---- ActionDecoder.map ---- */
    double mouse_x = 0.0;
    if (allow_yaw) {
        if (!discrete_yaw) {                                                             /* env:236 */
            double t = mul64(mouse, P.max_yaw_delta);
            mouse_x = LEAN ? div_const3(t, P.action_range, P.rcp_action_range) : div64(t, P.action_range);
        } else {                                                                         /* env:238 */
            double t = mul64(sub64(mouse, P.yaw_steps), P.max_yaw_delta);
            mouse_x = LEAN ? div_const3(t, P.yaw_steps, P.rcp_yaw_steps) : div64(t, P.yaw_steps);
        }
    }

    const uint32_t key_mask = (1u << P.num_keys) - 1u;
    const uint32_t last = (e.bits >> F_LAST_KEY_SHIFT) & 0xFu;
    uint32_t elapsed;   /* bit k: key k may be pressed again (env:241-242) */
    double now = 0.0;
    if (STAMPS) {
        now = sub64(P.time_limit, e.trem);
        elapsed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < P.num_keys)
                elapsed |= (now >= add64(e.stamp[k], P.key_delay) ? 1u : 0u) << k;
    } else {
        /* four 5-bit counters "ticks this key stays blocked" in one word: gather the is-zero bit
         * of each field (= elapsed), then decrement the non-zero fields by one */
        const uint32_t t = e.bits & TIMER_FIELD_MASK;
        const uint32_t z = ~((t | ((t | TIMER_HIGH) - TIMER_LOW)) >> 4) & TIMER_LOW; /* field == 0 */
        elapsed = ((z * 0x1111u) >> 12) & 0xFu;
        e.bits = e.bits - TIMER_LOW + z;           /* fields only: no borrow can leave a field */
    }
    const uint32_t down = keybits & (elapsed | last) & key_mask;                        /* env:243 */
    const uint32_t rising = down & ~last;                                               /* env:244 */
    if (STAMPS) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((rising >> k) & 1u)
                e.stamp[k] = now;                                                        /* env:246 */
    } else {
        const uint32_t m = ((rising * 0x1111u) & TIMER_LOW) * TIMER_MAX;   /* bit k -> field k mask */
        e.bits = (e.bits & ~m) | (((uint32_t)P.press_ticks * TIMER_LOW) & m);
    }

    /* env:251-261: smoothed keys in {0, 1/2, 1}; fmove / smove are truncations of f32(max) * that,
     * i.e. one of three / five values fixed by the config (tabulated in Params).  With prev = last
     * keys (smooth_keys) or prev = keys (no smoothing) twice the smoothed key is key + prev. */
    const uint32_t prev = P.smooth_keys ? last : down;
    const uint32_t both = down | (prev << 4);
    const int f2 = __popc(both & (0x11u << KEY_FORWARD));
    const int s2 = __popc(both & (0x11u << KEY_RIGHT)) - __popc(both & (0x11u << KEY_LEFT));
    const double fmove = P.fmove_tab[f2];
    const double smove = P.smove_tab[s2 + 2];

    bool jump = false;                                                                  /* env:267 */
    if (P.jump_mode == 2)
        jump = e.vz <= 16.0f;                                                           /* env:263 */
    else if (P.jump_mode == 1)
        jump = (down >> KEY_JUMP) & 1u;                                                 /* env:265 */

    e.yaw = add64(e.yaw, mouse_x);                                                       /* env:258 */
    e.bits = (e.bits & ~(0xFu << F_LAST_KEY_SHIFT)) | (down << F_LAST_KEY_SHIFT);       /* env:256 */
    if (move) {                     /* a compile-time nullptr in the step kernels: folds away */
        move->yaw = e.yaw;
        move->smove = smove;
        move->fmove = fmove;
        move->jump = jump;
    }

    /* ---- phys.apply ---- pitch = roll = 0 (env:490-491) so the matrix is [[cy, sy], [sy, -cy]] */
    double sy, cy;
    if (LEAN) {                                                                          /* phys:58-59 */
        const double a = div_const3(mul64(e.yaw, kTickConst[0]), 180.0, kTickConst[1]);
        if (Q1_POLY_SINCOS)
            sincos_rad(a, sy, cy);
        else
            sincos_ref<false>(a, sy, cy);
    } else {
        sincos_ref<false>(div64(mul64(e.yaw, kPi), 180.0), sy, cy);
    }
    bool og = e.bits & F_ON_GROUND, jr = e.bits & F_JUMP_RELEASED;
    move_body<LEAN>(e.vx, e.vy, e.vz, e.z, og, jr, cy, sy, sy, -cy, fmove, smove, jump, P.dt,
                    P.accel_dt, P.gravity_dt);
    e.bits = (e.bits & ~(F_ON_GROUND | F_JUMP_RELEASED)) | (og ? F_ON_GROUND : 0u) |
              (jr ? F_JUMP_RELEASED : 0u);

    /* ---- reward, time, done (env:500-506) ---- */
    if (speed_reward)
        reward = mul32(P.dt_f32, __fsqrt_rn(add32(mul32(e.vx, e.vx), mul32(e.vy, e.vy))));
    else
        reward = mul32(P.dt_f32, e.vy);
    e.trem = sub64(e.trem, P.dt);
    done = e.trem < 0.0;
}

/* ------------------------------------------------------------------ reset -------------------- */

/* env.VectorPhysEnv.reset_at (env:457-480) + ActionDecoder.reset_at (env:283-291) for one env,
 * with the global np.random draws replaced by a counter-based stream: five uniforms that are a pure
 * function of (seed, global env index, reset epoch).  np.random.uniform(x) is uniform(low=x,
 * high=1.0) = x + (1 - x) * u, which this keeps. */
template <bool STAMPS>
__device__ __forceinline__ void reset_env(const Params &P, Env &e, uint64_t gidx, uint32_t epoch)
{
    uint32_t w[12];
#pragma unroll
    for (uint32_t j = 0; j < 3; j++)
        philox4x32((uint32_t)gidx, (uint32_t)(gidx >> 32), epoch, 0x52455300u + j,
                   (uint32_t)P.seed, (uint32_t)(P.seed >> 32), w + 4 * j);
    double u0 = unit53(w[0], w[1]), u1 = unit53(w[2], w[3]), u2 = unit53(w[4], w[5]);
    double u3 = unit53(w[6], w[7]), u4 = unit53(w[8], w[9]);

    bool zs = u0 < P.zero_start_prob;
    e.z = (double)kInitialZ;
    e.vz = kInitialVz;
    e.yaw = zs ? (double)kInitialYawZero : add64(P.yaw_lo, mul64(sub64(P.yaw_hi, P.yaw_lo), u1));
    e.trem = zs ? P.time_limit : add64(P.time_limit, mul64(sub64(1.0, P.time_limit), u2));
    double speed = zs ? 0.0 : add64(P.max_initial_speed, mul64(sub64(1.0, P.max_initial_speed), u3));
    const double two_pi = 2.0 * kPi;
    double angle = add64(two_pi, mul64(sub64(1.0, two_pi), u4));
    if (P.hover) {
        speed = 320.0;
        angle = kPi / 2;
    }
    double sa, ca;
    sincos_ref<false>(angle, sa, ca);   /* angle in (1, 2 pi] */
    e.vx = __double2float_rn(mul64(speed, ca));
    e.vy = __double2float_rn(mul64(speed, sa));
    e.bits = F_JUMP_RELEASED | (zs ? F_ZERO_START : 0u); /* timers 0: every key may be pressed */
#pragma unroll
    for (int k = 0; k < 4; k++)
        e.stamp[k] = -P.key_delay;
}

/* ------------------------------------------------------------------ built-in policies -------- */

/* Synthetic action streams of q1_rollout, pure functions of (policy seed, global env, tick). */
__device__ __forceinline__ void policy_action(const Params &P, int policy, uint64_t pseed,
                                              uint64_t gidx, uint32_t tick_no, uint32_t &keybits,
                                              double &mouse)
{
    if (policy == 0) {
        uint32_t w[4];
        philox4x32((uint32_t)gidx, (uint32_t)(gidx >> 32), tick_no, 0x41435400u, (uint32_t)pseed,
                   (uint32_t)(pseed >> 32), w);
        keybits = w[0] & ((1u << P.num_keys) - 1u);
        if (!P.discrete_yaw) {
            double unit = mul64((double)w[1], 1.0 / 4294967296.0);
            mouse = (double)__double2float_rn(
                add64(-P.action_range, mul64(mul64(2.0, P.action_range), unit)));
        } else {
            mouse = (double)(w[1] % (uint32_t)(2 * (int)P.yaw_steps + 1));
        }
    } else {
        uint32_t phase = ((tick_no + (uint32_t)(gidx % 72u)) / 36u) & 1u;
        keybits = (phase == 0 ? 1u << KEY_LEFT : 1u << KEY_RIGHT) | (1u << KEY_FORWARD);
        if (P.num_keys == 4)
            keybits |= (tick_no & 1u) << KEY_JUMP;
        double turn = phase == 0 ? 1.5 : -1.5;
        if (!P.discrete_yaw) {
            mouse = (double)__double2float_rn(div64(mul64(turn, P.action_range), P.max_yaw_delta));
        } else {
            int steps = (int)P.yaw_steps;
            mouse = (double)(steps + (phase == 0 ? 1 : -1) * ((steps + 3) / 4));
        }
    }
}

} // namespace q1
