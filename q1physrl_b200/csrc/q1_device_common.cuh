/*
 * q1_device_common.cuh -- device helpers shared by the kernels of q1phys.cu and q1_actor.cu: access to
 * the tile-contiguous state blocks, observation / action array access, the episode-metric reduction
 * and the per-tick record rows of q1_record_view.
 */
#pragma once

#include "../../include/q1phys.h"
#include "q1_tick.cuh"

namespace q1 {

__device__ __forceinline__ unsigned char *state_block(const Params &P, int64_t i)
{
    return P.state + (i / kTile) * kTileBytes;
}

template <bool STAMPS>
__device__ __forceinline__ void load_env(const Params &P, int64_t i, Env &e)
{
    const unsigned char *blk = state_block(P, i);
    const int l = (int)(i % kTile);
    const float4 a = reinterpret_cast<const float4 *>(blk + kTileRecA)[l];
    const double2 b = reinterpret_cast<const double2 *>(blk + kTileRecB)[l];
    e.vx = a.x;
    e.vy = a.y;
    e.vz = a.z;
    e.bits = __float_as_uint(a.w);
    e.z = b.x;
    e.yaw = b.y;
    e.trem = reinterpret_cast<const double *>(blk + kTileTrem)[l];
    if (STAMPS) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            e.stamp[k] = k < P.num_keys ? P.stamps[(int64_t)k * P.n + i] : 0.0;
    }
}

template <bool STAMPS>
__device__ __forceinline__ void store_env(const Params &P, int64_t i, const Env &e)
{
    unsigned char *blk = state_block(P, i);
    const int l = (int)(i % kTile);
    reinterpret_cast<float4 *>(blk + kTileRecA)[l] =
        make_float4(e.vx, e.vy, e.vz, __uint_as_float(e.bits));
    reinterpret_cast<double2 *>(blk + kTileRecB)[l] = make_double2(e.z, e.yaw);
    reinterpret_cast<double *>(blk + kTileTrem)[l] = e.trem;
    if (STAMPS) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < P.num_keys)
                P.stamps[(int64_t)k * P.n + i] = e.stamp[k];
    }
}

__device__ __forceinline__ void store_obs(float *obs, int64_t i, const float o[6])
{
    float *row = obs + 6 * i;
    if ((reinterpret_cast<uintptr_t>(obs) & 7u) == 0) {
        float2 *r2 = reinterpret_cast<float2 *>(row);
        r2[0] = make_float2(o[0], o[1]);
        r2[1] = make_float2(o[2], o[3]);
        r2[2] = make_float2(o[4], o[5]);
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++)
            row[k] = o[k];
    }
}

/* (n, nk) u8 key actions -> bit mask; bit 0 of each byte is the action (env:228 astype(int), and
 * last_keys in {0,1} means `&` only ever sees bit 0). */
__device__ __forceinline__ uint32_t load_keys(const uint8_t *keys, int64_t i, int nk)
{
    if (nk == 4) {
        if ((reinterpret_cast<uintptr_t>(keys) & 3u) == 0) {
            uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(keys) + i);
            return (w & 1u) | ((w >> 7) & 2u) | ((w >> 14) & 4u) | ((w >> 21) & 8u);
        }
        const uint8_t *k = keys + 4 * i;
        return (__ldg(k) & 1u) | ((__ldg(k + 1) & 1u) << 1) | ((__ldg(k + 2) & 1u) << 2) |
               ((__ldg(k + 3) & 1u) << 3);
    }
    const uint8_t *k = keys + 3 * i;
    return (__ldg(k) & 1u) | ((__ldg(k + 1) & 1u) << 1) | ((__ldg(k + 2) & 1u) << 2);
}

__device__ __forceinline__ double load_mouse(const void *mouse, int kind, int64_t i)
{
    if (kind == Q1_MOUSE_F32)
        return (double)__ldg(reinterpret_cast<const float *>(mouse) + i);
    if (kind == Q1_MOUSE_I32)
        return (double)__ldg(reinterpret_cast<const int32_t *>(mouse) + i);
    return __ldg(reinterpret_cast<const double *>(mouse) + i);
}

/* -- episode metrics (q1physrl/train.py:54-57, 67-71) ------------------------------------------ */

__device__ __forceinline__ unsigned long long ordered_bits(double v)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

/* Called by full warps.  `finished`: this lane's episode ended this tick with return `ret`.
 * Warp-ballot first: a warp without a finished episode leaves after one instruction. */
__device__ __forceinline__ void report_episodes(const Params &P, bool finished, bool zs, double ret)
{
    unsigned any = __ballot_sync(0xffffffffu, finished);
    if (!any)
        return;
    double s = finished ? ret : 0.0, zsum = (finished && zs) ? ret : 0.0;
    double mx = finished ? ret : -INFINITY;
    int cnt = __popc(any), zcnt = __popc(__ballot_sync(0xffffffffu, finished && zs));
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        zsum += __shfl_xor_sync(0xffffffffu, zsum, off);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0) {
        if (zcnt) {
            atomicAdd(&P.metrics[0], zsum);
            atomicAdd(reinterpret_cast<unsigned long long *>(&P.metrics[1]), (unsigned long long)zcnt);
        }
        atomicAdd(&P.metrics[2], s);
        atomicAdd(reinterpret_cast<unsigned long long *>(&P.metrics[3]), (unsigned long long)cnt);
        atomicMax(reinterpret_cast<unsigned long long *>(&P.metrics[4]), ordered_bits(mx));
    }
}


/* -- per-tick record rows (q1_record_view; q1physrl/analyse.py:214-229) --------------------------- */

/* What analyse.py appends BEFORE the tick: movement state (analyse.py:218), observation the policy saw
 * (analyse.py:219; the hover override, env:483-485, happens inside the tick, after it) and the action
 * (analyse.py:220).  `row` = tick * n + env. */
__device__ __forceinline__ void record_before(const q1_record_view &rec, int64_t row, int nk, const Env &e,
                                              const float o[6], uint32_t keybits, double mouse)
{
    if (rec.vel) {
        rec.vel[3 * row] = e.vx;
        rec.vel[3 * row + 1] = e.vy;
        rec.vel[3 * row + 2] = e.vz;
    }
    if (rec.z_pos)
        rec.z_pos[row] = e.z;
    if (rec.on_ground)
        rec.on_ground[row] = (e.bits & F_ON_GROUND) != 0;
    if (rec.jump_released)
        rec.jump_released[row] = (e.bits & F_JUMP_RELEASED) != 0;
    if (rec.time_remaining)
        rec.time_remaining[row] = e.trem;
    if (rec.obs)
        store_obs(rec.obs, row, o);
    if (rec.keys)
        for (int k = 0; k < nk; k++)
            rec.keys[row * nk + k] = (keybits >> k) & 1u;
    if (rec.mouse)
        rec.mouse[row] = (float)mouse;
}

/* ... and AFTER it: the move command ActionDecoder.map made of the action (analyse.py:215-216, 221-224),
 * reward and done (analyse.py:226-228). */
__device__ __forceinline__ void record_after(const q1_record_view &rec, int64_t row, uint32_t record_flags,
                                             int jump_mode, const Move &mv, const float o[6], float reward,
                                             bool done)
{
    if (rec.yaw)
        rec.yaw[row] = mv.yaw;
    if (rec.smove)
        rec.smove[row] = (int64_t)mv.smove;
    if (rec.fmove)
        rec.fmove[row] = (int64_t)mv.fmove;
    if (rec.jump) {
        /* analyse.py:215-216 hands its shadow decoder the OBSERVATION's z velocity (quantised, divided
         * by 200), so with auto_jump it records obs <= 16 */
        bool j = mv.jump;
        if ((record_flags & Q1_RECORD_SHADOW_JUMP) && jump_mode == 2)
            j = o[5] <= 16.0f;
        rec.jump[row] = j;
    }
    if (rec.reward)
        rec.reward[row] = reward;
    if (rec.done)
        rec.done[row] = done;
}

} // namespace q1
