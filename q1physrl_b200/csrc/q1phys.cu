/*
 * q1phys.cu -- sm_100a kernels and the C ABI (include/q1phys.h) of the q1physrl_env movement step.
 *
 * Data layout in HBM.  The state is tile-contiguous: envs are grouped in blocks of 128 and block b
 * is 5120 contiguous bytes
 *   float4  rec_a[128]  {vx, vy, vz, bits}   bits = four 5-bit key timers | on_ground,
 *                                            jump_released, zero_start, last_keys[4], done_seen
 *   double2 rec_b[128]  {z_pos, yaw}
 *   double  trem[128]   time_remaining
 * = 40 B per env, read once and written once per tick as ONE bulk copy per tile each way.  Beside it:
 *   [stamp mode only: (nk, n) f64 key-press time stamps]
 *   epoch  u32 (n,)   reset count (RNG stream position), touched by resets only
 *   [TRACK only: f64 (n,) running episode return]
 *
 * Citations: phys = q1physrl_env/q1physrl_env/phys.py, env = q1physrl_env/q1physrl_env/env.py.
 */
#include "q1_internal.h"
#include "q1_device_common.cuh"
#include "q1_sample.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace q1;

/* ====================================================================== device side ========= */

namespace {

constexpr int kBlock = kTile;      /* threads per CTA: one env per thread, one state block per CTA tile */
/* Resident CTAs per SM the persistent step kernel is sized for.  Measured at 2^20 envs (us per tick,
 * same box): 8 CTAs x 64 registers 23.5, 7 x 72: 23.3, 6 x 80: 22.7, 5 x 96: 23.5, 4 x 128: 25.3.  With
 * the libm-exact sin/cos the tick holds more live f64 values; at 64 registers the schedule serialises. */
#ifndef Q1_STEP_CTAS
#define Q1_STEP_CTAS 6
#endif
constexpr int kStepCtasPerSm = Q1_STEP_CTAS;
constexpr int kStepCtasSmall = 7;   /* the variant for launches of few tiles per SM, see k_step_tma */

/* -- TMA (cp.async.bulk) + mbarrier plumbing, shared-memory accesses by 32-bit window address ----- */

__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(bar), "r"(parity)
                 : "memory");
}
/* L2 eviction policy for data that is streamed through once per tick (state, actions, results):
 * evict-first keeps the 126 MB L2 from filling with dirty lines that must be written back later in
 * bursts.  Q1_L2_HINTS: bit 0 = hint the bulk loads, bit 1 = hint the bulk stores. */
#ifndef Q1_L2_HINTS
#define Q1_L2_HINTS 2
#endif
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
/* global -> shared bulk copy that signals `bar` with the byte count when it lands */
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void *gsrc, uint32_t bytes, uint32_t bar,
                                          uint64_t policy)
{
    if (Q1_L2_HINTS & 1)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
                     "[%0], [%1], %2, [%3], %4;" ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar), "l"(policy)
                     : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar)
                     : "memory");
}
/* shared -> global bulk copy, tracked by this thread's bulk async-group */
__device__ __forceinline__ void bulk_store(void *gdst, uint32_t ssrc, uint32_t bytes, uint64_t policy)
{
    if (Q1_L2_HINTS & 2)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                     "r"(ssrc), "r"(bytes), "l"(policy)
                     : "memory");
    else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc),
                     "r"(bytes)
                     : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
/* all bulk groups of this thread have finished READING shared memory */
__device__ __forceinline__ void bulk_wait_read_all()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_smem_to_async_proxy()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds_d2(uint32_t a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_d(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts_f2(uint32_t a, float x, float y)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts_d2(uint32_t a, double x, double y)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void sts_d(uint32_t a, double x)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory");
}
__device__ __forceinline__ void sts_f(uint32_t a, float x)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory");
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t x)
{
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(x) : "memory");
}

/* -- env.VectorPhysEnv.vector_step (env:482-510): one lockstep tick ---------------------------- */

/* One tile = kTile envs.  Shared-memory image of a tile in flight.  Inputs (a ring of kInStages):
 * the state block and the action arrays as they lie in HBM; the state is rewritten in place and
 * stored back from there.  Outputs (a ring of kOutStages): the results as they will lie in HBM. */
enum : uint32_t {
    IN_STATE = 0,                                  /* one state block: rec_a | rec_b | trem */
    IN_MOUSE = IN_STATE + kTileBytes,              /* f32 / i32 mouse actions */
    IN_KEYS = IN_MOUSE + 4 * kTile,                /* kTile x num_keys bytes */
    IN_BYTES = IN_KEYS + 4 * kTile,
    OUT_OBS = 0,                                   /* float[kTile][6] */
    OUT_REWARD = OUT_OBS + 24 * kTile,
    OUT_DONE = OUT_REWARD + 4 * kTile,
    OUT_ZS = OUT_DONE + kTile,
    OUT_BYTES = OUT_ZS + kTile
};
static_assert(IN_BYTES % 128 == 0 && OUT_BYTES % 128 == 0, "every sub-buffer stays 16-byte aligned");
#ifndef Q1_IN_STAGES
#define Q1_IN_STAGES 3
#endif
constexpr int kInStages = Q1_IN_STAGES;   /* input ring depth: loads run kInStages - 1 tiles ahead */
constexpr int kOutStages = 2;

/* Persistent, TMA-pipelined step kernel (counter mode, full tiles, 16-byte aligned buffers, f32 or
 * i32 mouse).  gridDim.x = #SMs x kStepCtasPerSm; CTA c walks tiles c, c + grid, ...
 *
 * All global traffic is bulk copies (cp.async.bulk, the TMA engine): a tile's state block and
 * actions stream into an input stage two tiles ahead (3 copies, completion on an mbarrier); the
 * threads read their env from shared memory, run the fused tick, write the new state back in place
 * and the results into an output stage; after one CTA barrier the tile leaves as 5 bulk stores.
 * The issue work is spread over the four warps' lane 0 (state store / obs+reward / done+zs / next
 * loads), and every wait on an earlier bulk group sits one full tile after its issue, so no warp
 * ever blocks on the TMA engine.  Threads touch only shared memory (16-byte LDS/STS, 32-bit
 * addresses): no per-thread global address arithmetic, no load latency on the compute warps. */
/* 1: k_step_tma moves its bytes but skips the tick (measures the ceiling of the memory pipeline
 * alone; the results are meaningless).  Never set in the shipped build. */
#ifndef Q1_PASSTHROUGH
#define Q1_PASSTHROUGH 0
#endif

/* CTAS: resident CTAs per SM the build is register-limited for.  6 x 80 registers is the fastest at
 * 2^20 envs; 7 x 72 wins when a launch has few tiles per SM: 131 072 envs are 1024 tiles, which 148 x 7
 * = 1036 resident CTAs take in one wave where 148 x 6 = 888 need a second one for 136 of them
 * (5.6 vs 6.0 us per tick; 262 144 envs: 8.5 vs 8.9). */
template <bool TRACK, bool LEAN, bool COMMON, int CTAS>
__global__ void __launch_bounds__(kBlock, CTAS)
k_step_tma(const __grid_constant__ Params P, const uint8_t *__restrict__ keys,
           const void *__restrict__ mouse, int mouse_kind, float *__restrict__ obs,
           float *__restrict__ reward, uint8_t *__restrict__ done,
           uint8_t *__restrict__ zero_start, int auto_reset, int64_t tile_begin, int64_t tiles)
{
    /* tiles [tile_begin, tiles) of the handle; the global buffers are indexed by absolute env */
    __shared__ __align__(128) unsigned char in_mem[kInStages * IN_BYTES];
    __shared__ __align__(128) unsigned char out_mem[kOutStages * OUT_BYTES];
    __shared__ __align__(8) uint64_t full_bar[kInStages];
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const bool issuer = (tid & 31u) == 0;
    const uint32_t nk = (uint32_t)P.num_keys;
    const bool has_mouse = COMMON ? true : (bool)P.allow_yaw;
    const uint32_t mouse_bytes = has_mouse ? 4u * kTile : 0u;
    const uint32_t in_bytes = kTileBytes + nk * kTile + mouse_bytes;
    uint32_t in0 = smem_addr(in_mem), out0 = smem_addr(out_mem), bar0 = smem_addr(full_bar);
    asm volatile("" : "+r"(in0), "+r"(out0), "+r"(bar0)); /* pinned: no re-derivation per tile */

    const uint64_t l2pol = Q1_L2_HINTS ? l2_evict_first_policy() : 0;
    auto issue_loads = [&](uint32_t s, int64_t tile) {
        asm volatile("" : "+l"(tile)); /* address arithmetic stays inside the issuing lane's branch */
        const uint32_t st = in0 + s * IN_BYTES, bar = bar0 + s * 8u;
        mbar_expect_tx(bar, in_bytes);
        bulk_load(st + IN_STATE, P.state + tile * kTileBytes, kTileBytes, bar, l2pol);
        bulk_load(st + IN_KEYS, keys + tile * (nk * kTile), nk * kTile, bar, l2pol);
        if (mouse_bytes)
            bulk_load(st + IN_MOUSE, static_cast<const char *>(mouse) + tile * mouse_bytes, mouse_bytes, bar, l2pol);
    };

    /* Programmatic dependent launch: the next kernel in the stream may start placing its CTAs as
     * ours retire and run its own prologue; every kernel of this library waits for the full
     * completion (and memory visibility) of its predecessor before it touches global memory. */
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kInStages; s++)
            mbar_init(bar0 + s * 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_smem_to_async_proxy();
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (warp == 3 && issuer) { /* the loading lane primes all stages but one; the last fills after tile 0 */
#pragma unroll
        for (int s = 0; s < kInStages - 1; s++) {
            const int64_t tile = tile_begin + blockIdx.x + (int64_t)s * gridDim.x;
            if (tile < tiles)
                issue_loads(s, tile);
        }
    }

    uint32_t s = 0, parity = 0, so = 0;
    for (int64_t tile = tile_begin + blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t sb = in0 + s * IN_BYTES, ob = out0 + so * OUT_BYTES;
        mbar_wait(bar0 + s * 8u, parity);

        Env e;
        {
            const float4 a = lds_f4(sb + IN_STATE + kTileRecA + tid * 16u);
            const double2 b = lds_d2(sb + IN_STATE + kTileRecB + tid * 16u);
            e.vx = a.x;
            e.vy = a.y;
            e.vz = a.z;
            e.bits = __float_as_uint(a.w);
            e.z = b.x;
            e.yaw = b.y;
            e.trem = lds_d(sb + IN_STATE + kTileTrem + tid * 8u);
        }
        uint32_t keybits;
        if (nk == 4) {
            const uint32_t w = lds_u32(sb + IN_KEYS + tid * 4u);
            keybits = (w & 1u) | ((w >> 7) & 2u) | ((w >> 14) & 4u) | ((w >> 21) & 8u);
        } else {
            const uint32_t ka = sb + IN_KEYS + tid * 3u;
            keybits = (lds_u8(ka) & 1u) | ((lds_u8(ka + 1) & 1u) << 1) | ((lds_u8(ka + 2) & 1u) << 2);
        }
        double m = 0.0;
        if (has_mouse) {
            if (COMMON || mouse_kind == Q1_MOUSE_F32)
                m = (double)__uint_as_float(lds_u32(sb + IN_MOUSE + tid * 4u));
            else
                m = (double)(int32_t)lds_u32(sb + IN_MOUSE + tid * 4u);
        }
        float r;
        bool d;
        if (Q1_PASSTHROUGH) { /* traffic-only build: same bytes in and out, no arithmetic */
            r = e.vx + (float)m;
            d = keybits == 0xffu;
        } else {
            tick<false, LEAN, COMMON>(P, e, keybits, m, r, d);
        }
        const bool zs = e.bits & F_ZERO_START;
        bool finished = false;
        double ret = 0.0;
        const int64_t i = tile * kTile + tid;
        if (TRACK) {
            ret = add64(P.ep_return[i], (double)r);
            finished = d && !(e.bits & F_DONE_SEEN);
            if (finished)
                e.bits |= F_DONE_SEEN;
        }
        sts_f(ob + OUT_REWARD + tid * 4u, r);
        sts_u8(ob + OUT_DONE + tid, d ? 1u : 0u);
        sts_u8(ob + OUT_ZS + tid, zs ? 1u : 0u);
        if (d && auto_reset) {
            uint32_t ep = P.epoch[i] + 1u;
            P.epoch[i] = ep;
            reset_env<false>(P, e, P.env_index_base + (uint64_t)i, ep);
            if (TRACK)
                P.ep_return[i] = 0.0;
        } else if (TRACK) {
            P.ep_return[i] = ret;
        }
        float o[6];
        if (Q1_PASSTHROUGH) {
            o[0] = e.vx; o[1] = e.vy; o[2] = e.vz; o[3] = (float)e.z; o[4] = (float)e.yaw; o[5] = (float)e.trem;
        } else {
            observe<LEAN>(P, e, o);
        }
        sts_f2(ob + OUT_OBS + tid * 24u, o[0], o[1]);
        sts_f2(ob + OUT_OBS + tid * 24u + 8u, o[2], o[3]);
        sts_f2(ob + OUT_OBS + tid * 24u + 16u, o[4], o[5]);
        sts_f4(sb + IN_STATE + kTileRecA + tid * 16u, e.vx, e.vy, e.vz, __uint_as_float(e.bits));
        sts_d2(sb + IN_STATE + kTileRecB + tid * 16u, e.z, e.yaw);
        sts_d(sb + IN_STATE + kTileTrem + tid * 8u, e.trem);
        fence_smem_to_async_proxy();
        /* the stores this lane issued one tile ago have long finished reading shared memory; the
         * wait makes that a guarantee before the barrier lets anyone reuse those buffers */
        if (issuer)
            bulk_wait_read_all();
        __syncthreads();
        if (issuer) {
            int64_t t = tile;
            asm volatile("" : "+l"(t));
            if (warp == 0) {
                bulk_store(P.state + t * kTileBytes, sb + IN_STATE, kTileBytes, l2pol);
                bulk_commit();
            } else if (warp == 1) {
                bulk_store(obs + t * (6 * kTile), ob + OUT_OBS, 24 * kTile, l2pol);
                bulk_store(reward + t * kTile, ob + OUT_REWARD, 4 * kTile, l2pol);
                bulk_commit();
            } else if (warp == 2) {
                bulk_store(done + t * kTile, ob + OUT_DONE, kTile, l2pol);
                if (zero_start)
                    bulk_store(zero_start + t * kTile, ob + OUT_ZS, kTile, l2pol);
                bulk_commit();
            } else {
                /* refill the stage of the previous tile: all warps left it a barrier ago and its
                 * state store was drained before this barrier */
                const int64_t next = t + (kInStages - 1) * (int64_t)gridDim.x;
                if (next < tiles)
                    issue_loads(s >= 1 ? s - 1 : kInStages - 1, next);
            }
        }
        if (TRACK)
            report_episodes(P, finished, zs, ret);
        if (++s == kInStages) {
            s = 0;
            parity ^= 1u;
        }
        so ^= 1u;
    }
    if (issuer)
        bulk_wait_read_all(); /* shared memory must outlive the reads of the last bulk stores; their
                                 global writes complete with the grid */
}

/* The same tick, one env per thread with plain loads and stores: f64-stamp mode, ragged tails
 * (n not a multiple of kTile) and buffers that are not 16-byte aligned.  Covers envs [first, end). */
template <bool STAMPS, bool TRACK, bool LEAN>
__global__ void __launch_bounds__(kBlock)
k_step(const __grid_constant__ Params P, const uint8_t *__restrict__ keys,
       const void *__restrict__ mouse, int mouse_kind, float *__restrict__ obs,
       float *__restrict__ reward, uint8_t *__restrict__ done, uint8_t *__restrict__ zero_start,
       int auto_reset, int64_t first, int64_t end)
{
    /* same protocol as k_step_tma: let the next launch stage itself, touch global memory only after
     * every earlier launch has completed (no-ops when launched without the attribute) */
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int64_t i = first + (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool active = i < end;
    bool finished = false, zs = false;
    double ret = 0.0;
    if (active) {
        Env e;
        load_env<STAMPS>(P, i, e);
        uint32_t keybits = load_keys(keys, i, P.num_keys);
        double m = 0.0;
        if (P.allow_yaw)
            m = load_mouse(mouse, mouse_kind, i);
        float r;
        bool d;
        tick<STAMPS, LEAN, false>(P, e, keybits, m, r, d);
        zs = e.bits & F_ZERO_START;
        if (TRACK) {
            ret = add64(P.ep_return[i], (double)r);
            finished = d && !(e.bits & F_DONE_SEEN);
            if (finished)
                e.bits |= F_DONE_SEEN;
        }
        reward[i] = r;
        done[i] = d ? 1 : 0;
        if (zero_start)
            zero_start[i] = zs ? 1 : 0;
        if (d && auto_reset) {
            uint32_t ep = P.epoch[i] + 1u;
            P.epoch[i] = ep;
            reset_env<STAMPS>(P, e, P.env_index_base + (uint64_t)i, ep);
            if (TRACK)
                P.ep_return[i] = 0.0;
        } else if (TRACK) {
            P.ep_return[i] = ret;
        }
        float o[6];
        observe<LEAN>(P, e, o);
        store_obs(obs, i, o);
        store_env<STAMPS>(P, i, e);
    }
    if (TRACK)
        report_episodes(P, finished, zs, ret);
}

/* -- vector_reset / reset_at (env:428-480) -------------------------------------------------- */

template <bool STAMPS, bool TRACK, bool LEAN>
__global__ void __launch_bounds__(kBlock)
k_reset(const __grid_constant__ Params P, const uint8_t *__restrict__ mask, int64_t only,
        float *__restrict__ obs, int64_t obs_row_offset)
{
    int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (only >= 0)
        i = (i == 0) ? only : P.n;
    if (i >= P.n || (mask && !mask[i]))
        return;
    Env e;
    uint32_t ep = P.epoch[i] + 1u;
    P.epoch[i] = ep;
    reset_env<STAMPS>(P, e, P.env_index_base + (uint64_t)i, ep);
    store_env<STAMPS>(P, i, e);
    if (TRACK)
        P.ep_return[i] = 0.0;
    if (obs) {
        float o[6];
        observe<LEAN>(P, e, o);
        store_obs(obs, i + obs_row_offset, o);
    }
}

template <bool STAMPS, bool LEAN>
__global__ void __launch_bounds__(kBlock)
k_observe(const __grid_constant__ Params P, float *__restrict__ obs)
{
    int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= P.n)
        return;
    Env e;
    load_env<STAMPS>(P, i, e);
    float o[6];
    observe<LEAN>(P, e, o);
    store_obs(obs, i, o);
}

/* -- multi-tick rollout: state stays in registers for all ticks -------------------------------- */

/* Where a rollout's actions come from: a built-in device-side policy, or (RECORD launches) caller
 * arrays laid out [tick][env]. */
struct ActionFeed {
    int policy;                 /* Q1_POLICY_*; -1: the arrays below */
    uint64_t policy_seed;
    const uint8_t *keys;        /* (ticks, n, num_keys) */
    const void *mouse;          /* (ticks, n) of mouse_kind */
    int mouse_kind;
};

/* RECORD: the per-tick trajectory recorder behind analyse.eval_sim (q1physrl/analyse.py:197-240),
 * N envs at once: before each tick the env's movement state and observation (analyse.py:218-219),
 * the action, the move command ActionDecoder.map makes of it (analyse.py:215-216, taken from the
 * SAME tick<>() that advances the env rather than from a shadow decoder) and, after the tick, reward
 * and done are written as row t of the [tick][env] arrays of q1_record_view.  auto_reset = 0 keeps
 * the reference's behaviour (an env whose episode ended keeps stepping, env:505-506). */
template <bool STAMPS, bool TRACK, bool LEAN, bool RECORD>
__global__ void __launch_bounds__(kBlock)
k_rollout(const __grid_constant__ Params P, const ActionFeed feed, int ticks, uint32_t tick_base,
          float *__restrict__ obs, float *__restrict__ reward_sum, int auto_reset, uint32_t record_flags,
          const q1_record_view rec)
{
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool active = i < P.n;
    const int64_t ii = active ? i : 0;
    const uint64_t gidx = P.env_index_base + (uint64_t)ii;
    const int nk = P.num_keys;
    Env e;
    load_env<STAMPS>(P, ii, e);
    uint32_t ep = P.epoch[ii];
    double ret = TRACK ? P.ep_return[ii] : 0.0;
    float rsum = 0.0f;
    for (int t = 0; t < ticks; t++) {
        uint32_t keybits;
        double m;
        if (!RECORD || feed.policy >= 0) {
            policy_action(P, feed.policy, feed.policy_seed, gidx, tick_base + (uint32_t)t, keybits, m);
        } else {
            const int64_t row = (int64_t)t * P.n + ii;
            keybits = load_keys(feed.keys, row, nk);
            m = P.allow_yaw ? load_mouse(feed.mouse, feed.mouse_kind, row) : 0.0;
        }
        float r;
        bool d;
        if (RECORD) {
            const int64_t row = (int64_t)t * P.n + ii;
            float o[6];
            observe<LEAN>(P, e, o);           /* what the policy saw (the hover override, env:483-485,
                                                 happens inside the tick, after it) */
            if (active)
                record_before(rec, row, nk, e, o, keybits, m);
            Move mv;
            tick<STAMPS, LEAN, false>(P, e, keybits, m, r, d, &mv);
            if (active)
                record_after(rec, row, record_flags, P.jump_mode, mv, o, r, d);
        } else {
            tick<STAMPS, LEAN, false>(P, e, keybits, m, r, d);
        }
        rsum = add32(rsum, r);
        if (TRACK) {
            ret = add64(ret, (double)r);
            bool finished = active && d;
            if (RECORD && !auto_reset) {            /* report an episode once, as the step kernels do */
                finished = finished && !(e.bits & F_DONE_SEEN);
                if (finished)
                    e.bits |= F_DONE_SEEN;
            }
            report_episodes(P, finished, e.bits & F_ZERO_START, ret);
        }
        if (d && (!RECORD || auto_reset)) {
            ep += 1u;
            reset_env<STAMPS>(P, e, gidx, ep);
            ret = 0.0;
        }
    }
    if (!active)
        return;
    store_env<STAMPS>(P, i, e);
    P.epoch[i] = ep;
    if (TRACK)
        P.ep_return[i] = ret;
    if (reward_sum)
        reward_sum[i] = rsum;
    if (obs) {
        float o[6];
        observe<LEAN>(P, e, o);
        store_obs(obs, i, o);
    }
}

/* -- phys.apply on explicit arrays (phys:184-197) ----------------------------------------------- */

/* One row of phys.apply with general pitch / roll; dt_f32: the time_delta array was float32. */
__device__ __forceinline__ void phys_apply_row(double yaw, double pitch, double roll, bool has_pitch,
                                               bool has_roll, double fmove, double smove, bool jump,
                                               double dt, bool dt_f32, float &vx, float &vy,
                                               float &vz, double &z, bool &og, bool &jr)
{
    /* phys:58-66 */
    double sy, cy, sp = 0.0, cp = 1.0, sr = 0.0, cr = 1.0;
    sincos_ref(div64(mul64(yaw, kPi), 180.0), sy, cy);
    if (has_pitch)
        sincos_ref(div64(mul64(pitch, kPi), 180.0), sp, cp);
    if (has_roll)
        sincos_ref(div64(mul64(roll, kPi), 180.0), sr, cr);
    double fx = mul64(cp, cy);
    double rx = add64(mul64(mul64(mul64(-1.0, sr), sp), cy), mul64(mul64(-1.0, cr), -sy));
    double fy = mul64(cp, sy);
    double ry = add64(mul64(mul64(mul64(-1.0, sr), sp), sy), mul64(mul64(-1.0, cr), cy));
    if (dt_f32) {
        /* phys:78 f32(10) * f32(dt), phys:122 f32(800) * f32(dt): f32 products */
        const float dtf = (float)dt;
        move_body<false, true>(vx, vy, vz, z, og, jr, fx, rx, fy, ry, fmove, smove, jump, dt,
                               (double)mul32(10.0f, dtf), (double)mul32(800.0f, dtf));
    } else {
        move_body<false, false>(vx, vy, vz, z, og, jr, fx, rx, fy, ry, fmove, smove, jump, dt,
                                mul64(10.0, dt), mul64(800.0, dt));
    }
}

__global__ void __launch_bounds__(kBlock)
k_phys_apply(int64_t n, const double *__restrict__ yaw, const double *__restrict__ pitch,
             const double *__restrict__ roll, const double *__restrict__ fmove,
             const double *__restrict__ smove, const uint8_t *__restrict__ button2,
             const double *__restrict__ time_delta, int dt_f32, const double *__restrict__ z_pos,
             const float *__restrict__ vel, const uint8_t *__restrict__ on_ground,
             const uint8_t *__restrict__ jump_released, double *__restrict__ z_out,
             float *__restrict__ vel_out, uint8_t *__restrict__ og_out, uint8_t *__restrict__ jr_out)
{
    int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n)
        return;
    float vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
    double z = z_pos[i];
    bool og = on_ground[i] != 0, jr = jump_released[i] != 0;
    phys_apply_row(yaw[i], pitch ? pitch[i] : 0.0, roll ? roll[i] : 0.0, pitch != nullptr,
                   roll != nullptr, fmove[i], smove[i], button2[i] != 0, time_delta[i], dt_f32 != 0,
                   vx, vy, vz, z, og, jr);
    z_out[i] = z;
    vel_out[3 * i] = vx;
    vel_out[3 * i + 1] = vy;
    vel_out[3 * i + 2] = vz;
    og_out[i] = og;
    jr_out[i] = jr;
}

/* phys.apply when PlayerState.vel is FLOAT64 (PlayerState.from_df, phys:163-170: the notebook's comparison
 * with recorded game frames).  NumPy then keeps the friction speed (phys:85), the stored velocity (phys:190
 * assigns into an f64 array: no rounding) and the z velocity (phys:119-122) in f64.  dt_f32: 10 * dt and
 * 800 * dt are float32 products (phys:78, 122 with a float32 time_delta array). */
__global__ void __launch_bounds__(kBlock)
k_phys_apply_vel64(int64_t n, const double *__restrict__ yaw, const double *__restrict__ pitch,
                   const double *__restrict__ roll, const double *__restrict__ fmove,
                   const double *__restrict__ smove, const uint8_t *__restrict__ button2,
                   const double *__restrict__ time_delta, int dt_f32, const double *__restrict__ z_pos,
                   const double *__restrict__ vel, const uint8_t *__restrict__ on_ground,
                   const uint8_t *__restrict__ jump_released, double *__restrict__ z_out,
                   double *__restrict__ vel_out, uint8_t *__restrict__ og_out, uint8_t *__restrict__ jr_out)
{
    int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n)
        return;
    double sy, cy, sp = 0.0, cp = 1.0, sr = 0.0, cr = 1.0;                                  /* phys:58-66 */
    sincos_ref(div64(mul64(yaw[i], kPi), 180.0), sy, cy);
    if (pitch)
        sincos_ref(div64(mul64(pitch[i], kPi), 180.0), sp, cp);
    if (roll)
        sincos_ref(div64(mul64(roll[i], kPi), 180.0), sr, cr);
    const double fx = mul64(cp, cy);
    const double rx = add64(mul64(mul64(mul64(-1.0, sr), sp), cy), mul64(mul64(-1.0, cr), -sy));
    const double fy = mul64(cp, sy);
    const double ry = add64(mul64(mul64(mul64(-1.0, sr), sp), sy), mul64(mul64(-1.0, cr), cy));
    const double dt = time_delta[i];
    const float dtf = (float)dt;
    const bool was_on_ground = on_ground[i] != 0, jump = button2[i] != 0;
    double vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];

    double wx = add64(mul64(fx, fmove[i]), mul64(rx, smove[i]));                            /* phys:95-103 */
    double wy = add64(mul64(fy, fmove[i]), mul64(ry, smove[i]));
    double ws = __dsqrt_rn(add64(mul64(wx, wx), mul64(wy, wy)));
    double wdx = wx, wdy = wy;
    if (ws > 0.0) {
        wdx = div64(wx, ws);
        wdy = div64(wy, ws);
    }
    double wish_speed = ws < (double)kMaxSpeed ? ws : (double)kMaxSpeed;
    if (ws != ws)
        wish_speed = ws;
    if (was_on_ground) {                                                                    /* phys:83-90 */
        const double speed = __dsqrt_rn(add64(mul64(vx, vx), mul64(vy, vy)));
        const double control = speed > (double)kStopSpeed ? speed : (double)kStopSpeed;
        double new_speed = sub64(speed, mul64(mul64(dt, control), (double)kFriction));
        if (!(new_speed > 0.0))
            new_speed = 0.0;
        if (speed > 0.0) {
            const double ratio = div64(new_speed, speed);
            vx = mul64(vx, ratio);
            vy = mul64(vy, ratio);
        }
    }
    const double current = add64(mul64(vx, wdx), mul64(vy, wdy));                           /* phys:69-80 */
    const double clipped = (wish_speed > 30.0 && !was_on_ground) ? 30.0 : wish_speed;
    double add = sub64(clipped, current);
    if (!(add > 0.0))
        add = 0.0;
    double accel = mul64(dt_f32 ? (double)mul32(10.0f, dtf) : mul64(10.0, dt), wish_speed);
    if (add < accel)
        accel = add;
    vx = add64(vx, mul64(accel, wdx));
    vy = add64(vy, mul64(accel, wdy));

    const bool jr = (jump_released[i] != 0) | !jump;                                        /* phys:112-132 */
    const bool do_jump = was_on_ground && jump && jr;
    vz = add64(vz, do_jump ? (double)kJumpSpeed : 0.0);
    vz = sub64(vz, dt_f32 ? (double)mul32(800.0f, dtf) : mul64(800.0, dt));
    const double z = add64(z_pos[i], mul64(dt, vz));
    const bool og = z < (double)kFloorHeight;
    z_out[i] = og ? (double)kFloorHeight : z;
    vel_out[3 * i] = vx;
    vel_out[3 * i + 1] = vy;
    vel_out[3 * i + 2] = og ? 0.0 : vz;
    og_out[i] = og;
    jr_out[i] = jr;
}

/* q1physrl/analyse.py:92-118 `hypothetical_delta_speeds` in one launch: for every frame t and every
 * relative wish angle a, the ground-speed change of one phys.apply tick with yaw = base_yaw[t] +
 * rel_angle[a] and constant fmove / smove / time_delta.  out[a * n + t] = |v'| - |v| in f32. */
__global__ void __launch_bounds__(kBlock)
k_delta_speed_sweep(int64_t n, int64_t num_angles, const double *__restrict__ base_yaw,
                    const double *__restrict__ rel_angle, double fmove, double smove,
                    const uint8_t *__restrict__ button2, double time_delta, int dt_f32,
                    const double *__restrict__ z_pos, const float *__restrict__ vel,
                    const uint8_t *__restrict__ on_ground, const uint8_t *__restrict__ jump_released,
                    float *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t a = blockIdx.y;
    if (t >= n || a >= num_angles)
        return;
    float vx = vel[3 * t], vy = vel[3 * t + 1], vz = vel[3 * t + 2];
    double z = z_pos[t];
    bool og = on_ground[t] != 0, jr = jump_released[t] != 0;
    const float before = __fsqrt_rn(add32(mul32(vx, vx), mul32(vy, vy)));
    phys_apply_row(add64(base_yaw[t], rel_angle[a]), 0.0, 0.0, true, true, fmove, smove,
                   button2[t] != 0, time_delta, dt_f32 != 0, vx, vy, vz, z, og, jr);
    const float after = __fsqrt_rn(add32(mul32(vx, vx), mul32(vy, vy)));
    out[a * n + t] = __fsub_rn(after, before);
}

/* -- env.ActionDecoder.map on explicit decoder state (env:225-269), f64 stamps ------------------- */

__global__ void __launch_bounds__(kBlock)
k_decode(const __grid_constant__ Params P, uint8_t *__restrict__ last_keys,
         double *__restrict__ last_press, double *__restrict__ yaw,
         const uint8_t *__restrict__ keys, const double *__restrict__ mouse,
         const float *__restrict__ z_vel, const double *__restrict__ time_remaining,
         int64_t *__restrict__ smove, int64_t *__restrict__ fmove, uint8_t *__restrict__ jump)
{
    int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= P.n)
        return;
    const int nk = P.num_keys;
    double mouse_x = 0.0;
    if (P.allow_yaw) {
        double m = mouse[i];
        if (!P.discrete_yaw)
            mouse_x = div64(mul64(m, P.max_yaw_delta), P.action_range);
        else
            mouse_x = div64(mul64(sub64(m, P.yaw_steps), P.max_yaw_delta), P.yaw_steps);
    }
    const double now = sub64(P.time_limit, time_remaining[i]);
    int downs[4] = {0, 0, 0, 0}, lasts[4] = {0, 0, 0, 0};
    for (int k = 0; k < nk; k++) {
        int lk = last_keys[i * nk + k] & 1;
        bool elapsed = now >= add64(last_press[i * nk + k], P.key_delay);
        int d = (keys[i * nk + k] & 1) & ((elapsed ? 1 : 0) | lk);
        if (d & ~lk & 1)
            last_press[i * nk + k] = now;
        last_keys[i * nk + k] = (uint8_t)d;
        downs[k] = d;
        lasts[k] = lk;
    }
    int f2, s2;
    if (P.smooth_keys) {
        f2 = downs[KEY_FORWARD] + lasts[KEY_FORWARD];
        s2 = (downs[KEY_RIGHT] + lasts[KEY_RIGHT]) - (downs[KEY_LEFT] + lasts[KEY_LEFT]);
    } else {
        f2 = 2 * downs[KEY_FORWARD];
        s2 = 2 * (downs[KEY_RIGHT] - downs[KEY_LEFT]);
    }
    double fm = f2 == 2 ? P.fmove_full : (f2 == 1 ? P.fmove_half : 0.0);
    int as2 = s2 < 0 ? -s2 : s2;
    double sm = as2 == 2 ? P.smove_full : (as2 == 1 ? P.smove_half : 0.0);
    if (s2 < 0)
        sm = -sm;
    yaw[i] = add64(yaw[i], mouse_x);
    smove[i] = (int64_t)sm;
    fmove[i] = (int64_t)fm;
    jump[i] = P.auto_jump ? (z_vel[i] <= 16.0f) : (P.allow_jump ? (uint8_t)downs[KEY_JUMP] : 0);
}

/* -- self-test of the branch-free division sequences against the IEEE intrinsics ------------------ */

__device__ __forceinline__ uint64_t splitmix64(uint64_t &x)
{
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* random double with a uniform significand and an exponent in [-span, span] */
__device__ __forceinline__ double random_double(uint64_t &rng, int span, bool signed_)
{
    uint64_t w = splitmix64(rng);
    uint64_t mant = w & 0xFFFFFFFFFFFFFull;
    uint32_t sel = (uint32_t)(w >> 52) & 0x3FFu;
    if ((sel & 0x3Fu) == 0)
        mant = 0xFFFFFFFFFFFFFull;               /* significand of all ones: Markstein's exception */
    else if ((sel & 0x3Fu) == 1)
        mant = 0;
    else if ((sel & 0x3Fu) == 2)
        mant = 1;
    int e = (int)(splitmix64(rng) % (uint64_t)(2 * span + 1)) - span;
    uint64_t bits = ((uint64_t)(1023 + e) << 52) | mant;
    if (signed_ && (w >> 63))
        bits |= 0x8000000000000000ull;
    return __longlong_as_double((long long)bits);
}

/* sincos_ref on an array: the device build of q1_libm_sincos.cuh, exposed so that tests can compare
 * it bit for bit with the host C library the reference calls through NumPy */
__global__ void __launch_bounds__(256)
k_sincos(int64_t n, const double *__restrict__ x, double *__restrict__ s, double *__restrict__ c)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    double sv, cv;
    sincos_ref(x[i], sv, cv);
    s[i] = sv;
    c[i] = cv;
}

__global__ void __launch_bounds__(256)
k_selftest(uint64_t iters, uint64_t seed, unsigned long long *__restrict__ out)
{
    uint64_t rng = seed * 0x2545F4914F6CDD1Dull + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2 + 1;
    unsigned long long bad[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double consts[8] = {10.0, (double)10.08f, 180.0, 90.0, 5.0, 7.0, 4.0, 0.013888888888888};
    for (uint64_t it = 0; it < iters; it++) {
        /* 0: reciprocal, wide range */
        double b = random_double(rng, 200, false);
        double y = rcp_rn(b);
        bad[0] += __double_as_longlong(y) != __double_as_longlong(__drcp_rn(b));
        /* 1: quotient by a variable divisor, wide range */
        double a = random_double(rng, 200, true);
        bad[1] += __double_as_longlong(div_rcp(a, b, y)) != __double_as_longlong(__ddiv_rn(a, b));
        /* 2: quotient by constants the path uses (and a random constant) */
        double c = (it & 8) ? random_double(rng, 20, false) : consts[it & 7];
        double yc = __drcp_rn(c);
        bad[2] += __double_as_longlong(div_const(a, c, yc)) != __double_as_longlong(__ddiv_rn(a, c));
        if (short_division_ok_dev(c, yc))   /* the three-operation form, where its criterion holds */
            bad[2] += __double_as_longlong(div_const3(a, c, yc)) != __double_as_longlong(__ddiv_rn(a, c));
        /* 3: the physics ranges: wish velocity / wish speed, new_speed / speed */
        double ws = 1.0 + (double)(splitmix64(rng) >> 11) * (2000.0 / 9007199254740992.0);
        double wx = ((double)(splitmix64(rng) >> 11) * (2.0 / 9007199254740992.0) - 1.0) * ws;
        bad[3] += __double_as_longlong(div_rcp(wx, ws, rcp_rn(ws))) !=
                  __double_as_longlong(__ddiv_rn(wx, ws));
        float sp = __uint_as_float(0x30000000u + (uint32_t)(splitmix64(rng) % 0x16000000ull));
        double ns = (double)sp * ((double)(splitmix64(rng) >> 11) * (1.0 / 9007199254740992.0));
        bad[4] += __double_as_longlong(div_rcp(ns, (double)sp, rcp_rn((double)sp))) !=
                  __double_as_longlong(__ddiv_rn(ns, (double)sp));
    }
    /* 5: the f32 observation quotients, exhaustively: multiples of 16 over 200 and multiples of
     * 1/8 over 100, against the reference's f64 division rounded to f32 */
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t m = tid; m < (1ull << 21); m += nth) {
        float q = (float)((long long)m - (1ll << 20));
        float v = q * 16.0f;
        bad[5] += __float_as_uint(div_const32(v, 200.0f, 1.0f / 200.0f)) !=
                  __float_as_uint(__double2float_rn(__ddiv_rn((double)v, 200.0)));
        bad[6] += __float_as_uint(div_const32_long(v, 200.0f, 1.0f / 200.0f)) !=
                  __float_as_uint(__double2float_rn(__ddiv_rn((double)v, 200.0)));
    }
    for (uint64_t m = tid; m < (1ull << 25); m += nth) {
        float zq = (float)((long long)m - (1ll << 24)) * 0.125f;
        bad[5] += __float_as_uint(div_const32(zq, 100.0f, 1.0f / 100.0f)) !=
                  __float_as_uint(__double2float_rn(__ddiv_rn((double)zq, 100.0)));
        bad[7] += __float_as_uint(div_const32_long(zq, 100.0f, 1.0f / 100.0f)) !=
                  __float_as_uint(__double2float_rn(__ddiv_rn((double)zq, 100.0)));
    }
    for (int k = 0; k < 8; k++)
        if (bad[k])
            atomicAdd(&out[k], bad[k]);
}

/* -- q1physrl/action_dist.py: sampling from the policy's output distribution ---------------------- */

/* Q1PhysActionDist (action_dist.py:199-243) on a batch of policy outputs: one Categorical(2) per key
 * action, then GaussianSquashedGaussian for the mouse action (mean, log_std clipped as in
 * action_dist.py:67-76; squash = clip(NormalCDF(raw / 0.90685), 1e-6, 1 - 1e-6) * (high - low) + low,
 * action_dist.py:151, 186-192).  deterministic: argmax / squash(mean) (action_dist.py:84-88).
 * Noise: Philox4x32-10 keyed by `seed`, counter (env index, step). */
__global__ void __launch_bounds__(256)
k_sample_actions(int64_t n, int num_keys, const float *__restrict__ logits, float low, float high,
                 int deterministic, uint64_t seed, uint64_t step,
                 const uint64_t *__restrict__ step_device, uint64_t env_index_base,
                 uint8_t *__restrict__ keys, float *__restrict__ mouse)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n)
        return;
    if (step_device)
        step = *step_device; /* a replayed CUDA graph advances the noise stream through memory */
    float row[10];
    const int width = 2 * num_keys + 2;
    for (int k = 0; k < width; k++)
        row[k] = logits[i * width + k];
    float m;
    const uint32_t kb = sample_action_row(row, num_keys, low, high, deterministic != 0, seed, step,
                                          env_index_base + (uint64_t)i, &m);
    for (int k = 0; k < num_keys; k++)
        keys[i * num_keys + k] = (kb >> k) & 1u;
    mouse[i] = m;
}

} // namespace

/* ====================================================================== host side =========== */

static thread_local std::string g_error;

static int fail(int code, const std::string &msg)
{
    g_error = msg;
    return code;
}

/* the other translation units of the library report errors through the same thread-local slot */
int q1_set_error(int code, const std::string &msg) { return fail(code, msg); }

#define Q1_CUDA(call)                                                                             \
    do {                                                                                          \
        cudaError_t err__ = (call);                                                               \
        if (err__ != cudaSuccess)                                                                 \
            return fail(Q1_ECUDA, std::string(#call) + ": " + cudaGetErrorString(err__));         \
    } while (0)

struct q1_env {
    Params P{};
    q1_config cfg{};
    int device = 0;
    bool stamps = false;
    bool track = false;
    uint32_t flags = 0;
    void *pool = nullptr;
    size_t pool_bytes = 0;
    int state_bytes_per_env = 0;
    uint64_t ticks = 0;
    /* device scratch + stream of the *_host entry points */
    int sm_count = 148;
    bool pdl = true; /* launch the step kernel with programmatic stream serialization */
    int host_chunks = 2; /* pipeline depth of q1_step_host for large page-locked batches */
    bool balance_grid = false; /* step kernel: shrink the grid so that all CTAs walk equally many tiles */
    int64_t small_launch_tiles = 0; /* launches of at most this many tiles take the 7-CTAs-per-SM build */
    bool host_direct = true; /* q1_step_host: let the step kernel read / write page-locked host buffers
                                itself (mapped memory over PCIe) instead of staging them through HBM */
    cudaStream_t host_stream = nullptr;
    cudaStream_t in_stream = nullptr, out_stream = nullptr; /* the chunked pipeline of q1_step_host */
    cudaEvent_t ev_in[8] = {}, ev_done[8] = {};
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    void *bounce = nullptr;      /* page-locked, device-mapped staging of q1_step_host for small batches */
    size_t bounce_bytes = 0;
    char *bounce_dev = nullptr;  /* the address the device reaches `bounce` under */
    /* Set by every entry point that launches on a CALLER's stream (q1_step, q1_reset_all / _masked,
     * q1_rollout, q1_rollout_record, q1_observe).  The *_host entry points run on the handle's private
     * stream and return synchronised; when the handle has also been driven on caller streams they
     * first wait for the device, so that a host call never reads or writes state a caller-stream
     * launch is still working on.  Handles used through the *_host calls only never pay for it. */
    bool caller_streams_used = false;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess)
            prev = -1;
        if (prev == dev)
            prev = -1;          /* already current: nothing to switch, nothing to restore */
        else
            ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

/* see q1_env::caller_streams_used */
int order_after_caller_streams(q1_env *env)
{
    if (env->caller_streams_used) {
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) {
            g_error = std::string("cudaDeviceSynchronize: ") + cudaGetErrorString(err);
            return Q1_ECUDA;
        }
    }
    return Q1_OK;
}

/* env.Config -> the numbers the kernels use, each in the width the reference computes it in. */
int derive_params(const q1_config &c, Params &P, bool &counters_exact, bool numpy1_promotion = false)
{
    if (c.num_envs <= 0)
        return fail(Q1_EINVAL, "num_envs must be positive");
    if (!(c.time_delta > 0))
        return fail(Q1_EINVAL, "time_delta must be positive");
    if (c.discrete_yaw_steps != -1 && c.discrete_yaw_steps < 1)
        return fail(Q1_EINVAL, "discrete_yaw_steps must be -1 or >= 1");
    if (c.key_press_delay < 0)
        return fail(Q1_EINVAL, "key_press_delay must be >= 0");
    P.n = c.num_envs;
    P.dt = c.time_delta;
    P.time_limit = c.time_limit;
    P.key_delay = c.key_press_delay;
    /* env:230 `_MAX_YAW_SPEED * time_delta`, np.float32(720) times a Python float: float32 under
     * NumPy 2 (NEP 50, what the oracle and the golden fixtures were recorded with), float64 under the
     * NumPy 1.18 the reference pins (requirements_q1physrl.txt:33).  Equal only when time_delta is a
     * float32 value such as nothing the shipped configs use except through rounding: 1/72 and
     * 0.013888888888888 give 10 +- 4e-7 vs 10 exactly. */
    P.max_yaw_delta = numpy1_promotion ? 720.0 * c.time_delta : (double)(720.0f * (float)c.time_delta);
    P.action_range = c.action_range;
    P.yaw_steps = (double)c.discrete_yaw_steps;
    P.accel_dt = 10.0 * c.time_delta;
    P.gravity_dt = 800.0 * c.time_delta;
    P.fmove_half = std::trunc((double)(float)c.fmove_max * 0.5);
    P.fmove_full = std::trunc((double)(float)c.fmove_max);
    P.smove_half = std::trunc((double)(float)c.smove_max * 0.5);
    P.smove_full = std::trunc((double)(float)c.smove_max);
    P.rcp_action_range = 1.0 / c.action_range;
    P.rcp_yaw_steps = 1.0 / (double)c.discrete_yaw_steps;
    P.rcp_time_limit = 1.0 / c.time_limit;
    P.fmove_tab[0] = 0.0;
    P.fmove_tab[1] = P.fmove_half;
    P.fmove_tab[2] = P.fmove_full;
    P.smove_tab[0] = -P.smove_full;
    P.smove_tab[1] = -P.smove_half;
    P.smove_tab[2] = 0.0;
    P.smove_tab[3] = P.smove_half;
    P.smove_tab[4] = P.smove_full;
    P.zero_start_prob = c.zero_start_prob;
    P.yaw_lo = c.initial_yaw_lo;
    P.yaw_hi = c.initial_yaw_hi;
    P.max_initial_speed = c.max_initial_speed;
    P.dt_f32 = (float)c.time_delta;
    P.num_keys = q1_num_keys(&c);
    P.allow_yaw = c.allow_yaw != 0;
    P.discrete_yaw = c.discrete_yaw_steps != -1;
    P.speed_reward = c.speed_reward != 0;
    P.hover = c.hover != 0;
    P.smooth_keys = c.smooth_keys != 0;
    P.auto_jump = c.auto_jump != 0;
    P.allow_jump = c.allow_jump != 0;
    /* 5-bit countdown timers reproduce the f64 stamp comparison exactly when delay/dt is safely away
     * from an integer (rounding in TL - t_rem is ~1e-12 s), or when delay is 0. */
    double q = c.key_press_delay / c.time_delta;
    double frac = std::fabs(q - std::nearbyint(q));
    counters_exact = (c.key_press_delay == 0.0) || (frac > 1e-6 && q < 30.0); /* 5-bit counters */
    /* A fresh episode must satisfy "elapsed" at once, i.e. time_limit - t_rem >= -delay + delay = 0.
     * Resets draw t_rem = uniform(low=time_limit, high=1.0) (env:439, 466), which stays <= time_limit
     * only for time_limit >= 1; below that the f64 stamps are kept. */
    counters_exact = counters_exact && c.time_limit >= 1.0;
    P.delay_ticks = counters_exact ? (int32_t)std::ceil(q) : 0;
    P.press_ticks = P.delay_ticks > 0 ? P.delay_ticks - 1 : 0;
    P.jump_mode = c.auto_jump ? 2 : (c.allow_jump ? 1 : 0);
    return Q1_OK;
}

size_t align_up(size_t v) { return (v + 255u) & ~(size_t)255u; }

template <typename F> int dispatch(const q1_env *env, F &&f)
{
    auto with_lean = [&](auto st, auto tr) {
        return env->P.ieee_div ? f(st, tr, std::false_type{}) : f(st, tr, std::true_type{});
    };
    if (env->stamps)
        return env->track ? with_lean(std::true_type{}, std::true_type{})
                          : with_lean(std::true_type{}, std::false_type{});
    return env->track ? with_lean(std::false_type{}, std::true_type{})
                      : with_lean(std::false_type{}, std::false_type{});
}

int check_launch(const char *what)
{
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
        return fail(Q1_ECUDA, std::string(what) + " launch: " + cudaGetErrorString(err));
    return Q1_OK;
}

/* The address under which the device reaches the page-locked host buffer `p` (the same address
 * under unified addressing, possibly another one for cudaHostRegister'ed memory); NULL if the
 * buffer is not mapped into the device's address space. */
template <typename T> T *device_view(T *p)
{
    void *d = nullptr;
    if (cudaHostGetDevicePointer(&d, const_cast<void *>(static_cast<const void *>(p)), 0) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return static_cast<T *>(d);
}

int ensure_scratch(q1_env *env, size_t bytes)
{
    if (!env->host_stream)
        Q1_CUDA(cudaStreamCreateWithFlags(&env->host_stream, cudaStreamNonBlocking));
    if (env->scratch_bytes < bytes) {
        if (env->scratch)
            Q1_CUDA(cudaFree(env->scratch));
        env->scratch = nullptr;
        env->scratch_bytes = 0;
        Q1_CUDA(cudaMalloc(&env->scratch, bytes));
        env->scratch_bytes = bytes;
    }
    return Q1_OK;
}

} // namespace

namespace {
/* Bump allocator over one device buffer for the *_host conveniences without a handle. */
struct Staging {
    char *base = nullptr;
    size_t size = 0, off = 0;
    ~Staging()
    {
        if (base)
            cudaFree(base);
    }
    int reserve(size_t bytes)
    {
        Q1_CUDA(cudaMalloc(reinterpret_cast<void **>(&base), bytes));
        size = bytes;
        return Q1_OK;
    }
    template <typename T> T *take(size_t count)
    {
        T *p = reinterpret_cast<T *>(base + off);
        off = align_up(off + count * sizeof(T));
        return p;
    }
};
} // namespace

extern "C" {

const char *q1_last_error(void) { return g_error.c_str(); }

int q1_abi_version(void) { return Q1_ABI_VERSION; }

int q1_device_count(int *count)
{
    if (!count)
        return fail(Q1_EINVAL, "count is NULL");
    *count = 0;
    cudaError_t err = cudaGetDeviceCount(count);
    if (err != cudaSuccess) {
        *count = 0;
        return fail(Q1_ENODEV, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(err));
    }
    return Q1_OK;
}

int q1_num_keys(const q1_config *cfg)
{
    if (!cfg)
        return fail(Q1_EINVAL, "cfg is NULL");
    return (!cfg->auto_jump && cfg->allow_jump) ? 4 : 3; /* env:206-207 */
}

int q1_create(const q1_config *cfg, int device, uint64_t seed, uint64_t env_index_base,
              uint32_t flags, q1_env **out)
{
    if (!cfg || !out)
        return fail(Q1_EINVAL, "cfg / out is NULL");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(Q1_ENODEV, "no CUDA device: libq1phys has no CPU implementation");
    if (device < 0 || device >= count)
        return fail(Q1_EINVAL, "device index out of range");
    q1_env *env = new (std::nothrow) q1_env();
    if (!env)
        return fail(Q1_ENOMEM, "out of host memory");
    env->cfg = *cfg;
    env->device = device;
    env->flags = flags;
    bool counters_exact = false;
    int rc = derive_params(*cfg, env->P, counters_exact, (flags & Q1_F_NUMPY1_PROMOTION) != 0);
    if (rc != Q1_OK) {
        delete env;
        return rc;
    }
    env->stamps = (flags & Q1_F_FORCE_F64_STAMPS) || !counters_exact;
    env->track = flags & Q1_F_TRACK_RETURNS;
    env->pdl = getenv("Q1PHYS_NO_PDL") == nullptr;
    if (const char *hc = getenv("Q1PHYS_HOST_CHUNKS"))
        env->host_chunks = std::max(1, std::min(8, atoi(hc)));
    if (const char *bg = getenv("Q1PHYS_BALANCE_GRID"))
        env->balance_grid = atoi(bg) != 0;
    if (const char *hd = getenv("Q1PHYS_HOST_DIRECT"))
        env->host_direct = atoi(hd) != 0;
    env->small_launch_tiles = 4 * 7 * 148;    /* up to four waves of the small build (measured: see k_step_tma) */
    if (const char *sl = getenv("Q1PHYS_SMALL_TILES"))
        env->small_launch_tiles = atoll(sl);
    /* the reciprocal sequences assume positive divisors in a sane exponent range */
    auto sane = [](double v) { return v > 1e-100 && v < 1e100; };
    /* ... and that q = RN(a * RN(1/b)) is a faithful quotient (div_const3 in q1_tick.cuh) */
    auto short_division_ok = [](double b) { return std::fabs(std::fma(1.0 / b, b, -1.0)) <= 0x1p-54; };
    auto divisor_ok = [&](double b) { return sane(b) && short_division_ok(b); };
    env->P.ieee_div = (flags & Q1_F_IEEE_DIVISION) || !divisor_ok(cfg->time_limit) ||
                      (cfg->allow_yaw && cfg->discrete_yaw_steps == -1 && !divisor_ok(cfg->action_range)) ||
                      (cfg->allow_yaw && cfg->discrete_yaw_steps != -1 &&
                       !divisor_ok((double)cfg->discrete_yaw_steps));
    Params &P = env->P;
    P.seed = seed;
    P.env_index_base = env_index_base;

    const size_t n = (size_t)P.n;
    const int nk = P.num_keys;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes);
        return o;
    };
    const size_t blocks = (n + kTile - 1) / kTile;
    size_t o_state = take(blocks * kTileBytes);
    size_t o_stamps = env->stamps ? take(8 * n * nk) : 0;
    size_t o_epoch = take(4 * n);
    size_t o_ret = env->track ? take(8 * n) : 0;
    size_t o_metrics = env->track ? take(64) : 0;
    env->pool_bytes = off;
    env->state_bytes_per_env = 16 + 16 + 8 + (env->stamps ? 8 * nk : 0) + (env->track ? 8 : 0);

    DeviceGuard guard(device);
    if (!guard.ok) {
        delete env;
        return fail(Q1_ECUDA, "cudaSetDevice failed");
    }
    if (cudaDeviceGetAttribute(&env->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
        env->sm_count <= 0) {
        delete env;
        return fail(Q1_ECUDA, "cudaDeviceGetAttribute(multiProcessorCount) failed");
    }
    /* kStepCtasPerSm CTAs x 21 KB of staging must fit: ask for the large shared-memory carveout */
    {
#define Q1_TMA_VARIANTS(C)                                                                              \
    (const void *)k_step_tma<false, false, false, C>, (const void *)k_step_tma<false, false, true, C>,      \
    (const void *)k_step_tma<false, true, false, C>, (const void *)k_step_tma<false, true, true, C>,        \
    (const void *)k_step_tma<true, false, false, C>, (const void *)k_step_tma<true, false, true, C>,        \
    (const void *)k_step_tma<true, true, false, C>, (const void *)k_step_tma<true, true, true, C>
        const void *fns[] = {Q1_TMA_VARIANTS(kStepCtasPerSm), Q1_TMA_VARIANTS(kStepCtasSmall)};
#undef Q1_TMA_VARIANTS
        for (const void *f : fns) {
            cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            if (e != cudaSuccess) {
                delete env;
                return fail(Q1_ECUDA, std::string("cudaFuncSetAttribute(shared-memory carveout): ") +
                                          cudaGetErrorString(e));
            }
        }
    }
    cudaError_t err = cudaMalloc(&env->pool, env->pool_bytes);
    if (err != cudaSuccess) {
        delete env;
        return fail(err == cudaErrorMemoryAllocation ? Q1_ENOMEM : Q1_ECUDA,
                    std::string("cudaMalloc: ") + cudaGetErrorString(err));
    }
    err = cudaMemset(env->pool, 0, env->pool_bytes);
    if (err != cudaSuccess) {
        cudaFree(env->pool);
        delete env;
        return fail(Q1_ECUDA, std::string("cudaMemset: ") + cudaGetErrorString(err));
    }
    char *base = static_cast<char *>(env->pool);
    P.state = reinterpret_cast<unsigned char *>(base + o_state);
    P.stamps = env->stamps ? reinterpret_cast<double *>(base + o_stamps) : nullptr;
    P.epoch = reinterpret_cast<uint32_t *>(base + o_epoch);
    P.ep_return = env->track ? reinterpret_cast<double *>(base + o_ret) : nullptr;
    P.metrics = env->track ? reinterpret_cast<double *>(base + o_metrics) : nullptr;
    *out = env;
    return Q1_OK;
}

int q1_destroy(q1_env *env)
{
    if (!env)
        return Q1_OK;
    DeviceGuard guard(env->device);
    /* everything is released whatever fails; the first failure is what the caller hears about */
    cudaError_t first = cudaSuccess;
    auto note = [&](cudaError_t e) {
        if (e != cudaSuccess && first == cudaSuccess)
            first = e;
    };
    if (env->scratch)
        note(cudaFree(env->scratch));
    if (env->bounce)
        note(cudaFreeHost(env->bounce));
    if (env->host_stream)
        note(cudaStreamDestroy(env->host_stream));
    if (env->in_stream) {
        note(cudaStreamDestroy(env->in_stream));
        note(cudaStreamDestroy(env->out_stream));
        for (int c = 0; c < 8; c++) {
            note(cudaEventDestroy(env->ev_in[c]));
            note(cudaEventDestroy(env->ev_done[c]));
        }
    }
    if (env->pool)
        note(cudaFree(env->pool));
    delete env;
    if (first != cudaSuccess)
        return fail(Q1_ECUDA, std::string("q1_destroy: ") + cudaGetErrorString(first));
    return Q1_OK;
}

int q1_info(const q1_env *env, q1_env_info *out)
{
    if (!env || !out)
        return fail(Q1_EINVAL, "env / out is NULL");
    out->num_envs = env->P.n;
    out->num_keys = env->P.num_keys;
    out->device = env->device;
    out->f64_stamps = env->stamps;
    out->track_returns = env->track;
    out->key_delay_ticks = env->P.delay_ticks;
    out->state_bytes_per_env = env->state_bytes_per_env;
    out->env_index_base = env->P.env_index_base;
    out->seed = env->P.seed;
    out->ticks = env->ticks;
    return Q1_OK;
}

int q1_sync(q1_env *env, void *stream)
{
    if (!env)
        return fail(Q1_EINVAL, "env is NULL");
    DeviceGuard guard(env->device);
    Q1_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return Q1_OK;
}

static int reset_launch(q1_env *env, const uint8_t *mask, int64_t only, float *obs,
                        int64_t obs_row_offset, cudaStream_t s)
{
    unsigned grid = only >= 0 ? 1u : grid_for(env->P.n);
    return dispatch(env, [&](auto st, auto tr, auto ln) {
        k_reset<decltype(st)::value, decltype(tr)::value, decltype(ln)::value>
            <<<grid, kBlock, 0, s>>>(env->P, mask, only, obs, obs_row_offset);
        return check_launch("k_reset");
    });
}

int q1_reset_all(q1_env *env, float *obs, void *stream)
{
    if (!env)
        return fail(Q1_EINVAL, "env is NULL");
    DeviceGuard guard(env->device);
    env->caller_streams_used = true;
    return reset_launch(env, nullptr, -1, obs, 0, static_cast<cudaStream_t>(stream));
}

int q1_reset_masked(q1_env *env, const uint8_t *mask, float *obs, void *stream)
{
    if (!env || !mask)
        return fail(Q1_EINVAL, "env / mask is NULL");
    DeviceGuard guard(env->device);
    env->caller_streams_used = true;
    return reset_launch(env, mask, -1, obs, 0, static_cast<cudaStream_t>(stream));
}

/* mask_host == NULL: all envs.  Rows of envs that are not reset are left as they are in obs_host. */
static int reset_host(q1_env *env, const uint8_t *mask_host, float *obs_host)
{
    DeviceGuard guard(env->device);
    const size_t n = (size_t)env->P.n;
    size_t o_obs = align_up(n);
    int rc = ensure_scratch(env, o_obs + align_up(24 * n));
    if (rc == Q1_OK)
        rc = order_after_caller_streams(env);
    if (rc != Q1_OK)
        return rc;
    char *d = static_cast<char *>(env->scratch);
    cudaStream_t s = env->host_stream;
    uint8_t *d_mask = nullptr;
    float *d_obs = obs_host ? reinterpret_cast<float *>(d + o_obs) : nullptr;
    if (mask_host) {
        d_mask = reinterpret_cast<uint8_t *>(d);
        Q1_CUDA(cudaMemcpyAsync(d_mask, mask_host, n, cudaMemcpyHostToDevice, s));
        if (obs_host)
            Q1_CUDA(cudaMemcpyAsync(d_obs, obs_host, 24 * n, cudaMemcpyHostToDevice, s));
    }
    rc = reset_launch(env, d_mask, -1, d_obs, 0, s);
    if (rc != Q1_OK)
        return rc;
    if (obs_host)
        Q1_CUDA(cudaMemcpyAsync(obs_host, d_obs, 24 * n, cudaMemcpyDeviceToHost, s));
    Q1_CUDA(cudaStreamSynchronize(s));
    return Q1_OK;
}

int q1_reset_all_host(q1_env *env, float *obs_host)
{
    if (!env)
        return fail(Q1_EINVAL, "env is NULL");
    return reset_host(env, nullptr, obs_host);
}

int q1_reset_masked_host(q1_env *env, const uint8_t *mask_host, float *obs_host)
{
    if (!env || !mask_host)
        return fail(Q1_EINVAL, "env / mask is NULL");
    return reset_host(env, mask_host, obs_host);
}

int q1_reset_at_host(q1_env *env, int64_t index, float *obs6_host)
{
    if (!env)
        return fail(Q1_EINVAL, "env is NULL");
    if (index < 0 || index >= env->P.n)
        return fail(Q1_EINVAL, "env index out of range");
    DeviceGuard guard(env->device);
    int rc = ensure_scratch(env, 64);
    if (rc == Q1_OK)
        rc = order_after_caller_streams(env);
    if (rc != Q1_OK)
        return rc;
    float *d_obs = static_cast<float *>(env->scratch);
    rc = reset_launch(env, nullptr, index, obs6_host ? d_obs : nullptr, -index, env->host_stream);
    if (rc != Q1_OK)
        return rc;
    if (obs6_host)
        Q1_CUDA(cudaMemcpyAsync(obs6_host, d_obs, 6 * sizeof(float), cudaMemcpyDeviceToHost,
                                env->host_stream));
    Q1_CUDA(cudaStreamSynchronize(env->host_stream));
    return Q1_OK;
}

/* One tick for envs [begin, end) of the handle (begin a multiple of kTile); the buffers are the
 * full-size arrays, indexed by absolute env. */
static int step_range(q1_env *env, const uint8_t *keys, const void *mouse, int mouse_kind, float *obs,
                      float *reward, uint8_t *done, uint8_t *zero_start, int auto_reset,
                      int64_t begin, int64_t end, cudaStream_t s)
{
    auto aligned16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    /* full tiles go through the TMA-pipelined kernel when every buffer it bulk-copies is 16-byte
     * aligned (tile strides are multiples of 16 by construction); the rest takes the plain kernel */
    const int64_t tile_begin = begin / kBlock;
    int64_t tile_end = tile_begin;
    if (!env->stamps && mouse_kind != Q1_MOUSE_F64 && aligned16(keys) && aligned16(obs) &&
        aligned16(reward) && aligned16(done) &&
        (!zero_start || aligned16(zero_start)) && (!env->P.allow_yaw || aligned16(mouse)))
        tile_end = end / kBlock;
    /* a ragged tail costs a second launch (~5 us after the first); below 2^18 envs one launch of the
     * plain kernel over everything is the faster way to serve it */
    if (tile_end > tile_begin && end % kBlock != 0 && tile_end - tile_begin < 2048)
        tile_end = tile_begin;
    int rc = Q1_OK;
    if (tile_end > tile_begin) {
        /* the compiled-in configuration: continuous f32 mouse action, no hover, y-velocity reward */
        const bool common = env->P.allow_yaw && !env->P.discrete_yaw && !env->P.hover &&
                            !env->P.speed_reward && mouse_kind == Q1_MOUSE_F32;
        rc = dispatch(env, [&](auto, auto tr, auto ln) {
            const int64_t ntiles = tile_end - tile_begin;
            const bool small = ntiles <= env->small_launch_tiles;
            const int64_t resident = (int64_t)env->sm_count * (small ? kStepCtasSmall : kStepCtasPerSm);
            int64_t g = std::min<int64_t>(ntiles, resident);
            if (env->balance_grid && ntiles > resident) {
                /* same number of rounds, but every CTA walks (almost) the same number of tiles */
                const int64_t rounds = (ntiles + resident - 1) / resident;
                g = (ntiles + rounds - 1) / rounds;
            }
            unsigned grid = (unsigned)g;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(kBlock);
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = env->pdl ? 1 : 0;
            constexpr bool TR = decltype(tr)::value, LN = decltype(ln)::value;
            auto launch = [&](auto kernel) {
                return cudaLaunchKernelEx(&cfg, kernel, env->P, keys, mouse, mouse_kind, obs, reward, done,
                                          zero_start, auto_reset, tile_begin, tile_end);
            };
            cudaError_t err;
            if (small)
                err = common ? launch(k_step_tma<TR, LN, true, kStepCtasSmall>)
                             : launch(k_step_tma<TR, LN, false, kStepCtasSmall>);
            else
                err = common ? launch(k_step_tma<TR, LN, true, kStepCtasPerSm>)
                             : launch(k_step_tma<TR, LN, false, kStepCtasPerSm>);
            if (err != cudaSuccess)
                return fail(Q1_ECUDA, std::string("k_step_tma launch: ") + cudaGetErrorString(err));
            return check_launch("k_step_tma");
        });
    }
    const int64_t first = tile_end > tile_begin ? tile_end * kBlock : begin;
    if (rc == Q1_OK && first < end)
        rc = dispatch(env, [&](auto st, auto tr, auto ln) {
            /* launched as a programmatic dependent as well: after a TMA launch only its execution,
             * not its launch latency, is added to the tick (it waits for that launch to complete) */
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid_for(end - first));
            cfg.blockDim = dim3(kBlock);
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = env->pdl ? 1 : 0;
            cudaError_t err = cudaLaunchKernelEx(
                &cfg, k_step<decltype(st)::value, decltype(tr)::value, decltype(ln)::value>, env->P, keys,
                mouse, mouse_kind, obs, reward, done, zero_start, auto_reset, first, (int64_t)end);
            if (err != cudaSuccess)
                return fail(Q1_ECUDA, std::string("k_step launch: ") + cudaGetErrorString(err));
            return check_launch("k_step");
        });
    return rc;
}

int q1_step(q1_env *env, const uint8_t *keys, const void *mouse, int mouse_kind, float *obs,
            float *reward, uint8_t *done, uint8_t *zero_start, int auto_reset, void *stream)
{
    if (mouse_kind != Q1_MOUSE_F32 && mouse_kind != Q1_MOUSE_I32 && mouse_kind != Q1_MOUSE_F64)
        return fail(Q1_EINVAL, "unknown mouse_kind");
    if (!env || !keys || !obs || !reward || !done)
        return fail(Q1_EINVAL, "env / keys / obs / reward / done is NULL");
    if (env->P.allow_yaw && !mouse)
        return fail(Q1_EINVAL, "mouse is NULL but allow_yaw is set");
    DeviceGuard guard(env->device);
    env->caller_streams_used = true;
    int rc = step_range(env, keys, mouse, mouse_kind, obs, reward, done, zero_start, auto_reset, 0,
                        env->P.n, static_cast<cudaStream_t>(stream));
    if (rc == Q1_OK)
        env->ticks += 1;
    return rc;
}

int q1_host_alloc(uint64_t bytes, void **out)
{
    if (!out)
        return fail(Q1_EINVAL, "out is NULL");
    *out = nullptr;
    cudaError_t err = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (err != cudaSuccess)
        return fail(err == cudaErrorMemoryAllocation ? Q1_ENOMEM : Q1_ECUDA,
                    std::string("cudaHostAlloc: ") + cudaGetErrorString(err));
    return Q1_OK;
}

int q1_host_free(void *ptr)
{
    if (ptr)
        Q1_CUDA(cudaFreeHost(ptr));
    return Q1_OK;
}

/* Is `p` page-locked (cudaHostAlloc / cudaHostRegister) memory, i.e. can a copy be truly async? */
static bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int q1_step_host(q1_env *env, const uint8_t *keys, const void *mouse, int mouse_kind, float *obs,
                 float *reward, uint8_t *done, uint8_t *zero_start, int auto_reset)
{
    if (mouse_kind != Q1_MOUSE_F32 && mouse_kind != Q1_MOUSE_I32 && mouse_kind != Q1_MOUSE_F64)
        return fail(Q1_EINVAL, "unknown mouse_kind");
    const size_t mouse_size = mouse_kind == Q1_MOUSE_F64 ? 8 : 4;
    if (!env || !keys || !obs || !reward || !done)
        return fail(Q1_EINVAL, "env / keys / obs / reward / done is NULL");
    if (env->P.allow_yaw && !mouse)
        return fail(Q1_EINVAL, "mouse is NULL but allow_yaw is set");
    DeviceGuard guard(env->device);
    const size_t n = (size_t)env->P.n, nk = (size_t)env->P.num_keys;
    size_t o_keys = 0, o_mouse = align_up(n * nk), o_obs = o_mouse + align_up(8 * n);
    size_t o_rew = o_obs + align_up(24 * n), o_done = o_rew + align_up(4 * n);
    size_t o_zs = o_done + align_up(n), total = o_zs + align_up(n);
    if (!env->host_stream)
        Q1_CUDA(cudaStreamCreateWithFlags(&env->host_stream, cudaStreamNonBlocking));
    cudaStream_t s = env->host_stream;
    int rc = order_after_caller_streams(env);
    if (rc != Q1_OK)
        return rc;

    /* Up to kBounceAlways envs -- RLLib's 100 per worker, the gym-style single env -- the bounce route
     * below is taken without asking the driver what kind of memory the seven buffers are (six pointer
     * queries cost more than copying a few KB); larger batches are worth the question. */
    constexpr size_t kBounceAlways = 8192;
    const bool ask = n > kBounceAlways || !env->host_direct;
    const bool all_pinned = ask && is_pinned(keys) && is_pinned(obs) && is_pinned(reward) && is_pinned(done) &&
                            (!env->P.allow_yaw || is_pinned(mouse)) && (!zero_start || is_pinned(zero_start));
    if (all_pinned && env->host_direct) {
        /* Page-locked buffers are mapped into the device's address space: the step kernel bulk-loads
         * the actions from host memory and bulk-stores the results to host memory itself.  Both PCIe
         * directions then run concurrently for the whole launch, tile by tile, with no staging copy,
         * no chunk boundaries and one launch: a step costs max(H2D, D2H) of the wire. */
        const uint8_t *m_keys = device_view(keys);
        const void *m_mouse = env->P.allow_yaw ? device_view(mouse) : nullptr;
        float *m_obs = device_view(obs), *m_rew = device_view(reward);
        uint8_t *m_done = device_view(done), *m_zs = zero_start ? device_view(zero_start) : nullptr;
        if (m_keys && m_obs && m_rew && m_done && (!env->P.allow_yaw || m_mouse) && (!zero_start || m_zs)) {
            rc = step_range(env, m_keys, m_mouse, mouse_kind, m_obs, m_rew, m_done, m_zs, auto_reset, 0,
                            (int64_t)n, s);
            if (rc != Q1_OK)
                return rc;
            Q1_CUDA(cudaStreamSynchronize(s));
            env->ticks += 1;
            return Q1_OK;
        }
    }
    if (!all_pinned && env->host_direct && n <= (size_t)1 << 16) {
        /* Small batches in ordinary (pageable) memory -- RLLib's 100 envs per worker, the gym-style
         * single env: seven staged cudaMemcpy calls cost far more than the tick.  Bounce through one
         * page-locked, device-mapped buffer owned by the handle instead: two host memcpys in, ONE
         * launch that reads and writes that buffer across PCIe, one synchronisation, four memcpys out. */
        if (env->bounce_bytes < total) {
            if (env->bounce)
                cudaFreeHost(env->bounce);
            env->bounce = nullptr;
            env->bounce_bytes = 0;
            Q1_CUDA(cudaHostAlloc(&env->bounce, total, cudaHostAllocMapped));
            env->bounce_bytes = total;
            env->bounce_dev = device_view(static_cast<char *>(env->bounce));
        }
        char *h = static_cast<char *>(env->bounce);
        char *m = env->bounce_dev;
        if (m) {
            std::memcpy(h + o_keys, keys, n * nk);
            if (env->P.allow_yaw)
                std::memcpy(h + o_mouse, mouse, mouse_size * n);
            rc = step_range(env, reinterpret_cast<const uint8_t *>(m + o_keys), m + o_mouse, mouse_kind,
                            reinterpret_cast<float *>(m + o_obs), reinterpret_cast<float *>(m + o_rew),
                            reinterpret_cast<uint8_t *>(m + o_done),
                            zero_start ? reinterpret_cast<uint8_t *>(m + o_zs) : nullptr, auto_reset, 0,
                            (int64_t)n, s);
            if (rc != Q1_OK)
                return rc;
            Q1_CUDA(cudaStreamSynchronize(s));
            std::memcpy(obs, h + o_obs, 24 * n);
            std::memcpy(reward, h + o_rew, 4 * n);
            std::memcpy(done, h + o_done, n);
            if (zero_start)
                std::memcpy(zero_start, h + o_zs, n);
            env->ticks += 1;
            return Q1_OK;
        }
    }

    rc = ensure_scratch(env, total);
    if (rc != Q1_OK)
        return rc;
    char *d = static_cast<char *>(env->scratch);
    uint8_t *d_keys = reinterpret_cast<uint8_t *>(d + o_keys);
    char *d_mouse = d + o_mouse;
    float *d_obs = reinterpret_cast<float *>(d + o_obs), *d_rew = reinterpret_cast<float *>(d + o_rew);
    uint8_t *d_done = reinterpret_cast<uint8_t *>(d + o_done);
    uint8_t *d_zs = zero_start ? reinterpret_cast<uint8_t *>(d + o_zs) : nullptr;

    /* Large batches in page-locked memory with Q1PHYS_HOST_DIRECT=0 run as a pipeline over env
     * chunks: the action upload of chunk c+1, the tick of chunk c and the result download of chunk
     * c-1 overlap (the two PCIe directions are independent copy engines). */
    constexpr int kMaxChunks = 8;
    int chunks = 1;
    if (n >= (size_t)1 << 16 && all_pinned)
        chunks = env->host_chunks;
    if (chunks <= 1) {
        Q1_CUDA(cudaMemcpyAsync(d_keys, keys, n * nk, cudaMemcpyHostToDevice, s));
        if (env->P.allow_yaw)
            Q1_CUDA(cudaMemcpyAsync(d_mouse, mouse, mouse_size * n, cudaMemcpyHostToDevice, s));
        rc = step_range(env, d_keys, d_mouse, mouse_kind, d_obs, d_rew, d_done, d_zs, auto_reset, 0,
                        (int64_t)n, s);
        if (rc != Q1_OK)
            return rc;
        Q1_CUDA(cudaMemcpyAsync(obs, d_obs, 24 * n, cudaMemcpyDeviceToHost, s));
        Q1_CUDA(cudaMemcpyAsync(reward, d_rew, 4 * n, cudaMemcpyDeviceToHost, s));
        Q1_CUDA(cudaMemcpyAsync(done, d_done, n, cudaMemcpyDeviceToHost, s));
        if (zero_start)
            Q1_CUDA(cudaMemcpyAsync(zero_start, d_zs, n, cudaMemcpyDeviceToHost, s));
        Q1_CUDA(cudaStreamSynchronize(s));
        env->ticks += 1;
        return Q1_OK;
    }
    if (!env->in_stream) {
        Q1_CUDA(cudaStreamCreateWithFlags(&env->in_stream, cudaStreamNonBlocking));
        Q1_CUDA(cudaStreamCreateWithFlags(&env->out_stream, cudaStreamNonBlocking));
        for (int c = 0; c < kMaxChunks; c++) {
            Q1_CUDA(cudaEventCreateWithFlags(&env->ev_in[c], cudaEventDisableTiming));
            Q1_CUDA(cudaEventCreateWithFlags(&env->ev_done[c], cudaEventDisableTiming));
        }
    }
    /* chunk boundaries on 2 * kTile envs so that every sub-array offset stays 16-byte aligned */
    const size_t grain = 2 * kTile;
    const size_t per = ((n / chunks) / grain) * grain;
    auto chunk_begin = [&](int c) { return c == 0 ? (size_t)0 : (c >= chunks ? n : per * c); };
    for (int c = 0; c < chunks; c++) {
        const size_t b = chunk_begin(c), e = chunk_begin(c + 1);
        Q1_CUDA(cudaMemcpyAsync(d_keys + b * nk, keys + b * nk, (e - b) * nk, cudaMemcpyHostToDevice,
                                env->in_stream));
        if (env->P.allow_yaw)
            Q1_CUDA(cudaMemcpyAsync(d_mouse + b * mouse_size, static_cast<const char *>(mouse) + b * mouse_size,
                                    (e - b) * mouse_size, cudaMemcpyHostToDevice, env->in_stream));
        Q1_CUDA(cudaEventRecord(env->ev_in[c], env->in_stream));
    }
    for (int c = 0; c < chunks; c++) {
        const size_t b = chunk_begin(c), e = chunk_begin(c + 1);
        Q1_CUDA(cudaStreamWaitEvent(s, env->ev_in[c], 0));
        rc = step_range(env, d_keys, d_mouse, mouse_kind, d_obs, d_rew, d_done, d_zs, auto_reset,
                        (int64_t)b, (int64_t)e, s);
        if (rc != Q1_OK)
            return rc;
        Q1_CUDA(cudaEventRecord(env->ev_done[c], s));
        Q1_CUDA(cudaStreamWaitEvent(env->out_stream, env->ev_done[c], 0));
        Q1_CUDA(cudaMemcpyAsync(obs + 6 * b, d_obs + 6 * b, 24 * (e - b), cudaMemcpyDeviceToHost, env->out_stream));
        Q1_CUDA(cudaMemcpyAsync(reward + b, d_rew + b, 4 * (e - b), cudaMemcpyDeviceToHost, env->out_stream));
        Q1_CUDA(cudaMemcpyAsync(done + b, d_done + b, e - b, cudaMemcpyDeviceToHost, env->out_stream));
        if (zero_start)
            Q1_CUDA(cudaMemcpyAsync(zero_start + b, d_zs + b, e - b, cudaMemcpyDeviceToHost, env->out_stream));
    }
    Q1_CUDA(cudaStreamSynchronize(env->out_stream));
    Q1_CUDA(cudaStreamSynchronize(s));
    env->ticks += 1;
    return Q1_OK;
}

static int rollout_launch(q1_env *env, const ActionFeed &feed, int ticks, float *obs, float *reward_sum,
                          int auto_reset, uint32_t record_flags, const q1_record_view *rec, cudaStream_t s)
{
    int rc = dispatch(env, [&](auto st, auto tr, auto ln) {
        if (rec)
            k_rollout<decltype(st)::value, decltype(tr)::value, decltype(ln)::value, true>
                <<<grid_for(env->P.n), kBlock, 0, s>>>(env->P, feed, ticks, (uint32_t)env->ticks, obs,
                                                       reward_sum, auto_reset, record_flags, *rec);
        else
            k_rollout<decltype(st)::value, decltype(tr)::value, decltype(ln)::value, false>
                <<<grid_for(env->P.n), kBlock, 0, s>>>(env->P, feed, ticks, (uint32_t)env->ticks, obs,
                                                       reward_sum, 1, 0u, q1_record_view{});
        return check_launch("k_rollout");
    });
    if (rc == Q1_OK)
        env->ticks += (uint64_t)ticks;
    return rc;
}

int q1_rollout(q1_env *env, int policy, int ticks, uint64_t policy_seed, float *obs,
               float *reward_sum, void *stream)
{
    if (!env)
        return fail(Q1_EINVAL, "env is NULL");
    if (policy != Q1_POLICY_RANDOM && policy != Q1_POLICY_STRAFE_JUMP)
        return fail(Q1_EINVAL, "unknown policy");
    if (ticks < 0)
        return fail(Q1_EINVAL, "ticks must be >= 0");
    DeviceGuard guard(env->device);
    const ActionFeed feed = {policy, policy_seed, nullptr, nullptr, Q1_MOUSE_F32};
    env->caller_streams_used = true;
    return rollout_launch(env, feed, ticks, obs, reward_sum, 1, 0u, nullptr,
                          static_cast<cudaStream_t>(stream));
}

static int check_record_args(const q1_env *env, const q1_action_source *src, int ticks,
                             const q1_record_view *rec)
{
    if (!env || !src || !rec)
        return fail(Q1_EINVAL, "env / actions / record is NULL");
    if (ticks < 0)
        return fail(Q1_EINVAL, "ticks must be >= 0");
    if (src->kind == Q1_ACTIONS_BUILTIN) {
        if (src->builtin_policy != Q1_POLICY_RANDOM && src->builtin_policy != Q1_POLICY_STRAFE_JUMP)
            return fail(Q1_EINVAL, "unknown built-in policy");
    } else if (src->kind == Q1_ACTIONS_ARRAYS) {
        if (src->mouse_kind != Q1_MOUSE_F32 && src->mouse_kind != Q1_MOUSE_I32 && src->mouse_kind != Q1_MOUSE_F64)
            return fail(Q1_EINVAL, "unknown mouse_kind");
        if (ticks > 0 && (!src->keys || (env->P.allow_yaw && !src->mouse)))
            return fail(Q1_EINVAL, "action arrays are NULL");
    } else {
        return fail(Q1_EINVAL, "unknown action source kind");
    }
    return Q1_OK;
}

int q1_rollout_record(q1_env *env, const q1_action_source *actions, int ticks, int auto_reset,
                      uint32_t record_flags, const q1_record_view *record, float *final_obs, void *stream)
{
    int rc = check_record_args(env, actions, ticks, record);
    if (rc != Q1_OK)
        return rc;
    DeviceGuard guard(env->device);
    const ActionFeed feed = {actions->kind == Q1_ACTIONS_BUILTIN ? actions->builtin_policy : -1,
                             actions->policy_seed, actions->keys, actions->mouse, actions->mouse_kind};
    env->caller_streams_used = true;
    return rollout_launch(env, feed, ticks, final_obs, nullptr, auto_reset, record_flags, record,
                          static_cast<cudaStream_t>(stream));
}

int q1_rollout_record_host(q1_env *env, const q1_action_source *actions, int ticks, int auto_reset,
                           uint32_t record_flags, const q1_record_view *record, float *final_obs_host)
{
    int rc = check_record_args(env, actions, ticks, record);
    if (rc != Q1_OK)
        return rc;
    DeviceGuard guard(env->device);
    const size_t n = (size_t)env->P.n, nk = (size_t)env->P.num_keys, rows = n * (size_t)ticks;
    /* one staging buffer: [actions in | every requested record array | final obs] */
    struct Field { const void *host; size_t bytes, off; };
    Field f[15];
    const void *hosts[15] = {record->vel, record->z_pos, record->on_ground, record->jump_released,
                             record->time_remaining, record->obs, record->keys, record->mouse,
                             record->yaw, record->smove, record->fmove, record->jump, record->reward,
                             record->done, final_obs_host};
    const size_t width[15] = {12, 8, 1, 1, 8, 24, nk, 4, 8, 8, 8, 1, 4, 1, 0};
    const bool arrays = actions->kind == Q1_ACTIONS_ARRAYS;
    const size_t mouse_size = actions->mouse_kind == Q1_MOUSE_F64 ? 8 : 4;
    size_t off = 0;
    const size_t o_keys = off;
    off = align_up(off + (arrays ? rows * nk : 0));
    const size_t o_mouse = off;
    off = align_up(off + (arrays && env->P.allow_yaw ? rows * mouse_size : 0));
    for (int k = 0; k < 15; k++) {
        f[k].host = hosts[k];
        f[k].bytes = hosts[k] ? (k == 14 ? 24 * n : rows * width[k]) : 0;
        f[k].off = off;
        off = align_up(off + f[k].bytes);
    }
    rc = ensure_scratch(env, off + 256);
    if (rc == Q1_OK)
        rc = order_after_caller_streams(env);
    if (rc != Q1_OK)
        return rc;
    char *d = static_cast<char *>(env->scratch);
    cudaStream_t s = env->host_stream;
    if (arrays && rows) {
        Q1_CUDA(cudaMemcpyAsync(d + o_keys, actions->keys, rows * nk, cudaMemcpyHostToDevice, s));
        if (env->P.allow_yaw)
            Q1_CUDA(cudaMemcpyAsync(d + o_mouse, actions->mouse, rows * mouse_size, cudaMemcpyHostToDevice, s));
    }
    auto dev = [&](int k) -> void * { return f[k].bytes ? d + f[k].off : nullptr; };
    q1_record_view dv;
    dv.vel = static_cast<float *>(dev(0));
    dv.z_pos = static_cast<double *>(dev(1));
    dv.on_ground = static_cast<uint8_t *>(dev(2));
    dv.jump_released = static_cast<uint8_t *>(dev(3));
    dv.time_remaining = static_cast<double *>(dev(4));
    dv.obs = static_cast<float *>(dev(5));
    dv.keys = static_cast<uint8_t *>(dev(6));
    dv.mouse = static_cast<float *>(dev(7));
    dv.yaw = static_cast<double *>(dev(8));
    dv.smove = static_cast<int64_t *>(dev(9));
    dv.fmove = static_cast<int64_t *>(dev(10));
    dv.jump = static_cast<uint8_t *>(dev(11));
    dv.reward = static_cast<float *>(dev(12));
    dv.done = static_cast<uint8_t *>(dev(13));
    const ActionFeed feed = {arrays ? -1 : actions->builtin_policy, actions->policy_seed,
                             reinterpret_cast<const uint8_t *>(d + o_keys), d + o_mouse, actions->mouse_kind};
    rc = rollout_launch(env, feed, ticks, static_cast<float *>(dev(14)), nullptr, auto_reset, record_flags,
                        &dv, s);
    if (rc != Q1_OK)
        return rc;
    for (int k = 0; k < 15; k++)
        if (f[k].bytes)
            Q1_CUDA(cudaMemcpyAsync(const_cast<void *>(f[k].host), d + f[k].off, f[k].bytes,
                                    cudaMemcpyDeviceToHost, s));
    Q1_CUDA(cudaStreamSynchronize(s));
    return Q1_OK;
}

} /* extern "C" */

int q1_env_get_view(q1_env *env, bool on_caller_stream, q1_env_view *out)
{
    if (!env || !out)
        return fail(Q1_EINVAL, "env is NULL");
    out->P = env->P;
    out->device = env->device;
    out->stamps = env->stamps;
    out->track = env->track;
    out->ticks = env->ticks;
    out->sm_count = env->sm_count;
    if (on_caller_stream)
        env->caller_streams_used = true;
    return Q1_OK;
}

int q1_env_host_scratch(q1_env *env, size_t bytes, void **scratch, void **stream)
{
    if (!env)
        return fail(Q1_EINVAL, "env is NULL");
    DeviceGuard guard(env->device);
    int rc = ensure_scratch(env, bytes);
    if (rc == Q1_OK)
        rc = order_after_caller_streams(env);
    if (rc != Q1_OK)
        return rc;
    *scratch = env->scratch;
    *stream = env->host_stream;
    return Q1_OK;
}

extern "C" {

int q1_advance_ticks(q1_env *env, int64_t delta)
{
    if (!env)
        return fail(Q1_EINVAL, "env is NULL");
    if (delta < 0 && (uint64_t)(-delta) > env->ticks)
        return fail(Q1_EINVAL, "tick counter would become negative");
    env->ticks += (uint64_t)delta;
    return Q1_OK;
}

static int observe_launch(q1_env *env, float *obs, cudaStream_t s)
{
    return dispatch(env, [&](auto st, auto, auto ln) {
        k_observe<decltype(st)::value, decltype(ln)::value>
            <<<grid_for(env->P.n), kBlock, 0, s>>>(env->P, obs);
        return check_launch("k_observe");
    });
}

int q1_observe(q1_env *env, float *obs, void *stream)
{
    if (!env || !obs)
        return fail(Q1_EINVAL, "env / obs is NULL");
    DeviceGuard guard(env->device);
    env->caller_streams_used = true;
    return observe_launch(env, obs, static_cast<cudaStream_t>(stream));
}

int q1_observe_host(q1_env *env, float *obs_host)
{
    if (!env || !obs_host)
        return fail(Q1_EINVAL, "env / obs is NULL");
    DeviceGuard guard(env->device);
    const size_t n = (size_t)env->P.n;
    int rc = ensure_scratch(env, align_up(24 * n));
    if (rc != Q1_OK)
        return rc;
    float *d_obs = static_cast<float *>(env->scratch);
    rc = order_after_caller_streams(env);
    if (rc != Q1_OK)
        return rc;
    rc = observe_launch(env, d_obs, env->host_stream);
    if (rc != Q1_OK)
        return rc;
    Q1_CUDA(cudaMemcpyAsync(obs_host, d_obs, 24 * n, cudaMemcpyDeviceToHost, env->host_stream));
    Q1_CUDA(cudaStreamSynchronize(env->host_stream));
    return Q1_OK;
}

/* -- state copy-out / copy-in in the reference layout ----------------------------------------- */

/* -- exact checkpoint / resume: the raw device image of a handle ------------------------------ */

namespace {
struct SnapshotHeader {
    uint64_t magic, pool_bytes, ticks, seed, env_index_base;
    int64_t n;
    int32_t num_keys, stamps, track, abi;
    uint64_t config_hash;   /* FNV-1a of the q1_config and the create flags the image was taken under */
};

uint64_t config_hash(const q1_env *env)
{
    uint64_t h = 0xcbf29ce484222325ull;
    auto mix = [&](const void *p, size_t bytes) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < bytes; i++)
            h = (h ^ b[i]) * 0x100000001b3ull;
    };
    q1_config c = env->cfg;
    c.reserved = 0;
    mix(&c, sizeof c);
    const uint32_t flags = env->flags & (Q1_F_NUMPY1_PROMOTION | Q1_F_IEEE_DIVISION);
    mix(&flags, sizeof flags);
    return h;
}
constexpr uint64_t kSnapshotMagic = 0x51315048595353ull; /* "Q1PHYSS" */
} // namespace

int q1_snapshot_bytes(const q1_env *env, uint64_t *bytes)
{
    if (!env || !bytes)
        return fail(Q1_EINVAL, "env / bytes is NULL");
    *bytes = sizeof(SnapshotHeader) + env->pool_bytes;
    return Q1_OK;
}

int q1_snapshot_save_host(q1_env *env, void *buffer, uint64_t bytes)
{
    if (!env || !buffer)
        return fail(Q1_EINVAL, "env / buffer is NULL");
    if (bytes < sizeof(SnapshotHeader) + env->pool_bytes)
        return fail(Q1_EINVAL, "snapshot buffer too small (see q1_snapshot_bytes)");
    DeviceGuard guard(env->device);
    Q1_CUDA(cudaDeviceSynchronize());
    SnapshotHeader h = {kSnapshotMagic, env->pool_bytes, env->ticks, env->P.seed, env->P.env_index_base,
                        env->P.n, env->P.num_keys, env->stamps ? 1 : 0, env->track ? 1 : 0, Q1_ABI_VERSION,
                        config_hash(env)};
    std::memcpy(buffer, &h, sizeof h);
    Q1_CUDA(cudaMemcpy(static_cast<char *>(buffer) + sizeof h, env->pool, env->pool_bytes,
                       cudaMemcpyDeviceToHost));
    return Q1_OK;
}

int q1_snapshot_load_host(q1_env *env, const void *buffer, uint64_t bytes)
{
    if (!env || !buffer)
        return fail(Q1_EINVAL, "env / buffer is NULL");
    SnapshotHeader h;
    if (bytes < sizeof h)
        return fail(Q1_EINVAL, "snapshot truncated");
    std::memcpy(&h, buffer, sizeof h);
    if (h.magic != kSnapshotMagic || h.abi != Q1_ABI_VERSION)
        return fail(Q1_EINVAL, "not a libq1phys snapshot of this ABI version");
    if (h.n != env->P.n || h.num_keys != env->P.num_keys || h.stamps != (env->stamps ? 1 : 0) ||
        h.track != (env->track ? 1 : 0) || h.pool_bytes != env->pool_bytes ||
        bytes < sizeof h + h.pool_bytes)
        return fail(Q1_EINVAL, "snapshot was taken from a handle with another size / key count / flags");
    if (h.config_hash != config_hash(env))
        return fail(Q1_EINVAL, "snapshot was taken under another Config (time_delta, time_limit, key_press_delay "
                               "... differ): it would not continue the run it was saved from");
    DeviceGuard guard(env->device);
    Q1_CUDA(cudaDeviceSynchronize());
    Q1_CUDA(cudaMemcpy(env->pool, static_cast<const char *>(buffer) + sizeof h, env->pool_bytes,
                       cudaMemcpyHostToDevice));
    env->ticks = h.ticks;
    env->P.seed = h.seed;                      /* the reset stream continues where the snapshot left it */
    env->P.env_index_base = h.env_index_base;
    return Q1_OK;
}

int q1_get_state_host(q1_env *env, const q1_state_view *v)
{
    if (!env || !v)
        return fail(Q1_EINVAL, "env / view is NULL");
    DeviceGuard guard(env->device);
    Q1_CUDA(cudaDeviceSynchronize());
    const Params &P = env->P;
    const size_t n = (size_t)P.n;
    const int nk = P.num_keys;
    const size_t blocks = (n + kTile - 1) / kTile;
    std::vector<unsigned char> st(blocks * kTileBytes);
    Q1_CUDA(cudaMemcpy(st.data(), P.state, st.size(), cudaMemcpyDeviceToHost));
    auto rec_a = [&](size_t i) {
        return reinterpret_cast<float4 *>(st.data() + (i / kTile) * kTileBytes + kTileRecA) + i % kTile;
    };
    auto rec_b = [&](size_t i) {
        return reinterpret_cast<double2 *>(st.data() + (i / kTile) * kTileBytes + kTileRecB) + i % kTile;
    };
    auto trem_of = [&](size_t i) {
        return reinterpret_cast<double *>(st.data() + (i / kTile) * kTileBytes + kTileTrem) + i % kTile;
    };
    if (v->vel)
        for (size_t i = 0; i < n; i++) {
            v->vel[3 * i] = rec_a(i)->x;
            v->vel[3 * i + 1] = rec_a(i)->y;
            v->vel[3 * i + 2] = rec_a(i)->z;
        }
    for (size_t i = 0; i < n; i++) {
        if (v->z_pos)
            v->z_pos[i] = rec_b(i)->x;
        if (v->yaw)
            v->yaw[i] = rec_b(i)->y;
        if (v->time_remaining)
            v->time_remaining[i] = *trem_of(i);
    }
    auto bits_of = [&](size_t i) {
        uint32_t w;
        memcpy(&w, &rec_a(i)->w, 4);
        return w;
    };
    if (v->on_ground || v->jump_released || v->zero_start || v->last_keys) {
        for (size_t i = 0; i < n; i++) {
            uint32_t w = bits_of(i);
            if (v->on_ground)
                v->on_ground[i] = (w & F_ON_GROUND) != 0;
            if (v->jump_released)
                v->jump_released[i] = (w & F_JUMP_RELEASED) != 0;
            if (v->zero_start)
                v->zero_start[i] = (w & F_ZERO_START) != 0;
            if (v->last_keys)
                for (int k = 0; k < nk; k++)
                    v->last_keys[i * nk + k] = (w >> (F_LAST_KEY_SHIFT + k)) & 1u;
        }
    }
    if (v->last_press) {
        if (env->stamps) {
            std::vector<double> sp(n * nk);
            Q1_CUDA(cudaMemcpy(sp.data(), P.stamps, 8 * n * nk, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n; i++)
                for (int k = 0; k < nk; k++)
                    v->last_press[i * nk + k] = sp[(size_t)k * n + i];
        } else {
            /* Counter mode keeps "ticks this key stays blocked"; the stamp handed back
             * is the one that yields the same future decode decisions (exact stamps need
             * Q1_F_FORCE_F64_STAMPS). */
            for (size_t i = 0; i < n; i++) {
                double now = P.time_limit - *trem_of(i);
                uint32_t w = bits_of(i);
                for (int k = 0; k < nk; k++) {
                    int r = (w >> (TIMER_BITS * k)) & TIMER_MAX;
                    v->last_press[i * nk + k] =
                        r == 0 ? -P.key_delay : now - (double)(P.delay_ticks - r) * P.dt;
                }
            }
        }
    }
    if (v->episode_return) {
        if (!env->track)
            return fail(Q1_EINVAL, "episode_return requires Q1_F_TRACK_RETURNS");
        Q1_CUDA(cudaMemcpy(v->episode_return, P.ep_return, 8 * n, cudaMemcpyDeviceToHost));
    }
    return Q1_OK;
}

int q1_set_state_host(q1_env *env, const q1_state_view *v)
{
    if (!env || !v)
        return fail(Q1_EINVAL, "env / view is NULL");
    DeviceGuard guard(env->device);
    Q1_CUDA(cudaDeviceSynchronize());
    const Params &P = env->P;
    const size_t n = (size_t)P.n;
    const int nk = P.num_keys;
    {
        const size_t blocks = (n + kTile - 1) / kTile;
        std::vector<unsigned char> st(blocks * kTileBytes);
        Q1_CUDA(cudaMemcpy(st.data(), P.state, st.size(), cudaMemcpyDeviceToHost));
        const bool counters = v->last_press && !env->stamps;
        for (size_t i = 0; i < n; i++) {
            unsigned char *blk = st.data() + (i / kTile) * kTileBytes;
            float4 *ra = reinterpret_cast<float4 *>(blk + kTileRecA) + i % kTile;
            double2 *rb = reinterpret_cast<double2 *>(blk + kTileRecB) + i % kTile;
            double *tr = reinterpret_cast<double *>(blk + kTileTrem) + i % kTile;
            if (v->vel) {
                ra->x = v->vel[3 * i];
                ra->y = v->vel[3 * i + 1];
                ra->z = v->vel[3 * i + 2];
            }
            if (v->z_pos)
                rb->x = v->z_pos[i];
            if (v->yaw)
                rb->y = v->yaw[i];
            if (v->time_remaining)
                *tr = v->time_remaining[i];
            uint32_t x;
            memcpy(&x, &ra->w, 4);
            if (v->on_ground)
                x = (x & ~F_ON_GROUND) | (v->on_ground[i] ? F_ON_GROUND : 0u);
            if (v->jump_released)
                x = (x & ~F_JUMP_RELEASED) | (v->jump_released[i] ? F_JUMP_RELEASED : 0u);
            if (v->zero_start)
                x = (x & ~F_ZERO_START) | (v->zero_start[i] ? F_ZERO_START : 0u);
            if (v->last_keys) {
                x &= ~(0xFu << F_LAST_KEY_SHIFT);
                for (int k = 0; k < nk; k++)
                    x |= (uint32_t)(v->last_keys[i * nk + k] & 1u) << (F_LAST_KEY_SHIFT + k);
            }
            if (counters) {
                /* number of coming ticks j = 0, 1, .. for which now + j * dt >= stamp + delay
                 * (env:241-242) still fails */
                double now = P.time_limit - *tr;
                x &= ~TIMER_FIELD_MASK;
                for (int k = 0; k < nk; k++) {
                    double need = (v->last_press[i * nk + k] + P.key_delay - now) / P.dt;
                    double r = std::ceil(need - 1e-9);
                    if (!(r > 0))
                        r = 0;
                    if (r > TIMER_MAX)
                        r = TIMER_MAX;
                    x |= (uint32_t)r << (TIMER_BITS * k);
                }
            }
            memcpy(&ra->w, &x, 4);
        }
        Q1_CUDA(cudaMemcpy(P.state, st.data(), st.size(), cudaMemcpyHostToDevice));
    }
    if (v->last_press && env->stamps) {
        std::vector<double> st(n * nk);
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < nk; k++)
                st[(size_t)k * n + i] = v->last_press[i * nk + k];
        Q1_CUDA(cudaMemcpy(P.stamps, st.data(), 8 * n * nk, cudaMemcpyHostToDevice));
    }
    if (v->episode_return) {
        if (!env->track)
            return fail(Q1_EINVAL, "episode_return requires Q1_F_TRACK_RETURNS");
        Q1_CUDA(cudaMemcpy(P.ep_return, v->episode_return, 8 * n, cudaMemcpyHostToDevice));
    }
    return Q1_OK;
}

int q1_get_metrics_host(q1_env *env, int clear, q1_metrics *out)
{
    if (!env || !out)
        return fail(Q1_EINVAL, "env / out is NULL");
    if (!env->track)
        return fail(Q1_EINVAL, "metrics require Q1_F_TRACK_RETURNS");
    DeviceGuard guard(env->device);
    Q1_CUDA(cudaDeviceSynchronize());
    unsigned long long raw[5];
    Q1_CUDA(cudaMemcpy(raw, env->P.metrics, sizeof(raw), cudaMemcpyDeviceToHost));
    memcpy(&out->zero_start_return_sum, &raw[0], 8);
    out->zero_start_episodes = (int64_t)raw[1];
    memcpy(&out->return_sum, &raw[2], 8);
    out->episodes = (int64_t)raw[3];
    if (raw[4] == 0) {
        out->return_max = -INFINITY;
    } else {
        unsigned long long b = (raw[4] >> 63) ? (raw[4] & 0x7FFFFFFFFFFFFFFFull) : ~raw[4];
        memcpy(&out->return_max, &b, 8);
    }
    if (clear)
        Q1_CUDA(cudaMemset(env->P.metrics, 0, sizeof(raw)));
    return Q1_OK;
}

/* -- phys.apply ---------------------------------------------------------------------------------- */

int q1_phys_apply(int device, int64_t n, const double *yaw, const double *pitch, const double *roll,
                  const double *fmove, const double *smove, const uint8_t *button2,
                  const double *time_delta, int time_delta_f32, const double *z_pos, const float *vel,
                  const uint8_t *on_ground, const uint8_t *jump_released, double *z_pos_out,
                  float *vel_out, uint8_t *on_ground_out, uint8_t *jump_released_out, void *stream)
{
    if (n < 0)
        return fail(Q1_EINVAL, "n must be >= 0");
    if (n == 0)
        return Q1_OK;
    if (!yaw || !fmove || !smove || !button2 || !time_delta || !z_pos || !vel || !on_ground ||
        !jump_released || !z_pos_out || !vel_out || !on_ground_out || !jump_released_out)
        return fail(Q1_EINVAL, "a required array is NULL");
    DeviceGuard guard(device);
    if (!guard.ok)
        return fail(Q1_ENODEV, "cudaSetDevice failed: libq1phys has no CPU implementation");
    k_phys_apply<<<grid_for(n), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
        n, yaw, pitch, roll, fmove, smove, button2, time_delta, time_delta_f32, z_pos, vel, on_ground,
        jump_released, z_pos_out, vel_out, on_ground_out, jump_released_out);
    return check_launch("k_phys_apply");
}


int q1_phys_apply_host(int device, int64_t n, const double *yaw, const double *pitch,
                       const double *roll, const double *fmove, const double *smove,
                       const uint8_t *button2, const double *time_delta, int time_delta_f32,
                       const double *z_pos,
                       const float *vel, const uint8_t *on_ground, const uint8_t *jump_released,
                       double *z_pos_out, float *vel_out, uint8_t *on_ground_out,
                       uint8_t *jump_released_out)
{
    if (n < 0)
        return fail(Q1_EINVAL, "n must be >= 0");
    if (n == 0)
        return Q1_OK;
    if (!yaw || !fmove || !smove || !button2 || !time_delta || !z_pos || !vel || !on_ground ||
        !jump_released || !z_pos_out || !vel_out || !on_ground_out || !jump_released_out)
        return fail(Q1_EINVAL, "a required array is NULL");
    DeviceGuard guard(device);
    if (!guard.ok)
        return fail(Q1_ENODEV, "cudaSetDevice failed: libq1phys has no CPU implementation");
    const size_t N = (size_t)n;
    Staging st;
    int rc = st.reserve(align_up(8 * N) * 8 + align_up(12 * N) * 2 + align_up(N) * 5 + 4096);
    if (rc != Q1_OK)
        return rc;
    auto up = [&](auto *dst, const auto *src, size_t count) -> cudaError_t {
        return cudaMemcpy(dst, src, count * sizeof(*src), cudaMemcpyHostToDevice);
    };
    double *d_yaw = st.take<double>(N), *d_pitch = pitch ? st.take<double>(N) : nullptr;
    double *d_roll = roll ? st.take<double>(N) : nullptr;
    double *d_fm = st.take<double>(N), *d_sm = st.take<double>(N), *d_dt = st.take<double>(N);
    double *d_z = st.take<double>(N), *d_zo = st.take<double>(N);
    float *d_vel = st.take<float>(3 * N), *d_velo = st.take<float>(3 * N);
    uint8_t *d_b2 = st.take<uint8_t>(N), *d_og = st.take<uint8_t>(N), *d_jr = st.take<uint8_t>(N);
    uint8_t *d_ogo = st.take<uint8_t>(N), *d_jro = st.take<uint8_t>(N);
    Q1_CUDA(up(d_yaw, yaw, N));
    if (pitch)
        Q1_CUDA(up(d_pitch, pitch, N));
    if (roll)
        Q1_CUDA(up(d_roll, roll, N));
    Q1_CUDA(up(d_fm, fmove, N));
    Q1_CUDA(up(d_sm, smove, N));
    Q1_CUDA(up(d_dt, time_delta, N));
    Q1_CUDA(up(d_z, z_pos, N));
    Q1_CUDA(up(d_vel, vel, 3 * N));
    Q1_CUDA(up(d_b2, button2, N));
    Q1_CUDA(up(d_og, on_ground, N));
    Q1_CUDA(up(d_jr, jump_released, N));
    rc = q1_phys_apply(device, n, d_yaw, d_pitch, d_roll, d_fm, d_sm, d_b2, d_dt, time_delta_f32, d_z,
                       d_vel, d_og, d_jr, d_zo, d_velo, d_ogo, d_jro, nullptr);
    if (rc != Q1_OK)
        return rc;
    Q1_CUDA(cudaMemcpy(z_pos_out, d_zo, 8 * N, cudaMemcpyDeviceToHost));
    Q1_CUDA(cudaMemcpy(vel_out, d_velo, 12 * N, cudaMemcpyDeviceToHost));
    Q1_CUDA(cudaMemcpy(on_ground_out, d_ogo, N, cudaMemcpyDeviceToHost));
    Q1_CUDA(cudaMemcpy(jump_released_out, d_jro, N, cudaMemcpyDeviceToHost));
    return Q1_OK;
}

int q1_phys_apply_vel64_host(int device, int64_t n, const double *yaw, const double *pitch,
                             const double *roll, const double *fmove, const double *smove,
                             const uint8_t *button2, const double *time_delta, int time_delta_f32,
                             const double *z_pos, const double *vel, const uint8_t *on_ground,
                             const uint8_t *jump_released, double *z_pos_out, double *vel_out,
                             uint8_t *on_ground_out, uint8_t *jump_released_out)
{
    if (n < 0)
        return fail(Q1_EINVAL, "n must be >= 0");
    if (n == 0)
        return Q1_OK;
    if (!yaw || !fmove || !smove || !button2 || !time_delta || !z_pos || !vel || !on_ground ||
        !jump_released || !z_pos_out || !vel_out || !on_ground_out || !jump_released_out)
        return fail(Q1_EINVAL, "a required array is NULL");
    DeviceGuard guard(device);
    if (!guard.ok)
        return fail(Q1_ENODEV, "cudaSetDevice failed: libq1phys has no CPU implementation");
    const size_t N = (size_t)n;
    Staging st;
    int rc = st.reserve(align_up(8 * N) * 8 + align_up(24 * N) * 2 + align_up(N) * 5 + 4096);
    if (rc != Q1_OK)
        return rc;
    auto up = [&](auto *dst, const auto *src, size_t count) -> cudaError_t {
        return cudaMemcpy(dst, src, count * sizeof(*src), cudaMemcpyHostToDevice);
    };
    double *d_yaw = st.take<double>(N), *d_pitch = pitch ? st.take<double>(N) : nullptr;
    double *d_roll = roll ? st.take<double>(N) : nullptr;
    double *d_fm = st.take<double>(N), *d_sm = st.take<double>(N), *d_dt = st.take<double>(N);
    double *d_z = st.take<double>(N), *d_zo = st.take<double>(N);
    double *d_vel = st.take<double>(3 * N), *d_velo = st.take<double>(3 * N);
    uint8_t *d_b2 = st.take<uint8_t>(N), *d_og = st.take<uint8_t>(N), *d_jr = st.take<uint8_t>(N);
    uint8_t *d_ogo = st.take<uint8_t>(N), *d_jro = st.take<uint8_t>(N);
    Q1_CUDA(up(d_yaw, yaw, N));
    if (pitch)
        Q1_CUDA(up(d_pitch, pitch, N));
    if (roll)
        Q1_CUDA(up(d_roll, roll, N));
    Q1_CUDA(up(d_fm, fmove, N));
    Q1_CUDA(up(d_sm, smove, N));
    Q1_CUDA(up(d_dt, time_delta, N));
    Q1_CUDA(up(d_z, z_pos, N));
    Q1_CUDA(up(d_vel, vel, 3 * N));
    Q1_CUDA(up(d_b2, button2, N));
    Q1_CUDA(up(d_og, on_ground, N));
    Q1_CUDA(up(d_jr, jump_released, N));
    k_phys_apply_vel64<<<grid_for(n), kBlock>>>(n, d_yaw, d_pitch, d_roll, d_fm, d_sm, d_b2, d_dt,
                                                time_delta_f32, d_z, d_vel, d_og, d_jr, d_zo, d_velo, d_ogo,
                                                d_jro);
    rc = check_launch("k_phys_apply_vel64");
    if (rc != Q1_OK)
        return rc;
    Q1_CUDA(cudaMemcpy(z_pos_out, d_zo, 8 * N, cudaMemcpyDeviceToHost));
    Q1_CUDA(cudaMemcpy(vel_out, d_velo, 24 * N, cudaMemcpyDeviceToHost));
    Q1_CUDA(cudaMemcpy(on_ground_out, d_ogo, N, cudaMemcpyDeviceToHost));
    Q1_CUDA(cudaMemcpy(jump_released_out, d_jro, N, cudaMemcpyDeviceToHost));
    return Q1_OK;
}

int q1_delta_speed_sweep_host(int device, int64_t n, int64_t num_angles, const double *base_yaw,
                              const double *rel_angles, double fmove, double smove,
                              const uint8_t *button2, double time_delta, int time_delta_f32,
                              const double *z_pos, const float *vel, const uint8_t *on_ground,
                              const uint8_t *jump_released, float *delta_speed)
{
    if (n < 0 || num_angles < 0 || num_angles > 65535)
        return fail(Q1_EINVAL, "n must be >= 0 and num_angles in [0, 65535]");
    if (n == 0 || num_angles == 0)
        return Q1_OK;
    if (!base_yaw || !rel_angles || !button2 || !z_pos || !vel || !on_ground || !jump_released ||
        !delta_speed)
        return fail(Q1_EINVAL, "a required array is NULL");
    DeviceGuard guard(device);
    if (!guard.ok)
        return fail(Q1_ENODEV, "cudaSetDevice failed: libq1phys has no CPU implementation");
    const size_t N = (size_t)n, A = (size_t)num_angles;
    Staging st;
    int rc = st.reserve(align_up(8 * N) * 2 + align_up(8 * A) + align_up(12 * N) + align_up(N) * 3 +
                        align_up(4 * N * A) + 4096);
    if (rc != Q1_OK)
        return rc;
    double *d_yaw = st.take<double>(N), *d_rel = st.take<double>(A), *d_z = st.take<double>(N);
    float *d_vel = st.take<float>(3 * N), *d_out = st.take<float>(N * A);
    uint8_t *d_b2 = st.take<uint8_t>(N), *d_og = st.take<uint8_t>(N), *d_jr = st.take<uint8_t>(N);
    Q1_CUDA(cudaMemcpy(d_yaw, base_yaw, 8 * N, cudaMemcpyHostToDevice));
    Q1_CUDA(cudaMemcpy(d_rel, rel_angles, 8 * A, cudaMemcpyHostToDevice));
    Q1_CUDA(cudaMemcpy(d_z, z_pos, 8 * N, cudaMemcpyHostToDevice));
    Q1_CUDA(cudaMemcpy(d_vel, vel, 12 * N, cudaMemcpyHostToDevice));
    Q1_CUDA(cudaMemcpy(d_b2, button2, N, cudaMemcpyHostToDevice));
    Q1_CUDA(cudaMemcpy(d_og, on_ground, N, cudaMemcpyHostToDevice));
    Q1_CUDA(cudaMemcpy(d_jr, jump_released, N, cudaMemcpyHostToDevice));
    dim3 grid(grid_for(n), (unsigned)num_angles);
    k_delta_speed_sweep<<<grid, kBlock>>>(n, num_angles, d_yaw, d_rel, fmove, smove, d_b2, time_delta,
                                          time_delta_f32, d_z, d_vel, d_og, d_jr, d_out);
    rc = check_launch("k_delta_speed_sweep");
    if (rc != Q1_OK)
        return rc;
    Q1_CUDA(cudaMemcpy(delta_speed, d_out, 4 * N * A, cudaMemcpyDeviceToHost));
    return Q1_OK;
}

/* -- ActionDecoder.map ----------------------------------------------------------------------------- */

int q1_decode_host(const q1_config *cfg, int device, int64_t n, uint8_t *last_keys,
                   double *last_press, double *yaw, const uint8_t *keys, const double *mouse,
                   const float *z_vel, const double *time_remaining, int64_t *smove,
                   int64_t *fmove, uint8_t *jump)
{
    if (!cfg)
        return fail(Q1_EINVAL, "cfg is NULL");
    if (n < 0)
        return fail(Q1_EINVAL, "n must be >= 0");
    if (n == 0)
        return Q1_OK;
    if (!last_keys || !last_press || !yaw || !keys || !z_vel || !time_remaining || !smove ||
        !fmove || !jump)
        return fail(Q1_EINVAL, "a required array is NULL");
    if (cfg->allow_yaw && !mouse)
        return fail(Q1_EINVAL, "mouse is NULL but allow_yaw is set");
    q1_config c = *cfg;
    c.num_envs = n;
    Params P{};
    bool unused = false;
    int rc = derive_params(c, P, unused, (cfg->reserved & 1) != 0);
    if (rc != Q1_OK)
        return rc;
    DeviceGuard guard(device);
    if (!guard.ok)
        return fail(Q1_ENODEV, "cudaSetDevice failed: libq1phys has no CPU implementation");
    const size_t N = (size_t)n, nk = (size_t)P.num_keys;
    Staging st;
    rc = st.reserve(align_up(N * nk) * 2 + align_up(8 * N * nk) + align_up(8 * N) * 5 +
                    align_up(4 * N) + align_up(N) + 4096);
    if (rc != Q1_OK)
        return rc;
    uint8_t *d_lk = st.take<uint8_t>(N * nk), *d_keys = st.take<uint8_t>(N * nk);
    double *d_lp = st.take<double>(N * nk), *d_yaw = st.take<double>(N);
    double *d_mouse = st.take<double>(N), *d_tr = st.take<double>(N);
    int64_t *d_sm = st.take<int64_t>(N), *d_fm = st.take<int64_t>(N);
    float *d_zv = st.take<float>(N);
    uint8_t *d_jump = st.take<uint8_t>(N);
    Q1_CUDA(cudaMemcpy(d_lk, last_keys, N * nk, cudaMemcpyHostToDevice));
    Q1_CUDA(cudaMemcpy(d_keys, keys, N * nk, cudaMemcpyHostToDevice));
    Q1_CUDA(cudaMemcpy(d_lp, last_press, 8 * N * nk, cudaMemcpyHostToDevice));
    Q1_CUDA(cudaMemcpy(d_yaw, yaw, 8 * N, cudaMemcpyHostToDevice));
    if (cfg->allow_yaw)
        Q1_CUDA(cudaMemcpy(d_mouse, mouse, 8 * N, cudaMemcpyHostToDevice));
    Q1_CUDA(cudaMemcpy(d_tr, time_remaining, 8 * N, cudaMemcpyHostToDevice));
    Q1_CUDA(cudaMemcpy(d_zv, z_vel, 4 * N, cudaMemcpyHostToDevice));
    k_decode<<<grid_for(n), kBlock>>>(P, d_lk, d_lp, d_yaw, d_keys, d_mouse, d_zv, d_tr, d_sm, d_fm,
                                      d_jump);
    rc = check_launch("k_decode");
    if (rc != Q1_OK)
        return rc;
    Q1_CUDA(cudaMemcpy(last_keys, d_lk, N * nk, cudaMemcpyDeviceToHost));
    Q1_CUDA(cudaMemcpy(last_press, d_lp, 8 * N * nk, cudaMemcpyDeviceToHost));
    Q1_CUDA(cudaMemcpy(yaw, d_yaw, 8 * N, cudaMemcpyDeviceToHost));
    Q1_CUDA(cudaMemcpy(smove, d_sm, 8 * N, cudaMemcpyDeviceToHost));
    Q1_CUDA(cudaMemcpy(fmove, d_fm, 8 * N, cudaMemcpyDeviceToHost));
    Q1_CUDA(cudaMemcpy(jump, d_jump, N, cudaMemcpyDeviceToHost));
    return Q1_OK;
}

int q1_selftest_division(int device, uint64_t samples, uint64_t seed, uint64_t mismatches[8])
{
    if (!mismatches)
        return fail(Q1_EINVAL, "mismatches is NULL");
    DeviceGuard guard(device);
    if (!guard.ok)
        return fail(Q1_ENODEV, "cudaSetDevice failed: libq1phys has no CPU implementation");
    unsigned long long *d_out = nullptr;
    Q1_CUDA(cudaMalloc(&d_out, 8 * sizeof(unsigned long long)));
    Q1_CUDA(cudaMemset(d_out, 0, 8 * sizeof(unsigned long long)));
    const unsigned blocks = 148 * 8;
    uint64_t iters = (samples + (uint64_t)blocks * 256 - 1) / ((uint64_t)blocks * 256);
    k_selftest<<<blocks, 256>>>(iters, seed, d_out);
    int rc = check_launch("k_selftest");
    if (rc == Q1_OK) {
        cudaError_t err = cudaMemcpy(mismatches, d_out, 8 * sizeof(unsigned long long),
                                     cudaMemcpyDeviceToHost);
        if (err != cudaSuccess)
            rc = fail(Q1_ECUDA, std::string("k_selftest: ") + cudaGetErrorString(err));
    }
    cudaFree(d_out);
    return rc;
}

int q1_sincos_host(int device, int64_t n, const double *x, double *sin_out, double *cos_out)
{
    if (n < 0)
        return fail(Q1_EINVAL, "n must be >= 0");
    if (n == 0)
        return Q1_OK;
    if (!x || !sin_out || !cos_out)
        return fail(Q1_EINVAL, "x / sin_out / cos_out is NULL");
    DeviceGuard guard(device);
    if (!guard.ok)
        return fail(Q1_ENODEV, "cudaSetDevice failed: libq1phys has no CPU implementation");
    double *d = nullptr;
    Q1_CUDA(cudaMalloc(&d, 3 * (size_t)n * sizeof(double)));
    int rc = Q1_OK;
    cudaError_t err = cudaMemcpy(d, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
    if (err == cudaSuccess) {
        k_sincos<<<(unsigned)((n + 255) / 256), 256>>>(n, d, d + n, d + 2 * n);
        rc = check_launch("k_sincos");
    }
    if (err == cudaSuccess && rc == Q1_OK)
        err = cudaMemcpy(sin_out, d + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
    if (err == cudaSuccess && rc == Q1_OK)
        err = cudaMemcpy(cos_out, d + 2 * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (err != cudaSuccess)
        return fail(Q1_ECUDA, std::string("q1_sincos_host: ") + cudaGetErrorString(err));
    return rc;
}

int q1_sample_actions(int device, int64_t n, int num_keys, const float *logits, double action_low,
                      double action_high, int deterministic, uint64_t seed, uint64_t step,
                      const uint64_t *step_device, uint64_t env_index_base, uint8_t *keys,
                      float *mouse, void *stream)
{
    if (n < 0 || (num_keys != 3 && num_keys != 4))
        return fail(Q1_EINVAL, "n must be >= 0 and num_keys 3 or 4");
    if (n == 0)
        return Q1_OK;
    if (!logits || !keys || !mouse)
        return fail(Q1_EINVAL, "logits / keys / mouse is NULL");
    DeviceGuard guard(device);
    if (!guard.ok)
        return fail(Q1_ENODEV, "cudaSetDevice failed: libq1phys has no CPU implementation");
    k_sample_actions<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n, num_keys, logits, (float)action_low, (float)action_high, deterministic, seed, step,
        step_device, env_index_base, keys, mouse);
    return check_launch("k_sample_actions");
}

} /* extern "C" */
