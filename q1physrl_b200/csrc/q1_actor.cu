/*
 * q1_actor.cu -- the shipped PPO policy and the env it drives as ONE warp-specialised sm_100a kernel
 * (BASELINE config 5):
 *
 *   k_actor<LOOP = false>   obs -> tanh 256 -> tanh 256 -> logits -> Q1PhysActionDist sample -> the action
 *                           arrays q1_step consumes (q1_policy_act), tile after tile of 128 envs;
 *   k_actor<LOOP = true>    the closed loop: every CTA keeps the state of up to three tiles of 128 envs in
 *                           shared memory for T ticks and runs policy -> sample -> tick<>() -> observe ->
 *                           policy ... with no launch and no HBM traffic per tick (q1_policy_rollout);
 *                           optionally writes the per-tick record of analyse.eval_sim.
 * Reference: q1physrl checkpoints (RLLib fcnet default_policy/fc_1, fc_2, fc_out),
 * q1physrl/action_dist.py:84-101, 186-243, q1physrl_env/env.py:482-510.
 *
 * All three layers run on the 5th-generation tensor cores (tcgen05.mma, bf16 operands, fp32 accumulators
 * in tensor memory); activations never leave tensor memory.  28 warps in seven warpgroups, three roles,
 * synchronised by mbarriers only:
 *
 *   warps 0-7    ENV   two groups of four; a thread per env row: builds the layer-1 operand from the
 *                      observation, reads the logits, samples the action and (LOOP) runs the env tick for
 *                      its env.  The groups take alternate sequences (see the role's comment)
 *   warps 8-23   EPI   tanh epilogues: tcgen05.ld accumulator columns -> (+ bias) -> tanh -> bf16 ->
 *                      tcgen05.st as the next layer's A operand; thread = (row, column part)
 *   warp 24      MMA   issues every tcgen05.mma and tcgen05.commit (warps 25-27 only complete its
 *                      warpgroup: registers move between whole warpgroups, setmaxnreg)
 *
 * The tanh epilogues, not the MMAs, bound the policy (65 536 tanh per tile: 4096 cycles of the XU if all
 * went through MUFU.TANH; the MMAs of a tile are ~2600 cycles), and every publish of epilogue output to the
 * MMA warp (tcgen05.wait::st, fence, mbarrier arrive, the MMA warp's wake-up) costs a few hundred cycles
 * whatever it publishes.  So: few, large steps, and the XU never waits for the tensor pipe:
 *   layer 1   two N = 128 halves, both K-steps each.  Its epilogue runs in two halves of 128 columns, all
 *             four parts at once (part p: 32 columns of each half)
 *   layer 2   four N = 64 quarters.  The first K-half of all four is issued when the first half of the
 *             layer-1 activations is published and runs under the second half of that epilogue; the second
 *             K-halves follow quarter by quarter, and epilogue part q starts on quarter q as it completes
 *   layer 3   K-steps trail the layer-2 epilogue part by part; the logits land over quarter 3's first
 *             accumulator columns
 *   next tile (open policy step) its layer 1 is issued as soon as parts 0 and 1 have their accumulators in
 *             registers, into the regions this tile no longer needs, so the epilogue warps go from this
 *             tile's layer 2 straight into the next tile's layer 1
 * With two or more tiles per CTA the env tick of one tile runs under the policy phase of the next.
 *
 * Layer 1 on the tensor cores without bf16-quantising the observation (yaw / 90 would lose 3 degrees):
 * x = x_hi + x_mid + x_lo (three bf16, 24 bits), w = w_hi + w_lo; the five products that matter
 * (hi.hi, hi.lo, mid.hi, mid.lo, lo.hi) of the 6 inputs are 30 columns of one K = 32 operand, the last
 * two carry 1 x (b_hi, b_lo): the bias comes out of the MMA too.  Relative error ~2^-17 of |x w|.  That
 * operand is written by the env warps into shared memory (K-major, 128-byte swizzle, like the weights).
 *
 * Tensor memory (512 columns = four regions of 128), for an even sequence; an odd one uses the regions
 * rotated (A <-> C, B <-> D: tm_a .. tm_d below), which is what lets two consecutive tiles overlap:
 *   A [128,256)  layer-1 accumulator, columns 0..127; then layer-2 accumulators, quarters 2 and 3; then the
 *                logits over quarter 3's first 16 columns.  The NEXT tile's H1
 *   B [256,384)  layer-1 accumulator, columns 128..255; then H2: layer-2 activations, A operand of layer 3.
 *                The next tile's quarters 0 and 1
 *   C [0,128)    H1: layer-1 activations, A operand of layer 2.  The next tile's region A
 *   D [384,512)  layer-2 accumulators, quarters 0 and 1.  The next tile's region B
 * A thread only ever touches its own lane (= env row), so within a lane program order is enough; the
 * barriers order lanes against the MMAs, and the MMAs of one issuing thread execute in issue order.
 */
#include "q1_internal.h"
#include "q1_device_common.cuh"
#include "q1_sample.cuh"

#include <cuda_bf16.h>

#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace q1;

namespace {

constexpr int kRows = 128;   /* envs per tile = UMMA M */
constexpr int kHidden = 256; /* hidden width = K of layers 2 and 3, N of layers 1 and 2 */
constexpr int kOutPad = 16;  /* layer-3 N, zero-padded from 2 * num_keys + 2 = 8 or 10 */
constexpr int kObs = 6;
constexpr int kEpiParts = 4;  /* epilogue warps = 4 lane quadrants x 4 column parts (one warp of each part per scheduler) */
constexpr int kEnvWarps = 8; /* two groups of four (a group = the 128 rows of a tile), taking alternate sequences */
constexpr int kEpiWarps = 4 * kEpiParts, kMmaWarp = kEnvWarps + kEpiWarps;
constexpr int kPartThreads = 128; /* threads of one epilogue part */
constexpr int kThreads = 32 * (kMmaWarp + 4); /* the MMA warp brings its warpgroup: registers move between whole warpgroups */
constexpr int kMaxTiles = 3; /* tiles of env state a CTA keeps in shared memory (LOOP) */

/* shared-memory image; every UMMA operand block is 1024-byte aligned (128-byte swizzle atoms) */
constexpr uint32_t SM_B2 = 0;                                  /* W2^T: 4 K-atoms x 256 rows x 128 B */
constexpr uint32_t SM_B3 = SM_B2 + 4 * kHidden * 128;          /* W3^T: 4 K-atoms x 16 rows x 128 B */
constexpr uint32_t SM_B1 = SM_B3 + 4 * kOutPad * 128;          /* layer-1 operand: 256 rows x 128 B (K = 32 used) */
constexpr uint32_t SM_BIAS2 = SM_B1 + kHidden * 128;           /* fp32 [256] */
constexpr uint32_t SM_BIAS3 = SM_BIAS2 + kHidden * 4;          /* fp32 [16] */
constexpr uint32_t SM_WEIGHTS_END = SM_BIAS3 + kOutPad * 4;
constexpr uint32_t kImageBytes = SM_WEIGHTS_END - SM_B2;        /* what the host image holds */
static_assert(kImageBytes % 16 == 0, "bulk copies move multiples of 16 bytes");
enum : uint32_t { /* mbarriers, 8 bytes each.  A waiter tests a phase PARITY, so no barrier may complete two
                     phases between two waits of the same waiter: every barrier here completes once per tile
                     (B_X, B_D3: once per two) and the protocol keeps every waiter within one tile of every signaller. */
    B_W = 0,        /* weights have landed in shared memory */
    B_X = 1,        /* [2] layer-1 operand of the tile written (env rows -> MMA) */
    B_L1 = 3,       /* layer-1 accumulator complete (MMA -> epilogue) */
    B_H1 = 4,       /* [2] step h of the layer-1 activations stored by all four parts (epilogue -> MMA) */
    B_L2 = 6,       /* [4] layer-2 accumulator, quarter q, complete */
    B_H2 = 10,      /* [4] quarter q of the layer-2 activations stored */
    B_D3 = 14,      /* [2] logits of an even / odd sequence complete (MMA -> that sequence's env rows; one barrier
                       per parity because an env group that takes every other sequence waits on every other phase) */
    B_E = 16,       /* logits read (env rows -> epilogue: the next tile's activations may overwrite them) */
    B_R01 = 17,     /* layer-2 accumulators, quarters 0 and 1, are in registers (epilogue parts 0, 1 -> MMA:
                       region D may take the next tile's layer 1) */
    B_R3 = 18,      /* quarter 3's first columns are in registers (part 3 -> MMA: the logits may go there) */
    B_COUNT = 19
};
constexpr uint32_t SM_X = (SM_WEIGHTS_END + 1023) & ~1023u;     /* two layer-1 operands: 128 rows x 128 B (K = 32 used) */
constexpr uint32_t SM_BAR = SM_X + 2 * kRows * 128;
constexpr uint32_t SM_TMEM = SM_BAR + 8 * B_COUNT;
constexpr uint32_t SM_STATE = (SM_TMEM + 16 + 127) & ~127u;    /* LOOP: kMaxTiles x kSlotBytes */
constexpr uint32_t SLOT_EPOCH = kTileBytes;                     /* u32 [128] */
constexpr uint32_t SLOT_RETURN = SLOT_EPOCH + 4 * kRows;        /* f64 [128] */
constexpr uint32_t kSlotBytes = SLOT_RETURN + 8 * kRows;
constexpr uint32_t SM_TOTAL_ACT = SM_STATE;
constexpr uint32_t SM_TOTAL_LOOP = SM_STATE + kMaxTiles * kSlotBytes;
static_assert(SM_TOTAL_LOOP <= 232448, "one CTA per SM: at most 227 KB of shared memory");

/* tensor-memory columns (32-bit) of the tile of parity a = sequence & 1, see the map above */
__host__ __device__ constexpr uint32_t tm_a(uint32_t a) { return a ? 0u : 128u; }   /* L1 cols 0..127; then L2 quarters 2, 3 */
__host__ __device__ constexpr uint32_t tm_b(uint32_t a) { return a ? 384u : 256u; } /* L1 cols 128..255; then H2 */
__host__ __device__ constexpr uint32_t tm_c(uint32_t a) { return a ? 128u : 0u; }   /* H1 */
__host__ __device__ constexpr uint32_t tm_d(uint32_t a) { return a ? 256u : 384u; } /* L2 quarters 0, 1 */
__host__ __device__ constexpr uint32_t tm_l2(uint32_t a, uint32_t q) { return q < 2 ? tm_d(a) + 64u * q : tm_a(a) + 64u * (q - 2u); }
__host__ __device__ constexpr uint32_t tm_d3(uint32_t a) { return tm_a(a) + 64u; }  /* logits: over quarter 3's first columns */
static_assert(tm_a(1) == tm_c(0) && tm_b(1) == tm_d(0) && tm_c(1) == tm_a(0) && tm_d(1) == tm_b(0),
              "the next tile's regions are this tile's, rotated");

/* instruction descriptor of tcgen05.mma kind::f16: D = f32, A = B = bf16, both K-major, M = 128 */
__host__ __device__ constexpr uint32_t instr_desc(uint32_t n)
{
    return (1u << 4) /* c_format f32 */ | (1u << 7) /* a bf16 */ | (1u << 10) /* b bf16 */ |
           ((n >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
}

/* shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row atoms 1024 B apart */
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);       /* start address */
    d |= (uint64_t)1u << 16;                       /* leading byte offset: unused when swizzled */
    d |= (uint64_t)(1024u >> 4) << 32;             /* stride byte offset between 8-row atoms */
    d |= (uint64_t)1u << 46;                       /* descriptor version (Blackwell) */
    d |= (uint64_t)2u << 61;                       /* SWIZZLE_128B */
    return d;
}

/* the descriptor of `bytes` further on (bytes a multiple of 16, no carry out of the 14-bit address field:
 * shared memory is 228 KB, the field covers 256 KB) */
__device__ __forceinline__ uint64_t desc_at(uint64_t base, uint32_t bytes) { return base + (uint64_t)(bytes >> 4); }

__device__ __forceinline__ uint32_t saddr_of(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* A operand from tensor memory (lane = row, 16-bit elements packed two per column along K) */
/* The MMA warp runs its whole program on all 32 lanes, so that addresses and descriptors are warp-uniform
 * values the compiler can keep in the uniform datapath; only the tcgen05 instructions themselves are
 * predicated on the elected lane (`leader`).  Issued from a divergent `if (lane == 0)` region instead, every
 * MMA cost ~80 cycles of descriptor arithmetic and R2UR moves -- more than an N <= 128 MMA takes to execute. */
__device__ __forceinline__ void mma_bf16_ts(uint32_t leader, uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                            uint32_t idesc, bool accumulate)
{
    asm volatile("{\n\t"
                 ".reg .pred p, q;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "setp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
                 "}\n" ::"r"(tmem_d),
                 "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate), "r"(leader)
                 : "memory");
}
/* both operands from shared memory */
__device__ __forceinline__ void mma_bf16_ss(uint32_t leader, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, bool accumulate)
{
    asm volatile("{\n\t"
                 ".reg .pred p, q;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "setp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                 "}\n" ::"r"(tmem_d),
                 "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate), "r"(leader)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t leader, uint32_t bar)
{
    asm volatile("{\n\t"
                 ".reg .pred q;\n\t"
                 "setp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
                 "}\n" ::"r"(bar), "r"(leader)
                 : "memory");
}
/* 1 on exactly one lane of the (converged) warp, the same lane every time */
__device__ __forceinline__ uint32_t elect_leader()
{
    uint32_t is;
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "elect.sync _|p, 0xffffffff;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t"
                 "}\n" : "=r"(is));
    return is;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
/* Watchdog of the role handshakes: a wait that has polled kWatchdogPolls times (seconds; a healthy wait
 * takes microseconds) records who waited for what and raises g_fault[0]; every wait in every CTA then
 * falls through, so a protocol bug ends the launch with an error report (q1_policy_* return Q1_ECUDA
 * at their next call) instead of hanging the GPU. */
constexpr uint32_t kWatchdogPolls = 1u << 26;
__device__ unsigned int g_fault[8]; /* [0] raised, [1] tag of the first waiter that gave up, [2] sequence */

__device__ __forceinline__ bool bar_test(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, p;\n"
                 "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
/* tag = role (1 MMA, 2 EPI, 3 ENV) << 12 | barrier index << 7 | sub-step << 4.
 * Everything here is INLINE on purpose: a wait may sit between a tcgen05.ld and its wait::ld, and a
 * function call there lets the callee (or the caller's spill code around the call) use the very
 * registers the asynchronous load is still going to write. */
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity, uint32_t tag = 0, int64_t seq = 0)
{
    for (uint32_t polls = 0; !bar_test(bar, parity); polls++) {
        if ((polls & 1023u) == 1023u && *reinterpret_cast<volatile unsigned int *>(&g_fault[0])) {
            /* where the other roles stood when the first waiter gave up (first of each role) */
            atomicCAS(&g_fault[4 + ((tag >> 12) & 3u)], 0u, (tag & 0xFFFu) | ((uint32_t)seq << 12) | 0x80000000u);
            break;
        }
        if (polls >= kWatchdogPolls) {
            if (atomicExch(&g_fault[0], 1u) == 0u) {
                g_fault[1] = tag;
                g_fault[2] = (uint32_t)seq;
                g_fault[3] = blockIdx.x;
            }
            break;
        }
    }
}
/* The lanes of a warp leave a wait at different times; the whole-warp roles re-converge explicitly
 * because what follows a wait is a .sync.aligned tcgen05 instruction, which every lane must issue
 * together. */
__device__ __forceinline__ void bar_wait_warp(uint32_t bar, uint32_t parity, uint32_t tag = 0, int64_t seq = 0)
{
    bar_wait(bar, parity, tag, seq);
    __syncwarp();
}
__device__ __forceinline__ __attribute__((unused)) uint32_t pack_bf16(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
/* tanh of two pre-activations -> packed bf16x2 (what the next layer's A operand holds).  The special-
 * function unit is what bounds the epilogues: MUFU.TANH retires 4 lanes per clock per scheduler (a warp
 * instruction every 8 cycles, measured), and a tile needs 65 536 of them = 4096 cycles per SM against
 * ~2300 for its MMAs.
 *   Q1_POLICY_TANH 0 (default)  tanh.approx.f32 per element (relative error 2^-11), two MUFU per pair
 *   Q1_POLICY_TANH 1            tanh.approx.f16x2: ptxas splits it into two MUFU.TANH.F16 plus conversions
 *                               (no faster, measured)
 *   Q1_POLICY_TANH 2            tanh.approx.bf16x2: inputs rounded to bf16 first, ~3x the logit error
 * So the XU is relieved the other way round: Q1_POLICY_POLY_PAIRS of every 16 column pairs do not go to it
 * at all but through tanh2_poly_bf16 below on the FMA pipe, which the epilogue otherwise leaves idle. */
#ifndef Q1_POLICY_TANH
#define Q1_POLICY_TANH 0
#endif
/* Measured on one box, per 2^20 envs of k_actor<ACT> / per tick of the closed loop at 32 768 envs:
 *   pairs    0      2      3      4
 *   ACT    195.4  182.3  184.2  182.3 us
 *   LOOP    8.60   8.52   8.57   8.70 us     (there the chain env tick -> policy -> env tick of a tile, not
 * a pipe, sets the pace, and the env rows want the FMA pipe too).  ONE value for both, because the closed loop
 * is tested bit-identical to per-tick q1_policy_act + q1_step: 2. */
#ifndef Q1_POLICY_POLY_PAIRS
#define Q1_POLICY_POLY_PAIRS 2
#endif

/* tanh(x) ~ x P(x^2) on |x| <= 3.5, x clamped to that range first (beyond it tanh rounds to +-1 in bf16, and
 * so does this).  tools/make_tanh_poly.py: degree 17, minimax relative error 5.4e-4 (5.8e-4 evaluated in
 * float32) -- the size of MUFU.TANH's own 2^-11 and a quarter of the bf16 rounding of the stored activation;
 * 96 % of the results equal bf16(tanh x), the rest are the neighbouring bf16.  Evaluated on both elements
 * of the pair at once with the packed float32 instructions of sm_100 (FMUL2 / FFMA2, coefficients as
 * immediates): 4 FMNMX + 10 packed instructions per pair.  A NaN input clamps to +-3.5 (MUFU would return
 * NaN); q1_policy_check rejects non-finite weights and the observations of a finite state are finite. */
constexpr float kTanhClamp = 3.5f;
#define Q1_TANH_POLY(X) /* constant term first */                                                            \
    X(0.9994622468948364f) X(-0.32549557089805603f) X(0.1126757487654686f) X(-0.0303287822753191f)            \
    X(0.005662580020725727f) X(-0.0006888179923407733f) X(5.15220635861624e-05f) X(-2.1402802303782664e-06f)  \
    X(3.7685438769585744e-08f)
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint32_t tanh2_poly_bf16(float lo, float hi)
{
    lo = fminf(fmaxf(lo, -kTanhClamp), kTanhClamp);
    hi = fminf(fmaxf(hi, -kTanhClamp), kTanhClamp);
    const uint64_t x = pack_f32x2(lo, hi), t = mul_f32x2(x, x);
#define X(c) c,
    const float coef[9] = {Q1_TANH_POLY(X)};
#undef X
    uint64_t p = pack_f32x2(coef[8], coef[8]);
#pragma unroll
    for (int j = 7; j >= 0; j--)
        p = fma_f32x2(p, t, pack_f32x2(coef[j], coef[j]));
    p = mul_f32x2(p, x);
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p));
    return pack_bf16(a, b);
}
/* column pair e of a 32-column step: on the FMA pipe or the XU?  (spread evenly over the 16 pairs) */
template <uint32_t PAIRS>
__host__ __device__ constexpr bool pair_uses_poly(uint32_t e) { return (e * PAIRS) % 16u < PAIRS; }
__device__ __forceinline__ uint32_t tanh2_bf16(float lo, float hi)
{
#if Q1_POLICY_TANH == 1
    uint32_t h2, t2, out;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(hi), "f"(lo));
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t2) : "r"(h2));
    asm("{\n"
        ".reg .b16 l, h;\n"
        ".reg .f32 fl, fh;\n"
        "mov.b32 {l, h}, %1;\n"
        "cvt.f32.f16 fl, l;\n"
        "cvt.f32.f16 fh, h;\n"
        "cvt.rn.bf16x2.f32 %0, fh, fl;\n"
        "}" : "=r"(out) : "r"(t2));
    return out;
#elif Q1_POLICY_TANH == 2
    uint32_t x = pack_bf16(lo, hi), y;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
#else
    float a, b;
    asm("tanh.approx.f32 %0, %1;" : "=f"(a) : "f"(lo));
    asm("tanh.approx.f32 %0, %1;" : "=f"(b) : "f"(hi));
    return pack_bf16(a, b);
#endif
}
template <uint32_t PAIRS>
__device__ __forceinline__ uint32_t tanh2_pair(uint32_t e, float lo, float hi)
{
    return pair_uses_poly<PAIRS>(e) ? tanh2_poly_bf16(lo, hi) : tanh2_bf16(lo, hi);
}
/* consecutive 32-bit columns of this thread's TMEM lane */
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t v[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t v[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
/* tcgen05.ld is asynchronous: the *_issue forms only start it, tmem_ld_wait*() completes it.  The wait
 * also "redefines" the destination registers for the compiler (empty asm with them as in/out operands),
 * so that no use of them can be scheduled above it. */
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t v[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                   "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                   "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
                   "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t v[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                   "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                   "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void pin16(uint32_t *v)
{
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                      "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]),
                      "+r"(v[14]), "+r"(v[15]));
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t v[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t v[4])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3])
                 : "memory");
}
/* N consecutive columns, N in {8, 16, 32} (loads) / {4, 8, 16} (stores) */
template <int N> __device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t *v)
{
    if (N == 8)
        tmem_ld8_issue(taddr, v);
    else if (N == 16)
        tmem_ld16_issue(taddr, v);
    else
        tmem_ld32_issue(taddr, v);
}
template <int N> __device__ __forceinline__ void tmem_ld_wait(uint32_t *v)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (N == 8) {
        asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]));
    } else {
        pin16(v);
        if (N == 32)
            pin16(v + 16);
    }
}
template <int N> __device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t *v)
{
    if (N == 4)
        tmem_st4(taddr, v);
    else if (N == 8)
        tmem_st8(taddr, v);
    else
        tmem_st16(taddr, v);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32"
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                   "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                   "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

/* x = hi + mid + lo in three bf16 (24 significant bits): the bit patterns */
__device__ __forceinline__ void split3(float x, uint32_t &hi, uint32_t &mid, uint32_t &lo)
{
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);          /* exact */
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);         /* exact */
    const __nv_bfloat16 l = __float2bfloat16_rn(r2);
    hi = __bfloat16_as_ushort(h);
    mid = __bfloat16_as_ushort(m);
    lo = __bfloat16_as_ushort(l);
}

/* The layer-1 A operand of one env row: K = 32 bf16 in 16 packed columns.  Input k occupies K indices
 * 5k .. 5k+4 = (hi, hi, mid, mid, lo), matching the weight rows (w_hi, w_lo, w_hi, w_lo, w_hi) the host
 * lays out; K = 30, 31 are 1.0 against (bias_hi, bias_lo). */
__device__ __forceinline__ void layer1_operand(const float o[kObs], uint32_t cols[16])
{
    uint32_t a[32];
#pragma unroll
    for (int k = 0; k < kObs; k++) {
        uint32_t h, m, l;
        split3(o[k], h, m, l);
        a[5 * k] = h;
        a[5 * k + 1] = h;
        a[5 * k + 2] = m;
        a[5 * k + 3] = m;
        a[5 * k + 4] = l;
    }
    a[30] = a[31] = 0x3F80u; /* bf16 1.0 */
#pragma unroll
    for (int j = 0; j < 16; j++)
        cols[j] = a[2 * j] | (a[2 * j + 1] << 16);
}

/* -DQ1_ACTOR_TRACE=1: CTA 0 stamps the SM clock at its protocol events into g_trace[sequence][event]
 * (tools/trace_actor.py prints the timeline).  Never set in the shipped build. */
#ifndef Q1_ACTOR_TRACE
#define Q1_ACTOR_TRACE 0
#endif
#if Q1_ACTOR_TRACE
#ifndef Q1_ACTOR_TRACE_BASE
#define Q1_ACTOR_TRACE_BASE 0 /* first of the 16 sequences recorded */
#endif
#ifndef Q1_ACTOR_TRACE_BLOCK
#define Q1_ACTOR_TRACE_BLOCK 0
#endif
__device__ long long g_trace[16][64];
__device__ long long g_block_cycles[256];   /* (entry -> exit cycles) << 8 | SM id, per CTA */
#if Q1_ACTOR_TRACE == 2 /* only the entry / exit stamps: the code between them is the shipped build's */
#define TRACE(seq, ev) do { } while (0)
#else
#define TRACE(seq, ev)                                                                                     \
    do {                                                                                                   \
        if (blockIdx.x == Q1_ACTOR_TRACE_BLOCK && (seq) >= Q1_ACTOR_TRACE_BASE &&                         \
            (seq) < Q1_ACTOR_TRACE_BASE + 16 && (threadIdx.x & 31u) == 0)                                 \
            g_trace[(seq) - Q1_ACTOR_TRACE_BASE][(ev)] = clock64();                                       \
    } while (0)
#endif
#else
#define TRACE(seq, ev) do { } while (0)
#endif

struct ActorArgs {
    const unsigned char *image;
    int64_t n;                 /* envs (ACT: rows of obs; LOOP: the handle's num_envs) */
    int num_keys;
    float low, high;
    int deterministic;
    uint64_t seed;             /* of the sampling noise */
    uint64_t step;             /* position of the noise stream (ACT: this call; LOOP: first tick) */
    const uint64_t *step_device;
    uint64_t env_index_base;
    /* ACT */
    const float *obs;
    uint8_t *keys;
    float *mouse;
    float *logits_out;
    /* LOOP */
    int ticks;
    int auto_reset;
    int64_t tile_begin, tile_end; /* tiles [begin, end) of the handle go to this launch */
    uint32_t record_flags;
    q1_record_view rec;
    int64_t rec_tick0;         /* record row of this launch's first tick */
    float *final_obs;
    float *reward_sum;
};

/* one env out of / into a shared-memory state block */
__device__ __forceinline__ void slot_load(const unsigned char *blk, int l, Env &e)
{
    const float4 a = reinterpret_cast<const float4 *>(blk + kTileRecA)[l];
    const double2 b = reinterpret_cast<const double2 *>(blk + kTileRecB)[l];
    e.vx = a.x;
    e.vy = a.y;
    e.vz = a.z;
    e.bits = __float_as_uint(a.w);
    e.z = b.x;
    e.yaw = b.y;
    e.trem = reinterpret_cast<const double *>(blk + kTileTrem)[l];
}
__device__ __forceinline__ void slot_store(unsigned char *blk, int l, const Env &e)
{
    reinterpret_cast<float4 *>(blk + kTileRecA)[l] = make_float4(e.vx, e.vy, e.vz, __uint_as_float(e.bits));
    reinterpret_cast<double2 *>(blk + kTileRecB)[l] = make_double2(e.z, e.yaw);
    reinterpret_cast<double *>(blk + kTileTrem)[l] = e.trem;
}

/* 28 warps (the SM hands registers out per four warps) start with 72 registers each = 64 512 of the 65 536.
 * The MMA warpgroup -- one working warp -- gives all but 24 back, which is exactly what lets the six env and
 * epilogue warpgroups grow to the 80 they need (32 + 32 accumulator columns in flight + 16 packed results). */
template <uint32_t N>
__device__ __forceinline__ void warpgroup_reg_inc()
{
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <uint32_t N>
__device__ __forceinline__ void warpgroup_reg_dec()
{
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <bool LOOP, bool TRACK, bool LEAN, bool RECORD>
__global__ void __launch_bounds__(kThreads, 1)
k_actor(const __grid_constant__ Params P, const __grid_constant__ ActorArgs A)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t s0 = saddr_of(smem);
    auto bar = [&](uint32_t b) { return s0 + SM_BAR + 8u * b; };
#if Q1_ACTOR_TRACE
    const long long trace_entry = clock64();
    if (blockIdx.x == Q1_ACTOR_TRACE_BLOCK && tid == 0) {
        g_trace[0][62] = trace_entry;    /* kernel entry (event 63: exit) */
        unsigned long long ns;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
        g_trace[0][60] = (long long)ns;  /* the same two moments in nanoseconds: 60, 61 */
    }
#endif

    /* this CTA's work: LOOP: tiles tile_begin + blockIdx.x + j * gridDim.x (j < k), each for `ticks`
     * ticks, visited round-robin; ACT: the same tile walk, each tile once */
    const int64_t all_tiles = LOOP ? A.tile_end - A.tile_begin : (A.n + kRows - 1) / kRows;
    const int64_t first_tile = (LOOP ? A.tile_begin : 0) + blockIdx.x;
    const int64_t mine = ((LOOP ? A.tile_begin : 0) + all_tiles - first_tile + gridDim.x - 1) / gridDim.x;
    const int k = LOOP ? (int)mine : 1;                         /* tile slots of this CTA */
    const int64_t S = LOOP ? (int64_t)k * A.ticks : mine;      /* policy evaluations ("sequences") */
    auto tile_of = [&](int64_t s) { return first_tile + (LOOP ? (s % k) : s) * (int64_t)gridDim.x; };

    if (tid == 0) {
        bar_init(bar(B_W), 1);
        bar_init(bar(B_X + 0), kRows);
        bar_init(bar(B_X + 1), kRows);
        bar_init(bar(B_L1), 1);
        bar_init(bar(B_D3 + 0), 1);
        bar_init(bar(B_D3 + 1), 1);
        bar_init(bar(B_E), kRows);
        bar_init(bar(B_R01), 2 * kPartThreads);
        bar_init(bar(B_R3), kPartThreads);
        for (int b = 0; b < 2; b++)
            bar_init(bar(B_H1 + b), 4 * kPartThreads);
        for (int b = 0; b < 4; b++) {
            bar_init(bar(B_H2 + b), kPartThreads);
            bar_init(bar(B_L2 + b), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        /* the weight image -> shared memory, in 32 KB bulk copies */
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar(B_W)), "r"(kImageBytes)
                     : "memory");
        for (uint32_t off = 0; off < kImageBytes; off += 32768u) {
            const uint32_t len = kImageBytes - off < 32768u ? kImageBytes - off : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s0 + SM_B2 + off), "l"(A.image + off), "r"(len), "r"(bar(B_W))
                         : "memory");
        }
    }
    if (warp == 0) { /* one warp owns the TMEM allocation: all 512 columns */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s0 + SM_TMEM), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (LOOP) { /* the CTA's env state: HBM -> shared memory, once */
        for (int j = 0; j < k; j++) {
            const int64_t tile = tile_of(j);
            unsigned char *slot = smem + SM_STATE + j * kSlotBytes;
            const int4 *src = reinterpret_cast<const int4 *>(P.state + tile * kTileBytes);
            for (uint32_t x = tid; x < kTileBytes / 16; x += kThreads)
                reinterpret_cast<int4 *>(slot)[x] = src[x];
            for (uint32_t x = tid; x < kRows; x += kThreads) {
                const int64_t i = tile * kRows + x;
                reinterpret_cast<uint32_t *>(slot + SLOT_EPOCH)[x] = i < P.n ? P.epoch[i] : 0u;
                if (TRACK)
                    reinterpret_cast<double *>(slot + SLOT_RETURN)[x] = i < P.n ? P.ep_return[i] : 0.0;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + SM_TMEM);
    uint64_t step = A.step;
    if (A.step_device)
        step = *A.step_device;

    if (warp > kMmaWarp) {
        warpgroup_reg_dec<24>(); /* the three warps that only came to make a warpgroup */
    } else if (warp == kMmaWarp) {
        /* ================================================================ MMA issuer ============ */
        warpgroup_reg_dec<24>();
        const uint32_t leader = elect_leader();
        bar_wait_warp(bar(B_W), 0, 4096u + B_W * 128u, 0);
        const uint64_t dX = smem_desc(s0 + SM_X), dB1 = smem_desc(s0 + SM_B1), dB2 = smem_desc(s0 + SM_B2),
                       dB3 = smem_desc(s0 + SM_B3);
        /* layer 1 of sequence s2: L1 (128 x 256) = X (128 x 32, shared memory) . W1op, as two N = 128 halves
         * (the two accumulator regions of a tile are not adjacent for odd tiles) */
        auto layer1 = [&](int64_t s2) {
            const uint32_t par = (uint32_t)s2 & 1u;
#pragma unroll
            for (uint32_t half = 0; half < 2; half++)
#pragma unroll
                for (uint32_t ks = 0; ks < 2; ks++)
                    mma_bf16_ss(leader, tmem + (half ? tm_b(par) : tm_a(par)),
                                desc_at(dX, par * (kRows * 128u) + ks * 32u),
                                desc_at(dB1, half * (128u * 128u) + ks * 32u), instr_desc(128), ks > 0);
            mma_commit(leader, bar(B_L1));
        };
        bool l1_issued = false; /* layer 1 of the sequence about to start went out under the previous one */
        for (int64_t s = 0; s < S; s++) {
            const uint32_t par = (uint32_t)s & 1u, ph = (uint32_t)s & 1u;
            TRACE(s, 0);
            if (!l1_issued) {
                bar_wait_warp(bar(B_X + par), (uint32_t)(s >> 1) & 1u, 4096u + B_X * 128u, s);
                tc_fence_after();
                TRACE(s, 1);
                /* its regions are free: the previous tile's MMAs that read them were issued before these and
                 * execute before them, and this warp has seen that tile's layer-2 epilogue publish every part */
                layer1(s);
                TRACE(s, 2);
            }
            /* layer 2, quarter q: D (128 x 64) = H1 . W2[:, 64q .. 64q+63] */
            auto layer2_kstep = [&](uint32_t q, uint32_t ks, bool accumulate) {
                mma_bf16_ts(leader, tmem + tm_l2(par, q), tmem + tm_c(par) + ks * 8u,
                            desc_at(dB2, (ks >> 2) * (kHidden * 128u) + q * (64u * 128u) + (ks & 3u) * 32u),
                            instr_desc(64), accumulate);
            };
            /* Layer 2 trails the layer-1 epilogue, which publishes the activations in two halves of 128
             * columns (= 8 K-steps each; K-steps may accumulate in any order).  The first half is everything
             * region A held, so A may take quarters 2 and 3 at once: the first K-half of ALL four quarters runs
             * on the tensor pipe while the second half of the tanh runs on the XU, and after the second
             * publication quarter q is complete after q + 1 half-quarters instead of q + 1 whole ones. */
#pragma unroll
            for (uint32_t h = 0; h < 2; h++) {
                bar_wait_warp(bar(B_H1 + h), ph, 4096u + B_H1 * 128u + (h << 4), s);
                tc_fence_after();
                TRACE(s, 3 + h);
#pragma unroll
                for (uint32_t q = 0; q < 4; q++) {
#pragma unroll
                    for (uint32_t j = 0; j < 8; j++)
                        layer2_kstep(q, 8u * h + j, h + j > 0);
                    if (h == 1) {
                        mma_commit(leader, bar(B_L2 + q));
                        TRACE(s, 11 + q);
                    }
                }
            }
            /* Regions C (H1: every layer-2 MMA has been issued) and D (quarters 0 and 1, once the epilogue has
             * them in registers) of this tile are regions A and B of the next.  In the open policy step the next
             * operand has been waiting since the env rows read the previous logits, so its layer 1 goes out right
             * here, under this tile's layer-2 epilogue, and the epilogue warps run from one tile into the next
             * without waiting for the logits.  In the closed loop the next operand is the END of another tile's
             * env tick -- the loop is bound by that chain, not by the pipes -- and waiting for it here would
             * hold back this tile's logits: there layer 1 goes out at the top of the loop.  (Looking without
             * waiting, mbarrier.test_wait, is no way out: ~300 cycles of this warp per look that fails.) */
            l1_issued = false;
            if (!LOOP && s + 1 < S) {
                bar_wait_warp(bar(B_R01), ph, 4096u + B_R01 * 128u, s);
                bar_wait_warp(bar(B_X + (par ^ 1u)), (uint32_t)((s + 1) >> 1) & 1u, 4096u + B_X * 128u + 16u, s + 1);
                tc_fence_after();
                TRACE(s + 1, 1);
                layer1(s + 1);
                TRACE(s + 1, 2);
                l1_issued = true;
            }
            /* layer 3: D3 (128 x 16) += H2[:, 64q .. 64q+63] . W3 (padded), trailing the layer-2 epilogue part
             * by part.  The logits go over the first columns of quarter 3's
             * accumulator: the last place the next tile's activations reach. */
            auto layer3 = [&](uint32_t q, bool accumulate) {
                bar_wait_warp(bar(B_H2 + q), ph, 4096u + B_H2 * 128u + (q << 4), s);
                tc_fence_after();
                TRACE(s, 5 + q);
#pragma unroll
                for (uint32_t kk = 0; kk < 4; kk++) {
                    const uint32_t ks = 4u * q + kk;
                    mma_bf16_ts(leader, tmem + tm_d3(par), tmem + tm_b(par) + ks * 8u,
                                desc_at(dB3, (ks >> 2) * (kOutPad * 128u) + (ks & 3u) * 32u),
                                instr_desc(kOutPad), accumulate || kk > 0);
                }
                TRACE(s, 15 + q);
            };
            TRACE(s, 9);
            bar_wait_warp(bar(B_R3), ph, 4096u + B_R3 * 128u, s);
            TRACE(s, 10);
            layer3(0, false);
            layer3(1, true);
            layer3(2, true);
            layer3(3, true);
            mma_commit(leader, bar(B_D3 + par));
        }
    } else if (warp >= kEnvWarps) {
        /* ================================================================ tanh epilogues ======== */
        warpgroup_reg_inc<80>();
        const uint32_t quad = warp & 3u, part = (warp - kEnvWarps) >> 2; /* TMEM lanes 32 quad .., column part */
        const uint32_t lane_base = tmem + ((quad * 32u) << 16);
        bar_wait_warp(bar(B_W), 0, 8192u + B_W * 128u, 0);
        for (int64_t s = 0; s < S; s++) {
            const uint32_t ph = (uint32_t)s & 1u;
            constexpr uint32_t kPoly = Q1_POLICY_POLY_PAIRS;
            uint32_t va[32], vb[32], p[16];
            /* ---- layer-1 epilogue (bias already in the MMA) -> tanh -> H1 ---- */
            if (quad == 0) TRACE(s, 20 + 8 * part + 0);
            bar_wait_warp(bar(B_L1), ph, 8192u + B_L1 * 128u, s);
            tc_fence_after();
            if (quad == 0) TRACE(s, 20 + 8 * part + 1);
            /* Half h: this part's 32 of the accumulator columns 128 h .. 128 h + 127 (region A, then B).  A loop
             * that is NOT unrolled: unrolled, ptxas interleaves the arithmetic of the two halves for latency and
             * the first publication -- which the layer-2 MMAs are waiting for -- sinks to the end (seen in the
             * SASS as soon as the halves contain dependent FFMA2 chains).  (Written out instead, with the second
             * half's tcgen05.ld in flight under the first half's arithmetic and its wait::ld placed after the
             * first publication, the order holds too, but ptxas sinks that load next to its wait and the
             * larger live set costs more than the latency it could hide: 190.5 vs 179.3 us, measured.)
             * H1 goes where the PREVIOUS tile kept its layer-2 quarters 2 (first half) and 3 (second half, with
             * its logits over the first columns): the store waits until those have been read -- by the
             * epilogue part that took quarter 2, by the env rows -- which is long ago unless a role fell behind. */
#pragma unroll 1
            for (uint32_t h = 0; h < 2; h++) {
                tmem_ld_issue<32>(lane_base + (h ? tm_b(ph) : tm_a(ph)) + 32u * part, va);
                tmem_ld_wait<32>(va);
#pragma unroll
                for (uint32_t e = 0; e < 16; e++)
                    p[e] = tanh2_pair<kPoly>(e, __uint_as_float(va[2 * e]), __uint_as_float(va[2 * e + 1]));
                if (s >= 1) {
                    if (h == 0)
                        bar_wait_warp(bar(B_H2 + 2), ph ^ 1u, 8192u + B_H2 * 128u + (2u << 4) + 1u, s);
                    else
                        bar_wait_warp(bar(B_E), ph ^ 1u, 8192u + B_E * 128u, s);
                }
                tmem_st<16>(lane_base + tm_c(ph) + 64u * h + 16u * part, p);
                tmem_st_wait();
                tc_fence_before();
                bar_arrive(bar(B_H1 + h));
                if (quad == 0) TRACE(s, 20 + 8 * part + 2 + h);
            }
            /* ---- layer-2 epilogue: accumulator + bias -> tanh -> H2 (region B: this lane is done with the
             * layer-1 columns that were there).  Part q takes quarter q as it completes.  (Sharing the last
             * quarter between two parts shortened the tail of a tile while tiles did not overlap; now that the
             * epilogue warps run on into the next tile, equal work per part is worth more: measured.) ---- */
            auto half_quarter = [&](uint32_t q, uint32_t hh, const uint32_t *v) {  /* 32 columns -> 16 of H2 */
                const float *b = reinterpret_cast<const float *>(smem + SM_BIAS2) + 64u * q + 32u * hh;
#pragma unroll
                for (uint32_t e = 0; e < 16; e++)
                    p[e] = tanh2_pair<kPoly>(e, __uint_as_float(v[2 * e]) + b[2u * e], __uint_as_float(v[2 * e + 1]) + b[2u * e + 1u]);
                tmem_st<16>(lane_base + tm_b(ph) + 32u * q + 16u * hh, p);
            };
            {
                bar_wait_warp(bar(B_L2 + part), ph, 8192u + B_L2 * 128u + (part << 4), s);
                tc_fence_after();
                if (quad == 0) TRACE(s, 20 + 8 * part + 4);
                tmem_ld_issue<32>(lane_base + tm_l2(ph, part), va);
                tmem_ld_wait<32>(va);
                if (part == 3) { /* the columns the logits will overwrite */
                    tc_fence_before();
                    bar_arrive(bar(B_R3));
                }
                tmem_ld_issue<32>(lane_base + tm_l2(ph, part) + 32u, vb);
                half_quarter(part, 0, va);
                tmem_ld_wait<32>(vb);
                if (part < 2) { /* the whole quarter is in registers */
                    tc_fence_before();
                    bar_arrive(bar(B_R01));
                }
                half_quarter(part, 1, vb);
                tmem_st_wait();
                tc_fence_before();
                bar_arrive(bar(B_H2 + part));
                if (quad == 0) TRACE(s, 20 + 8 * part + 5);
            }
        }
    } else {
        /* ================================================================ env rows =============== */
        warpgroup_reg_inc<80>();
        /* Two groups of four warps; warp w of a group owns TMEM lanes 32 (w & 3) .. + 31 = those rows of the tile.
         * The sample + tick of a tile is one long dependent instruction stream per warp (~7000 cycles in the
         * closed loop, more than the policy phase of a tile), so with ONE group the env rows, not the tensor
         * pipe or the XU, set the pace.  With an even number of tiles per CTA (and in the open policy step) the
         * groups take alternate sequences -- disjoint tiles, so no state is shared between them; otherwise
         * group 0 takes every sequence as before and group 1 leaves. */
        const uint32_t group = warp >> 2, row = tid & 127u;
        const bool two_groups = !LOOP || (k & 1) == 0;
        const uint32_t lane_base = tmem + (((warp & 3u) * 32u) << 16);
        const float *bias3 = reinterpret_cast<const float *>(smem + SM_BIAS3);
        const int width = 2 * A.num_keys + 2;
        const int64_t stride = (LOOP && k < 2) ? 1 : 2; /* how far ahead the layer-1 operand is prepared */
        const int64_t s_first = two_groups ? group : (group == 0 ? 0 : S), s_step = two_groups ? 2 : 1;

        /* the layer-1 operand of sequence s2 -> X[s2 & 1] */
        auto prepare = [&](int64_t s2) {
            const int64_t i = tile_of(s2) * kRows + row;
            float o[kObs] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            if (LOOP) {
                Env e;
                slot_load(smem + SM_STATE + (s2 % k) * kSlotBytes, row, e);
                observe<LEAN>(P, e, o);
            } else if (i < A.n) {
                if ((reinterpret_cast<uintptr_t>(A.obs) & 7u) == 0) {
                    const float2 *p2 = reinterpret_cast<const float2 *>(A.obs + i * kObs);
                    const float2 a = __ldg(p2), b = __ldg(p2 + 1), c = __ldg(p2 + 2);
                    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y;
                } else {
#pragma unroll
                    for (int q = 0; q < kObs; q++)
                        o[q] = __ldg(A.obs + i * kObs + q);
                }
            }
            uint32_t cols[16];
            layer1_operand(o, cols);
            /* row `row` of the K-major, 128-byte-swizzled operand: 16-byte chunk j (K = 8j .. 8j+7) sits at
             * chunk position j ^ (row & 7) of the row's 128 bytes; chunks 4..7 (K >= 32) are never read */
            unsigned char *xrow = smem + SM_X + ((uint32_t)s2 & 1u) * (kRows * 128u) + row * 128u;
#pragma unroll
            for (uint32_t j = 0; j < 4; j++)
                *reinterpret_cast<uint4 *>(xrow + ((j ^ (row & 7u)) << 4)) =
                    make_uint4(cols[4 * j], cols[4 * j + 1], cols[4 * j + 2], cols[4 * j + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy stores -> the MMA's reads */
            bar_arrive(bar(B_X + ((uint32_t)s2 & 1u)));
        };
        bar_wait_warp(bar(B_W), 0, 12288u + B_W * 128u + 0u, 0);
        if (two_groups) {
            if ((int64_t)group < S)
                prepare(group);
        } else {
            for (int64_t s2 = s_first; s2 < stride && s2 < S; s2++)
                prepare(s2);
        }

        float rsum[kMaxTiles];
#pragma unroll
        for (int j = 0; j < kMaxTiles; j++)
            rsum[j] = 0.0f;
        for (int64_t s = s_first; s < S; s += s_step) {
            if ((warp & 3u) == 0) TRACE(s, 52);
            /* the noise of this row's action needs no logits: drawn while the policy runs */
            const ActionNoise noise = draw_action_noise(
                A.deterministic != 0, A.seed, LOOP ? step + (uint64_t)(s / k) : step,
                (LOOP ? P.env_index_base : A.env_index_base) + (uint64_t)(tile_of(s) * kRows + row));
            bar_wait_warp(bar(B_D3 + ((uint32_t)s & 1u)), (uint32_t)(s >> 1) & 1u, 12288u + B_D3 * 128u, s);
            tc_fence_after();
            if ((warp & 3u) == 0) TRACE(s, 53);
            uint32_t v[16];
            tmem_ld16(lane_base + tm_d3((uint32_t)s & 1u), v);
            tc_fence_before();
            bar_arrive(bar(B_E));
            if ((warp & 3u) == 0) TRACE(s, 54);
            float lg[10];
#pragma unroll
            for (int q = 0; q < 10; q++)
                lg[q] = __uint_as_float(v[q]) + bias3[q];
            if (!LOOP && s + stride < S) /* nothing of it depends on this tile's action: the MMA warp wants it early */
                prepare(s + stride);
            const int64_t tile = tile_of(s);
            const int64_t i = tile * kRows + row;
            const bool active = i < (LOOP ? P.n : A.n);
            if (!LOOP) {
                if (active) {
                    float m;
                    const uint32_t kb = apply_action_noise(lg, A.num_keys, A.low, A.high, A.deterministic != 0, noise, &m);
                    for (int q = 0; q < A.num_keys; q++)
                        A.keys[i * A.num_keys + q] = (kb >> q) & 1u;
                    A.mouse[i] = m;
                    if (A.logits_out)
                        for (int q = 0; q < width; q++)
                            A.logits_out[i * width + q] = lg[q];
                }
            } else {
                /* ---- one env tick: env.VectorPhysEnv.vector_step (env:482-510) on the sampled action ---- */
                const int slot_no = (int)(s % k);
                const int64_t tick_no = s / k;
                unsigned char *slot = smem + SM_STATE + slot_no * kSlotBytes;
                const uint64_t gidx = P.env_index_base + (uint64_t)i;
                float m;
                const uint32_t keybits = apply_action_noise(lg, A.num_keys, A.low, A.high, A.deterministic != 0, noise, &m);
#if Q1_ACTOR_TRACE == 1 /* the stamp is only a stamp if the value exists by then */
                if (__float_as_uint(m) == 0x7fc12345u && keybits == 77u) TRACE(s, 59);
#endif
                if ((warp & 3u) == 0) TRACE(s, 57);
                Env e;
                slot_load(slot, row, e);
                float r;
                bool d;
                if (RECORD) {
                    const int64_t rrow = (A.rec_tick0 + tick_no) * P.n + i;
                    float o[6];
                    observe<LEAN>(P, e, o);
                    if (active)
                        record_before(A.rec, rrow, A.num_keys, e, o, keybits, (double)m);
                    Move mv;
                    tick<false, LEAN, false>(P, e, keybits, (double)m, r, d, &mv);
                    if (active)
                        record_after(A.rec, rrow, A.record_flags, P.jump_mode, mv, o, r, d);
                } else {
                    tick<false, LEAN, false>(P, e, keybits, (double)m, r, d);
                }
#if Q1_ACTOR_TRACE == 1
                if (__float_as_uint(r) == 0x7fc12345u && d) TRACE(s, 59);
#endif
                if ((warp & 3u) == 0) TRACE(s, 58);
#pragma unroll
                for (int j = 0; j < kMaxTiles; j++)
                    if (j == slot_no)
                        rsum[j] = add32(rsum[j], r);
                if (TRACK) {
                    double *retp = reinterpret_cast<double *>(slot + SLOT_RETURN) + row;
                    double ret = add64(*retp, (double)r);
                    bool finished = active && d;
                    if (!A.auto_reset) { /* report an episode once, as the step kernels do */
                        finished = finished && !(e.bits & F_DONE_SEEN);
                        if (finished)
                            e.bits |= F_DONE_SEEN;
                    }
                    report_episodes(P, finished, e.bits & F_ZERO_START, ret);
                    *retp = (d && A.auto_reset) ? 0.0 : ret;
                }
                if (d && A.auto_reset) {
                    uint32_t *epp = reinterpret_cast<uint32_t *>(slot + SLOT_EPOCH) + row;
                    const uint32_t ep = *epp + 1u;
                    *epp = ep;
                    reset_env<false>(P, e, gidx, ep);
                }
                slot_store(slot, row, e);
            }
            if ((warp & 3u) == 0) TRACE(s, 55);
            if (LOOP && s + stride < S)
                prepare(s + stride); /* reads only this thread's own row of the slot */
            if ((warp & 3u) == 0) TRACE(s, 56);
        }
        if (LOOP) { /* results of the launch: final observation and per-env reward sum */
            for (int j = 0; j < k; j++) {
                const int64_t i = tile_of(j) * kRows + row;
                if (i >= P.n || (two_groups ? ((uint32_t)j & 1u) != group : group != 0))
                    continue;
                if (A.final_obs) {
                    Env e;
                    slot_load(smem + SM_STATE + j * kSlotBytes, row, e);
                    float o[6];
                    observe<LEAN>(P, e, o);
                    store_obs(A.final_obs, i, o);
                }
                if (A.reward_sum) {
                    float acc = 0.0f;
#pragma unroll
                    for (int q = 0; q < kMaxTiles; q++)
                        if (q == j)
                            acc = rsum[q];
                    A.reward_sum[i] = acc;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (LOOP) { /* the CTA's env state: shared memory -> HBM */
        for (int j = 0; j < k; j++) {
            const int64_t tile = tile_of(j);
            const unsigned char *slot = smem + SM_STATE + j * kSlotBytes;
            int4 *dst = reinterpret_cast<int4 *>(P.state + tile * kTileBytes);
            for (uint32_t x = tid; x < kTileBytes / 16; x += kThreads)
                dst[x] = reinterpret_cast<const int4 *>(slot)[x];
            for (uint32_t x = tid; x < kRows; x += kThreads) {
                const int64_t i = tile * kRows + x;
                if (i < P.n) {
                    P.epoch[i] = reinterpret_cast<const uint32_t *>(slot + SLOT_EPOCH)[x];
                    if (TRACK)
                        P.ep_return[i] = reinterpret_cast<const double *>(slot + SLOT_RETURN)[x];
                }
            }
        }
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
#if Q1_ACTOR_TRACE
    if (blockIdx.x == Q1_ACTOR_TRACE_BLOCK && tid == 0) {
        g_trace[0][63] = clock64();
        unsigned long long ns;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
        g_trace[0][61] = (long long)ns;
    }
    if (tid == 0 && blockIdx.x < 256) {
        unsigned smid;
        asm("mov.u32 %0, %smid;" : "=r"(smid));
        g_block_cycles[blockIdx.x] = ((clock64() - trace_entry) << 8) | (long long)(smid & 255u);
    }
#endif
}

uint16_t to_bf16(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u)
        return (uint16_t)((u >> 16) | 0x40u);
    u += 0x7FFFu + ((u >> 16) & 1u); /* round to nearest even */
    return (uint16_t)(u >> 16);
}
float from_bf16(uint16_t h)
{
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

struct DeviceScope { /* switch to a device for the duration of a call */
    int prev = -1;
    bool ok = true;
    explicit DeviceScope(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess)
            prev = -1;
        if (prev == dev)
            prev = -1;
        else
            ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceScope()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

/* Reads (and clears) the watchdog record of the kernels launched so far on the current device; the
 * caller has synchronised.  -> Q1_OK or Q1_ECUDA with the record in the error text. */
int check_fault()
{
    unsigned int f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(f, g_fault, sizeof f) != cudaSuccess)
        return q1_set_error(Q1_ECUDA, "k_actor: cannot read the watchdog record");
    if (!f[0])
        return Q1_OK;
    const unsigned int zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_fault, zero, sizeof zero);
    static const char *roles[] = {"?", "MMA", "EPI", "ENV"};
    std::string others;
    for (int r = 1; r <= 3; r++)
        if (f[4 + r])
            others += std::string("; a waiting ") + roles[r] + " thread stood at barrier " +
                      std::to_string((f[4 + r] >> 7) & 31u) + " sub-step " + std::to_string((f[4 + r] >> 4) & 7u) +
                      " sequence " + std::to_string((f[4 + r] >> 12) & 0x7FFFFu);
    return q1_set_error(Q1_ECUDA, std::string("k_actor watchdog: the ") + roles[(f[1] >> 12) & 3u] +
                                      " role gave up waiting for barrier " + std::to_string((f[1] >> 7) & 31u) +
                                      " (sub-step " + std::to_string((f[1] >> 4) & 7u) + ") at sequence " +
                                      std::to_string(f[2]) + " in CTA " + std::to_string(f[3]) + others +
                                      "; the launch's results are invalid");
}

template <typename K> cudaError_t allow_smem(K kernel, uint32_t bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

} // namespace

extern "C" {

int q1_policy_create(int device, int num_keys, const float *w1, const float *b1, const float *w2,
                     const float *b2, const float *w3, const float *b3, q1_policy **out)
{
    if (!out || !w1 || !b1 || !w2 || !b2 || !w3 || !b3)
        return q1_set_error(Q1_EINVAL, "a weight array / out is NULL");
    *out = nullptr;
    if (num_keys != 3 && num_keys != 4)
        return q1_set_error(Q1_EINVAL, "num_keys must be 3 or 4");
    const int width = 2 * num_keys + 2;
    std::vector<unsigned char> img(kImageBytes, 0);
    auto put = [&](uint32_t region, uint32_t rows, int nrow, int k, uint16_t h) {
        /* element (row nrow, K index k) of an operand stored K-major with 128-byte swizzle */
        const uint32_t c = (uint32_t)k >> 3, e = (uint32_t)k & 7u;
        const uint32_t off = region - SM_B2 + (c >> 3) * (rows * 128u) + (uint32_t)nrow * 128u +
                             (((c & 7u) ^ ((uint32_t)nrow & 7u)) << 4) + e * 2u;
        memcpy(&img[off], &h, 2);
    };
    for (int k = 0; k < kHidden; k++)
        for (int nrow = 0; nrow < kHidden; nrow++)
            put(SM_B2, kHidden, nrow, k, to_bf16(w2[k * kHidden + nrow])); /* fc_2 kernel is (in, out) */
    for (int k = 0; k < kHidden; k++)
        for (int nrow = 0; nrow < width; nrow++)
            put(SM_B3, kOutPad, nrow, k, to_bf16(w3[k * width + nrow]));
    /* layer 1 as a K = 32 operand (see layer1_operand): per input (w_hi, w_lo, w_hi, w_lo, w_hi), then the
     * bias split the same way against two columns of ones */
    for (int u = 0; u < kHidden; u++) {
        for (int k = 0; k < kObs; k++) {
            const float w = w1[k * kHidden + u];            /* fc_1 kernel is (in, out) */
            const uint16_t hi = to_bf16(w), lo = to_bf16(w - from_bf16(hi));
            const uint16_t pat[5] = {hi, lo, hi, lo, hi};
            for (int q = 0; q < 5; q++)
                put(SM_B1, kHidden, u, 5 * k + q, pat[q]);
        }
        const uint16_t bh = to_bf16(b1[u]), bl = to_bf16(b1[u] - from_bf16(bh));
        put(SM_B1, kHidden, u, 30, bh);
        put(SM_B1, kHidden, u, 31, bl);
    }
    memcpy(&img[SM_BIAS2 - SM_B2], b2, kHidden * 4);
    memcpy(&img[SM_BIAS3 - SM_B2], b3, width * 4);

    DeviceScope scope(device);
    if (!scope.ok)
        return q1_set_error(Q1_ENODEV, "cudaSetDevice failed: libq1phys has no CPU implementation");
    q1_policy *p = new (std::nothrow) q1_policy();
    if (!p)
        return q1_set_error(Q1_ENOMEM, "out of host memory");
    p->device = device;
    p->num_keys = num_keys;
    cudaError_t err = cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (err == cudaSuccess)
        err = cudaMalloc(&p->image, kImageBytes);
    if (err == cudaSuccess)
        err = cudaMemcpy(p->image, img.data(), kImageBytes, cudaMemcpyHostToDevice);
    if (err == cudaSuccess)
        err = allow_smem(k_actor<false, false, true, false>, SM_TOTAL_ACT);
#define Q1_ALLOW(TR, LN, RC)                                                    \
    if (err == cudaSuccess)                                                     \
        err = allow_smem(k_actor<true, TR, LN, RC>, SM_TOTAL_LOOP);
    Q1_ALLOW(false, false, false) Q1_ALLOW(false, false, true) Q1_ALLOW(false, true, false)
    Q1_ALLOW(false, true, true) Q1_ALLOW(true, false, false) Q1_ALLOW(true, false, true)
    Q1_ALLOW(true, true, false) Q1_ALLOW(true, true, true)
#undef Q1_ALLOW
    if (err != cudaSuccess) {
        if (p->image)
            cudaFree(p->image);
        delete p;
        return q1_set_error(Q1_ECUDA, std::string("q1_policy_create: ") + cudaGetErrorString(err));
    }
    *out = p;
    return Q1_OK;
}

#if Q1_ACTOR_TRACE
int q1_actor_trace(long long *out) /* 16 x 64 stamps of CTA 0 */
{
    return cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 16 * 64) == cudaSuccess ? 0 : -1;
}
int q1_actor_block_cycles(long long *out) /* 256 entries, see g_block_cycles */
{
    return cudaMemcpyFromSymbol(out, g_block_cycles, sizeof(long long) * 256) == cudaSuccess ? 0 : -1;
}
#endif

/* Synchronises the device and reports a watchdog record of the policy kernels, if any. */
int q1_policy_check(q1_policy *p)
{
    if (!p)
        return q1_set_error(Q1_EINVAL, "policy is NULL");
    DeviceScope scope(p->device);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess)
        return q1_set_error(Q1_ECUDA, std::string("q1_policy_check: ") + cudaGetErrorString(err));
    return check_fault();
}

int q1_policy_destroy(q1_policy *p)
{
    if (!p)
        return Q1_OK;
    DeviceScope scope(p->device);
    cudaError_t err = cudaFree(p->image);
    delete p;
    if (err != cudaSuccess)
        return q1_set_error(Q1_ECUDA, std::string("q1_policy_destroy: ") + cudaGetErrorString(err));
    return Q1_OK;
}

int q1_policy_act(q1_policy *p, int64_t n, const float *obs, double action_low, double action_high,
                  int deterministic, uint64_t seed, uint64_t step, const uint64_t *step_device,
                  uint64_t env_index_base, uint8_t *keys, float *mouse, float *logits_out, void *stream)
{
    if (!p || !obs || !keys || !mouse)
        return q1_set_error(Q1_EINVAL, "policy / obs / keys / mouse is NULL");
    if (n < 0)
        return q1_set_error(Q1_EINVAL, "n must be >= 0");
    if (n == 0)
        return Q1_OK;
    DeviceScope scope(p->device);
    if (!scope.ok)
        return q1_set_error(Q1_ECUDA, "cudaSetDevice failed");
    ActorArgs a = {};
    a.image = p->image;
    a.n = n;
    a.num_keys = p->num_keys;
    a.low = (float)action_low;
    a.high = (float)action_high;
    a.deterministic = deterministic;
    a.seed = seed;
    a.step = step;
    a.step_device = step_device;
    a.env_index_base = env_index_base;
    a.obs = obs;
    a.keys = keys;
    a.mouse = mouse;
    a.logits_out = logits_out;
    const int64_t tiles = (n + kRows - 1) / kRows;
    const unsigned grid = (unsigned)(tiles < p->sm_count ? tiles : p->sm_count);
    Params unused = {};
    k_actor<false, false, true, false><<<grid, kThreads, SM_TOTAL_ACT, static_cast<cudaStream_t>(stream)>>>(unused, a);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
        return q1_set_error(Q1_ECUDA, std::string("k_actor launch: ") + cudaGetErrorString(err));
    return Q1_OK;
}

/* see include/q1phys.h */
static int rollout_launch(q1_policy *p, const q1_env_view &ev, int ticks, int auto_reset, int deterministic,
                          uint64_t seed, double action_low, double action_high, uint32_t record_flags,
                          const q1_record_view *rec, float *final_obs, float *reward_sum, cudaStream_t s)
{
    ActorArgs a = {};
    a.image = p->image;
    a.n = ev.P.n;
    a.num_keys = p->num_keys;
    a.low = (float)action_low;
    a.high = (float)action_high;
    a.deterministic = deterministic;
    a.seed = seed;
    a.step = ev.ticks;
    a.env_index_base = ev.P.env_index_base;
    a.ticks = ticks;
    a.auto_reset = auto_reset;
    a.record_flags = record_flags;
    if (rec)
        a.rec = *rec;
    a.final_obs = final_obs;
    a.reward_sum = reward_sum;
    const int64_t tiles = (ev.P.n + kRows - 1) / kRows;
    const int64_t per_launch = (int64_t)p->sm_count * kMaxTiles;
    for (int64_t t0 = 0; t0 < tiles; t0 += per_launch) {
        a.tile_begin = t0;
        a.tile_end = t0 + per_launch < tiles ? t0 + per_launch : tiles;
        const int64_t count = a.tile_end - a.tile_begin;
        /* as many CTAs as SMs, unless there are fewer tiles; tiles spread evenly over them */
        const unsigned grid = (unsigned)(count < p->sm_count ? count : p->sm_count);
#define Q1_LAUNCH(TR, LN, RC) \
    k_actor<true, TR, LN, RC><<<grid, kThreads, SM_TOTAL_LOOP, s>>>(ev.P, a)
        const bool lean = !ev.P.ieee_div;
        if (ev.track) {
            if (lean) { if (rec) Q1_LAUNCH(true, true, true); else Q1_LAUNCH(true, true, false); }
            else { if (rec) Q1_LAUNCH(true, false, true); else Q1_LAUNCH(true, false, false); }
        } else {
            if (lean) { if (rec) Q1_LAUNCH(false, true, true); else Q1_LAUNCH(false, true, false); }
            else { if (rec) Q1_LAUNCH(false, false, true); else Q1_LAUNCH(false, false, false); }
        }
#undef Q1_LAUNCH
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess)
            return q1_set_error(Q1_ECUDA, std::string("k_actor launch: ") + cudaGetErrorString(err));
    }
    return Q1_OK;
}

static int rollout_check(q1_policy *p, q1_env *env, int ticks, const q1_env_view &ev)
{
    (void)env;
    if (ticks < 0)
        return q1_set_error(Q1_EINVAL, "ticks must be >= 0");
    if (ev.stamps)
        return q1_set_error(Q1_EINVAL, "q1_policy_rollout needs the counter form of the key timers (not "
                                       "Q1_F_FORCE_F64_STAMPS / a config that forces f64 stamps): step such "
                                       "envs with q1_policy_act + q1_step");
    if (ev.P.num_keys != p->num_keys)
        return q1_set_error(Q1_EINVAL, "the policy's action head does not match the env's key count");
    if (ev.device != p->device)
        return q1_set_error(Q1_EINVAL, "policy and env live on different devices");
    return Q1_OK;
}

int q1_policy_rollout(q1_policy *p, q1_env *env, int ticks, int auto_reset, int deterministic, uint64_t seed,
                      double action_low, double action_high, uint32_t record_flags,
                      const q1_record_view *record, float *final_obs, float *reward_sum, void *stream)
{
    if (!p || !env)
        return q1_set_error(Q1_EINVAL, "policy / env is NULL");
    q1_env_view ev;
    int rc = q1_env_get_view(env, true, &ev);
    if (rc == Q1_OK)
        rc = rollout_check(p, env, ticks, ev);
    if (rc != Q1_OK)
        return rc;
    if (ticks == 0)
        return Q1_OK;
    DeviceScope scope(p->device);
    rc = rollout_launch(p, ev, ticks, auto_reset, deterministic, seed, action_low, action_high, record_flags,
                        record, final_obs, reward_sum, static_cast<cudaStream_t>(stream));
    if (rc == Q1_OK)
        rc = q1_advance_ticks(env, ticks);
    return rc;
}

int q1_policy_rollout_host(q1_policy *p, q1_env *env, int ticks, int auto_reset, int deterministic,
                           uint64_t seed, double action_low, double action_high, uint32_t record_flags,
                           const q1_record_view *record, float *final_obs_host)
{
    if (!p || !env || !record)
        return q1_set_error(Q1_EINVAL, "policy / env / record is NULL");
    q1_env_view ev;
    int rc = q1_env_get_view(env, false, &ev);
    if (rc == Q1_OK)
        rc = rollout_check(p, env, ticks, ev);
    if (rc != Q1_OK)
        return rc;
    if (ticks == 0)
        return Q1_OK;
    const size_t n = (size_t)ev.P.n, nk = (size_t)ev.P.num_keys, rows = n * (size_t)ticks;
    const void *hosts[15] = {record->vel, record->z_pos, record->on_ground, record->jump_released,
                             record->time_remaining, record->obs, record->keys, record->mouse,
                             record->yaw, record->smove, record->fmove, record->jump, record->reward,
                             record->done, final_obs_host};
    const size_t width[15] = {12, 8, 1, 1, 8, 24, nk, 4, 8, 8, 8, 1, 4, 1, 0};
    size_t bytes[15], offs[15], off = 0;
    for (int q = 0; q < 15; q++) {
        bytes[q] = hosts[q] ? (q == 14 ? 24 * n : rows * width[q]) : 0;
        offs[q] = off;
        off = (off + bytes[q] + 255) & ~(size_t)255;
    }
    void *scratch = nullptr, *stream = nullptr;
    rc = q1_env_host_scratch(env, off + 256, &scratch, &stream);
    if (rc != Q1_OK)
        return rc;
    DeviceScope scope(p->device);
    char *d = static_cast<char *>(scratch);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    void *dev[15];
    for (int q = 0; q < 15; q++)
        dev[q] = bytes[q] ? d + offs[q] : nullptr;
    q1_record_view dv;
    dv.vel = static_cast<float *>(dev[0]);
    dv.z_pos = static_cast<double *>(dev[1]);
    dv.on_ground = static_cast<uint8_t *>(dev[2]);
    dv.jump_released = static_cast<uint8_t *>(dev[3]);
    dv.time_remaining = static_cast<double *>(dev[4]);
    dv.obs = static_cast<float *>(dev[5]);
    dv.keys = static_cast<uint8_t *>(dev[6]);
    dv.mouse = static_cast<float *>(dev[7]);
    dv.yaw = static_cast<double *>(dev[8]);
    dv.smove = static_cast<int64_t *>(dev[9]);
    dv.fmove = static_cast<int64_t *>(dev[10]);
    dv.jump = static_cast<uint8_t *>(dev[11]);
    dv.reward = static_cast<float *>(dev[12]);
    dv.done = static_cast<uint8_t *>(dev[13]);
    rc = rollout_launch(p, ev, ticks, auto_reset, deterministic, seed, action_low, action_high, record_flags,
                        &dv, static_cast<float *>(dev[14]), nullptr, s);
    if (rc != Q1_OK)
        return rc;
    for (int q = 0; q < 15; q++)
        if (bytes[q]) {
            cudaError_t err = cudaMemcpyAsync(const_cast<void *>(hosts[q]), dev[q], bytes[q], cudaMemcpyDeviceToHost, s);
            if (err != cudaSuccess)
                return q1_set_error(Q1_ECUDA, std::string("q1_policy_rollout_host: ") + cudaGetErrorString(err));
        }
    cudaError_t err = cudaStreamSynchronize(s);
    if (err != cudaSuccess)
        return q1_set_error(Q1_ECUDA, std::string("q1_policy_rollout_host: ") + cudaGetErrorString(err));
    rc = check_fault();
    if (rc != Q1_OK)
        return rc;
    return q1_advance_ticks(env, ticks);
}

} /* extern "C" */
