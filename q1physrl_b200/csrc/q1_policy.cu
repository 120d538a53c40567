/*
 * q1_policy.cu -- the shipped PPO policy fused into ONE sm_100a kernel: obs -> tanh 256 -> tanh 256
 * -> logits -> Q1PhysActionDist sample -> the action arrays q1_step consumes.
 * Reference: q1physrl checkpoints (RLLib fcnet default_policy/fc_1, fc_2, fc_out) and
 * q1physrl/action_dist.py.
 *
 * This is the one dense contraction near the path, so it is the one place the 5th-generation
 * tensor cores are used.  Per CTA (one per SM, 512 threads = 128 rows x 4 column groups, persistent
 * over tiles of 128 envs); the hidden activations live in TENSOR MEMORY only:
 *   layer 1 (K = 6)    fp32 on the CUDA cores from the fp32 observation (bf16 would quantise yaw /
 *                      90 to ~3 degrees), tanh.approx.bf16x2, tcgen05.st into TMEM as the A operand
 *   layer 2 (256x256)  16 x tcgen05.mma (M = 128, N = 256, K = 16; A from TMEM, B = W2^T from shared
 *                      memory, fp32 accumulator in TMEM), issued by one thread, completion through
 *                      tcgen05.commit on an mbarrier.  It runs while the CUDA cores compute layer 1
 *                      of the NEXT tile into the other activation buffer.
 *   epilogue 2         tcgen05.ld 32 columns at a time -> + bias -> tanh -> bf16 -> tcgen05.st over
 *                      the consumed layer-1 activations
 *   layer 3 (256x10)   16 x tcgen05.mma with N = 16 (weights zero-padded), accumulator over the
 *                      consumed layer-2 columns
 *   epilogue 3         tcgen05.ld the logit columns -> + bias -> sample_action_row -> keys / mouse
 * TMEM budget (512 columns): two activation buffers of 128 columns (256 bf16 per lane) + 256
 * accumulator columns.  The weights (128 KB of bf16 W2^T pre-swizzled into the UMMA shared-memory
 * layout on the host, 8 KB W3^T, fp32 W1 and biases) are copied into shared memory once per CTA
 * with cp.async.bulk.
 */
#include "../../include/q1phys.h"
#include "q1_sample.cuh"

#include <cuda_bf16.h>

#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace q1;

int q1_set_error(int code, const std::string &msg); /* q1phys.cu: thread-local last error */

namespace {

constexpr int kRows = 128;   /* envs per tile = UMMA M */
constexpr int kHidden = 256; /* hidden width = K of layers 2 and 3, N of layer 2 */
constexpr int kOutPad = 16;  /* layer-3 N, zero-padded from 2 * num_keys + 2 = 8 or 10 */
constexpr int kObs = 6;
constexpr int kGroups = 4;   /* column groups: thread = (row, group); 4 warps per scheduler hide latency */
constexpr int kThreads = kRows * kGroups;

/* shared-memory image; every UMMA operand block is 1024-byte aligned (128-byte swizzle atoms) */
constexpr uint32_t SM_B2 = 0;                                  /* 4 K-atoms x 256 rows x 128 B */
constexpr uint32_t SM_B3 = SM_B2 + 4 * kHidden * 128;          /* 4 K-atoms x 16 rows x 128 B */
constexpr uint32_t SM_W1 = SM_B3 + 4 * kOutPad * 128;          /* fp32 [6][256] */
constexpr uint32_t SM_B1 = SM_W1 + kObs * kHidden * 4;         /* fp32 [256] */
constexpr uint32_t SM_BIAS2 = SM_B1 + kHidden * 4;             /* fp32 [256] */
constexpr uint32_t SM_BIAS3 = SM_BIAS2 + kHidden * 4;          /* fp32 [16] */
constexpr uint32_t SM_WEIGHTS_END = SM_BIAS3 + kOutPad * 4;
constexpr uint32_t SM_BAR = (SM_WEIGHTS_END + 15u) & ~15u;     /* 3 mbarriers */
constexpr uint32_t SM_TMEM = SM_BAR + 32;                      /* TMEM base address */
constexpr uint32_t SM_TOTAL = SM_TMEM + 16;
constexpr uint32_t kImageBytes = SM_WEIGHTS_END - SM_B2;        /* what the host image holds */
static_assert(kImageBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

/* tensor-memory columns (32-bit): two activation buffers (256 bf16 per lane each), accumulators */
constexpr uint32_t TM_H0 = 0, TM_H1 = 128, TM_D = 256;

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

__device__ __forceinline__ uint32_t saddr_of(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* A operand from tensor memory (lane = row, 16-bit elements packed two per column along K) */
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            bool accumulate)
{
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
                 "}\n" ::"r"(tmem_d),
                 "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
/* tanh of two pre-activations -> packed bf16x2.  Q1_POLICY_TANH_BF16X2 rounds the inputs to bf16
 * first and spends one MUFU on the pair (faster, ~3x the logit error); the default keeps fp32
 * inputs (tanh.approx.f32, relative error 2^-11) and rounds only the results. */
#ifndef Q1_POLICY_TANH_BF16X2
#define Q1_POLICY_TANH_BF16X2 0
#endif
__device__ __forceinline__ uint32_t tanh2_bf16(float lo, float hi)
{
#if Q1_POLICY_TANH_BF16X2
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
/* 16 consecutive 32-bit columns of this thread's TMEM lane */
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t v[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
/* 32 consecutive fp32 accumulator columns of this thread's TMEM lane */
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32])
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
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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

__global__ void __launch_bounds__(kThreads, 1)
k_policy_act(const unsigned char *__restrict__ image, int64_t n, int num_keys,
             const float *__restrict__ obs, float low, float high, int deterministic, uint64_t seed,
             uint64_t step, const uint64_t *__restrict__ step_device, uint64_t env_index_base,
             uint8_t *__restrict__ keys, float *__restrict__ mouse, float *__restrict__ logits_out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t row = tid % kRows, group = tid / kRows; /* warp % 4 == row / 32: its TMEM lanes */
    const uint32_t s0 = saddr_of(smem);
    const uint32_t bar_w = s0 + SM_BAR, bar_l2 = s0 + SM_BAR + 8, bar_l3 = s0 + SM_BAR + 16;
    const int width = 2 * num_keys + 2;
    if (step_device)
        step = *step_device;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_w) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_l2) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_l3) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        /* the weight image -> shared memory, in 32 KB bulk copies */
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_w), "r"(kImageBytes)
                     : "memory");
        for (uint32_t off = 0; off < kImageBytes; off += 32768u) {
            const uint32_t len = kImageBytes - off < 32768u ? kImageBytes - off : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s0 + SM_B2 + off), "l"(image + off), "r"(len), "r"(bar_w)
                         : "memory");
        }
    }
    if (warp == 0) { /* one warp owns the TMEM allocation: 512 columns (256 for layer 2, 16 for layer 3) */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s0 + SM_TMEM), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + SM_TMEM);
    const uint32_t tmem_lane = tmem + (((warp & 3u) * 32u) << 16); /* this warp's 32 TMEM lanes */
    bar_wait(bar_w, 0);

    const float4 *w1 = reinterpret_cast<const float4 *>(smem + SM_W1);
    const float4 *b1 = reinterpret_cast<const float4 *>(smem + SM_B1);
    const float *bias2 = reinterpret_cast<const float *>(smem + SM_BIAS2);
    const float *bias3 = reinterpret_cast<const float *>(smem + SM_BIAS3);
    const int64_t tiles = (n + kRows - 1) / kRows;
    constexpr uint32_t kUnits = kHidden / kGroups;          /* hidden units per thread: 64 */

    /* layer 1 of one tile: this thread's 64 hidden units of its row -> 32 TMEM columns of buffer h */
    auto layer1 = [&](int64_t tile, uint32_t tm_h) {
        const int64_t i = tile * kRows + row;
        float o[kObs] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        if (i < n) {
            if ((reinterpret_cast<uintptr_t>(obs) & 7u) == 0) {
                const float2 *p2 = reinterpret_cast<const float2 *>(obs + i * kObs);
                const float2 a = __ldg(p2), b = __ldg(p2 + 1), c = __ldg(p2 + 2);
                o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y;
            } else {
#pragma unroll
                for (int k = 0; k < kObs; k++)
                    o[k] = __ldg(obs + i * kObs + k);
            }
        }
#pragma unroll
        for (uint32_t half = 0; half < 2; half++) {
            uint32_t cols[16];
#pragma unroll
            for (uint32_t q = 0; q < 8; q++) {              /* 4 units -> 2 packed columns */
                const uint32_t g4 = (group * kUnits + half * 32u + q * 4u) / 4u;
                float4 acc = b1[g4];
#pragma unroll
                for (int k = 0; k < kObs; k++) {
                    const float4 w = w1[k * (kHidden / 4) + g4];
                    acc.x = fmaf(o[k], w.x, acc.x);
                    acc.y = fmaf(o[k], w.y, acc.y);
                    acc.z = fmaf(o[k], w.z, acc.z);
                    acc.w = fmaf(o[k], w.w, acc.w);
                }
                cols[2 * q] = tanh2_bf16(acc.x, acc.y);
                cols[2 * q + 1] = tanh2_bf16(acc.z, acc.w);
            }
            tmem_st16(tmem_lane + tm_h + group * (kUnits / 2) + half * 16u, cols);
        }
        tmem_st_wait();
    };

    uint32_t parity = 0, hb = 0;
    if ((int64_t)blockIdx.x < tiles)
        layer1(blockIdx.x, TM_H0);
    tc_fence_before();
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t tm_h = hb ? TM_H1 : TM_H0, tm_hn = hb ? TM_H0 : TM_H1;
        const int64_t i = tile * kRows + row;
        const bool active = i < n;
        /* ---- layer 2: D[128 x 256] = H1[128 x 256] . W2, A from tensor memory ---- */
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < kHidden / 16; k++) {
                const uint64_t db = smem_desc(s0 + SM_B2 + (k >> 2) * (kHidden * 128u) + (k & 3u) * 32u);
                mma_bf16_ts(tmem + TM_D, tmem + tm_h + k * 8u, db, instr_desc(kHidden), k > 0);
            }
            mma_commit(bar_l2);
        }
        /* ---- meanwhile: layer 1 of the next tile into the other activation buffer ---- */
        if (tile + gridDim.x < tiles)
            layer1(tile + gridDim.x, tm_hn);
        bar_wait(bar_l2, parity);
        tc_fence_after();
        /* ---- epilogue 2: + bias, tanh, bf16, over the consumed layer-1 activations ---- */
#pragma unroll 1
        for (uint32_t c32 = group * (kHidden / 32 / kGroups); c32 < (group + 1) * (kHidden / 32 / kGroups); c32++) {
            uint32_t v[32], cols[16];
            tmem_ld32(tmem_lane + TM_D + c32 * 32u, v);
#pragma unroll
            for (uint32_t e = 0; e < 16; e++) {
                const uint32_t col = c32 * 32u + 2u * e;
                cols[e] = tanh2_bf16(__uint_as_float(v[2 * e]) + bias2[col],
                                     __uint_as_float(v[2 * e + 1]) + bias2[col + 1u]);
            }
            tmem_st16(tmem_lane + tm_h + c32 * 16u, cols);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        /* ---- layer 3: D3[128 x 16] = H2[128 x 256] . W3 (padded), over the consumed columns of D ---- */
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < kHidden / 16; k++) {
                const uint64_t db = smem_desc(s0 + SM_B3 + (k >> 2) * (kOutPad * 128u) + (k & 3u) * 32u);
                mma_bf16_ts(tmem + TM_D, tmem + tm_h + k * 8u, db, instr_desc(kOutPad), k > 0);
            }
            mma_commit(bar_l3);
        }
        bar_wait(bar_l3, parity);
        tc_fence_after();
        /* ---- epilogue 3: logits -> sampled action (one thread per row) ---- */
        if (group == 0) {
            uint32_t v[16];
            tmem_ld16(tmem_lane + TM_D, v);
            float lg[10];
#pragma unroll
            for (int k = 0; k < 10; k++)
                lg[k] = __uint_as_float(v[k]) + bias3[k];
            if (active) {
                float m;
                const uint32_t kb = sample_action_row(lg, num_keys, low, high, deterministic != 0, seed,
                                                      step, env_index_base + (uint64_t)i, &m);
                for (int k = 0; k < num_keys; k++)
                    keys[i * num_keys + k] = (kb >> k) & 1u;
                mouse[i] = m;
                if (logits_out)
                    for (int k = 0; k < width; k++)
                        logits_out[i * width + k] = lg[k];
            }
        }
        tc_fence_before(); /* the next tile's MMAs overwrite the accumulators group 0 just read */
        __syncthreads();
        parity ^= 1u;
        hb ^= 1u;
    }
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
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

} // namespace

struct q1_policy {
    int device = 0;
    int num_keys = 4;
    int sm_count = 148;
    unsigned char *image = nullptr; /* device copy of the shared-memory weight image */
};

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
    auto put = [&](uint32_t region, uint32_t rows, int nrow, int k, float value) {
        /* element (row nrow, K index k) of an operand stored K-major with 128-byte swizzle */
        const uint32_t c = (uint32_t)k >> 3, e = (uint32_t)k & 7u;
        const uint32_t off = region - SM_B2 + (c >> 3) * (rows * 128u) + (uint32_t)nrow * 128u +
                             (((c & 7u) ^ ((uint32_t)nrow & 7u)) << 4) + e * 2u;
        const uint16_t h = to_bf16(value);
        memcpy(&img[off], &h, 2);
    };
    for (int k = 0; k < kHidden; k++)
        for (int nrow = 0; nrow < kHidden; nrow++)
            put(SM_B2, kHidden, nrow, k, w2[k * kHidden + nrow]); /* fc_2 kernel is (in, out) */
    for (int k = 0; k < kHidden; k++)
        for (int nrow = 0; nrow < width; nrow++)
            put(SM_B3, kOutPad, nrow, k, w3[k * width + nrow]);
    memcpy(&img[SM_W1 - SM_B2], w1, kObs * kHidden * 4);
    memcpy(&img[SM_B1 - SM_B2], b1, kHidden * 4);
    memcpy(&img[SM_BIAS2 - SM_B2], b2, kHidden * 4);
    memcpy(&img[SM_BIAS3 - SM_B2], b3, width * 4);

    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess)
        return q1_set_error(Q1_ENODEV, "cudaSetDevice failed: libq1phys has no CPU implementation");
    q1_policy *p = new (std::nothrow) q1_policy();
    if (!p)
        return q1_set_error(Q1_ENOMEM, "out of host memory");
    p->device = device;
    p->num_keys = num_keys;
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaError_t err = cudaMalloc(&p->image, kImageBytes);
    if (err == cudaSuccess)
        err = cudaMemcpy(p->image, img.data(), kImageBytes, cudaMemcpyHostToDevice);
    if (err == cudaSuccess)
        err = cudaFuncSetAttribute(k_policy_act, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
    if (prev >= 0)
        cudaSetDevice(prev);
    if (err != cudaSuccess) {
        if (p->image)
            cudaFree(p->image);
        delete p;
        return q1_set_error(Q1_ECUDA, std::string("q1_policy_create: ") + cudaGetErrorString(err));
    }
    *out = p;
    return Q1_OK;
}

int q1_policy_destroy(q1_policy *p)
{
    if (!p)
        return Q1_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(p->device);
    cudaFree(p->image);
    if (prev >= 0)
        cudaSetDevice(prev);
    delete p;
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
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(p->device) != cudaSuccess)
        return q1_set_error(Q1_ECUDA, "cudaSetDevice failed");
    const int64_t tiles = (n + kRows - 1) / kRows;
    const unsigned grid = (unsigned)(tiles < p->sm_count ? tiles : p->sm_count);
    k_policy_act<<<grid, kThreads, SM_TOTAL, static_cast<cudaStream_t>(stream)>>>(
        p->image, n, p->num_keys, obs, (float)action_low, (float)action_high, deterministic, seed, step,
        step_device, env_index_base, keys, mouse, logits_out);
    cudaError_t err = cudaGetLastError();
    if (prev >= 0)
        cudaSetDevice(prev);
    if (err != cudaSuccess)
        return q1_set_error(Q1_ECUDA, std::string("k_policy_act launch: ") + cudaGetErrorString(err));
    return Q1_OK;
}

} /* extern "C" */
