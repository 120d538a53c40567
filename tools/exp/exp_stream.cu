// Streaming experiments behind DESIGN.md section 4: how fast can 123 MB per launch move through
// (a) a TMA bulk load -> shared memory -> TMA bulk store pipeline, (b) plain vectorised ld/st,
// out of place and in place, at the step kernel's launch size?  Standalone:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o exp_stream exp_stream.cu && ./exp_stream
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void bulk_load(uint32_t s, const void *g, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s), "l"(g), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void *g, uint32_t s, uint32_t bytes, uint64_t pol, int hint)
{
    if (hint)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(g), "r"(s), "r"(bytes), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(s), "r"(bytes) : "memory");
}

// One thread per CTA drives an S-stage ring of `tile` bytes: load tile k, store it, refill.
extern __shared__ __align__(128) unsigned char smem[];
__global__ void k_tma_copy(const char *src, char *dst, int64_t tiles, uint32_t tile, int stages, int hint)
{
    if (threadIdx.x != 0) return;
    uint32_t base = smem_addr(smem), bars = base + stages * tile;
    for (int s = 0; s < stages; s++) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint64_t pol = evict_first();
    int64_t t0 = blockIdx.x, g = gridDim.x;
    for (int s = 0; s < stages - 1; s++) {
        int64_t t = t0 + s * g;
        if (t < tiles) { mbar_expect(bars + 8 * s, tile); bulk_load(base + s * tile, src + t * tile, tile, bars + 8 * s); }
    }
    int s = 0; uint32_t parity = 0;
    for (int64_t t = t0; t < tiles; t += g) {
        mbar_wait(bars + 8 * s, parity);
        // the store issued one tile ago must have finished reading its stage before that stage is refilled
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        int64_t nt = t + (int64_t)(stages - 1) * g;
        int ps = s == 0 ? stages - 1 : s - 1;
        if (nt < tiles) { mbar_expect(bars + 8 * ps, tile); bulk_load(base + ps * tile, src + nt * tile, tile, bars + 8 * ps); }
        bulk_store(dst + t * tile, base + s * tile, tile, pol, hint);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (++s == stages) { s = 0; parity ^= 1; }
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// plain vectorised copy, grid-stride, 4 x 16 B in flight per thread
__global__ void k_ldst_copy(const uint4 *src, uint4 *dst, int64_t n16)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) dst[i] = src[i];
}

int main(int argc, char **argv)
{
    const int64_t n = argc > 1 ? atoll(argv[1]) : (1 << 20);
    const int64_t half = 117 * n / 2 / 24576 * 24576;      // bytes read = bytes written per launch
    const int ring = 4, steps = n <= (1 << 20) ? 3000 : 800;
    std::vector<char *> a(ring), b(ring);
    for (int r = 0; r < ring; r++) { CK(cudaMalloc(&a[r], half)); CK(cudaMalloc(&b[r], half)); CK(cudaMemset(a[r], 1, half)); CK(cudaMemset(b[r], 2, half)); }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(k_tma_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_tma_copy, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    auto timeit = [&](const char *name, auto launch) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            for (int i = 0; i < 50; i++) launch(i % ring);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int i = 0; i < steps; i++) launch(i % ring);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        CK(cudaGetLastError());
        printf("%-44s %8.2f us  %6.0f GB/s\n", name, best / steps * 1e3, 2.0 * half / (best / steps * 1e-3) / 1e9);
    };
    printf("n = %lld envs: %.1f MB read + %.1f MB written per launch\n", (long long)n, half / 1e6, half / 1e6);
    timeit("cudaMemcpyAsync D2D", [&](int r) { cudaMemcpyAsync(b[r], a[r], half, cudaMemcpyDeviceToDevice, 0); });
    for (int blocks : {sms * 8, sms * 16, sms * 32}) {
        char name[64]; snprintf(name, sizeof name, "ld/st copy, %d CTAs x 256", blocks);
        timeit(name, [&](int r) { k_ldst_copy<<<blocks, 256>>>((const uint4 *)a[r], (uint4 *)b[r], half / 16); });
    }
    struct Cfg { int tile, stages, ctas, hint, inplace; };
    const Cfg cfgs[] = {{6144, 3, 8, 1, 0}, {6144, 3, 16, 1, 0}, {6144, 4, 12, 1, 0}, {12288, 3, 6, 1, 0}, {12288, 4, 4, 1, 0},
                        {24576, 3, 3, 1, 0}, {24576, 4, 2, 1, 0}, {6144, 3, 8, 0, 0}, {12288, 3, 6, 0, 0},
                        {6144, 3, 8, 1, 1}, {12288, 3, 6, 1, 1}, {6144, 4, 12, 1, 1}};
    for (const Cfg &c : cfgs) {
        char name[96];
        snprintf(name, sizeof name, "TMA %s tile %5d x %d stages, %2d CTAs/SM%s", c.inplace ? "in-place " : "copy     ", c.tile, c.stages, c.ctas, c.hint ? "" : ", no hint");
        size_t sm = (size_t)c.tile * c.stages + 64;
        int64_t tiles = half / c.tile;
        timeit(name, [&](int r) { k_tma_copy<<<sms * c.ctas, 32, sm>>>(a[r], c.inplace ? a[r] : b[r], tiles, c.tile, c.stages, c.hint); });
    }
    return 0;
}
