// Ceiling probe for k_emit's output pattern: every CTA bulk-stores two unaligned ~11.9 KB tiles from shared memory
// (cp.async.bulk shared -> global), optionally after gathering 64 random 128-byte windows and after a delay.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tile_store tile_store.cu && ./tile_store
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 8) k_store(uint8_t *out1, uint8_t *out2, uint32_t tile_bytes, const uint32_t *hap, uint64_t hap_words,
                                                  int gather, int delay_cycles, int fill, const uint8_t *qfill, uint32_t stage_cap) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(8) unsigned long long mbar;
    const uint32_t sm = (uint32_t)__cvta_generic_to_shared(smem), mb = (uint32_t)__cvta_generic_to_shared(&mbar);
    const uint64_t g = (uint64_t)blockIdx.x * tile_bytes;  // multiple of 16 when tile_bytes is
    if (fill && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(2u * stage_cap) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm), "l"(qfill), "r"(2u * stage_cap), "r"(mb) : "memory");
    }
    uint32_t acc = 0;
    if (gather) {  // 64 windows at random places, half a warp each (as the sequence phase loads them)
        // gather: 1 = 64 independent windows of 16 words; 2 = 32 places, two windows 50 words apart (a read pair);
        //         3 = as 1 but only 11 words per window; 4 = as 1 inside the first 32 MB (L2 resident); 5 = as 2 with 11 words
        uint64_t x = (uint64_t)blockIdx.x * 0x9E3779B97F4A7C15ull + 12345;
        const uint64_t span = gather == 4 ? (32ull << 20) / 4 : hap_words - 128;
        const uint32_t lanes = (gather == 3 || gather == 5) ? 11u : 16u;
        for (int r = threadIdx.x >> 4; r < 64; r += 8) {
            const int place = (gather == 2 || gather == 5) ? (r >> 1) : r;
            uint64_t y = x + (uint64_t)place * 0xD1B54A32D192ED03ull;
            y ^= y >> 29; y *= 0xBF58476D1CE4E5B9ull; y ^= y >> 32;
            const uint64_t w = (y % span) + (((gather == 2 || gather == 5) && (r & 1)) ? 50u : 0u);
            if ((threadIdx.x & 15) < lanes) acc += hap[w + (threadIdx.x & 15)];
        }
    }
    if (delay_cycles) {
        const long long t0 = clock64();
        while (clock64() - t0 < delay_cycles) {}
    }
    __syncthreads();
    if (fill) {
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(mb) : "memory");
    }
    if (gather && acc == 0xdeadbeef) smem[threadIdx.x] = 1;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out1 + g), "r"(sm), "r"(tile_bytes) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out2 + g), "r"(sm + stage_cap), "r"(tile_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// persistent CTAs with two stage sets: delay (stands for building a tile), then a committed, un-awaited bulk store of the
// set; a set is waited for (wait_group.read 1) only before it is reused
__global__ void __launch_bounds__(256, 4) k_store_p(uint8_t *out1, uint8_t *out2, uint32_t tile_bytes, uint32_t n_tiles, int delay_cycles, uint32_t stage_cap,
                                                    int jitter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t it = 0;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const uint32_t sm = sm0 + (it & 1u) * 2u * stage_cap;
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        int d = delay_cycles;
        if (jitter) d = (int)(((t * 2654435761u) >> 16) % (2u * (uint32_t)delay_cycles + 1u));
        if (d) {
            const long long t0 = clock64();
            while (clock64() - t0 < d) {}
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint64_t g = (uint64_t)t * tile_bytes;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out1 + g), "r"(sm), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out2 + g), "r"(sm + stage_cap), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// one tile per CTA, delay drawn per tile (0 .. 2 * delay): do equal CTAs move in lock-step?
__global__ void __launch_bounds__(128, 8) k_store_j(uint8_t *out1, uint8_t *out2, uint32_t tile_bytes, int delay_cycles, uint32_t stage_cap, int jitter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t sm = (uint32_t)__cvta_generic_to_shared(smem);
    int d = delay_cycles;
    if (jitter) d = (int)(((blockIdx.x * 2654435761u) >> 16) % (2u * (uint32_t)delay_cycles + 1u));
    const long long t0 = clock64();
    while (clock64() - t0 < d) {}
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t g = (uint64_t)blockIdx.x * tile_bytes;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out1 + g), "r"(sm), "r"(tile_bytes) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out2 + g), "r"(sm + stage_cap), "r"(tile_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

__global__ void k_fill64(uint64_t *p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0x4949494949494949ull;
}

int main() {
    const uint32_t tile_bytes = 11904, stage_cap = 12512;  // 32 records of 372 bytes
    const uint32_t n_tiles = 183864;                        // one chunk of the C5 bench
    uint8_t *o1, *o2, *qf;
    uint32_t *hap;
    const uint64_t hap_words = 1500000000ull / 4;
    cudaMalloc(&o1, (size_t)n_tiles * tile_bytes + 256);
    cudaMalloc(&o2, (size_t)n_tiles * tile_bytes + 256);
    cudaMalloc(&qf, 2 * stage_cap);
    cudaMalloc(&hap, hap_words * 4);
    cudaMemset(qf, 'I', 2 * stage_cap);
    cudaMemset(hap, 1, hap_words * 4);
    cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const double bytes = 2.0 * n_tiles * tile_bytes;
    struct { const char *name; int gather, delay, fill; } cases[] = {
        {"store only", 0, 0, 0}, {"store + tma fill", 0, 0, 1}, {"store + gather 64 x 16 words", 1, 0, 0},
        {"store + gather 32 pairs x 2 x 16 words", 2, 0, 0}, {"store + gather 64 x 11 words", 3, 0, 0},
        {"store + gather 32 pairs x 2 x 11 words", 5, 0, 0}, {"store + gather 64 x 16 words in 32 MB", 4, 0, 0},
        {"store + gather + fill", 1, 0, 1}, {"store + gather + fill + 4000 cyc", 1, 4000, 1}};
    for (auto &c : cases) {
        float best = 1e9f;
        for (int it = 0; it < 6; ++it) {
            cudaEventRecord(a);
            k_store<<<n_tiles, 128, 2 * stage_cap>>>(o1, o2, tile_bytes, hap, hap_words, c.gather, c.delay, c.fill, qf, stage_cap);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (it >= 2 && ms < best) best = ms;
        }
        printf("%-42s %.3f ms  %.0f GB/s of text\n", c.name, best, bytes / best / 1e6);
    }
    cudaFuncSetAttribute(k_store_p, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k_store_j, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int jitter = 0; jitter < 2; ++jitter)
        for (int delay : {4000, 8000}) {
            float best = 1e9f;
            for (int it = 0; it < 5; ++it) {
                cudaEventRecord(a);
                k_store_j<<<n_tiles, 128, 2 * stage_cap>>>(o1, o2, tile_bytes, delay, stage_cap, jitter);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                if (it >= 2 && ms < best) best = ms;
            }
            printf("one tile per CTA, delay %5d%s      %.3f ms  %.0f GB/s (delay alone: %.3f ms)\n", delay, jitter ? " +-100%" : "       ", best, bytes / best / 1e6,
                   delay / 1.965e6 * n_tiles / 1184.0);
        }
    for (int jitter = 0; jitter < 2; ++jitter)
        for (int delay : {0, 2000, 4000, 8000}) {
            float best = 1e9f;
            for (int it = 0; it < 5; ++it) {
                cudaEventRecord(a);
                k_store_p<<<148 * 4, 256, 4 * stage_cap>>>(o1, o2, tile_bytes, n_tiles, delay, stage_cap, jitter);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                if (it >= 2 && ms < best) best = ms;
            }
            printf("persistent 2 sets, delay %5d%s      %.3f ms  %.0f GB/s (delay alone: %.3f ms)\n", delay, jitter ? " +-100%" : "       ", best, bytes / best / 1e6,
                   delay / 1.965e6 * n_tiles / 592.0);
        }
    {
        float best = 1e9f;
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(a);
            k_fill64<<<148 * 16, 256>>>((uint64_t *)o1, (size_t)n_tiles * tile_bytes / 8);
            k_fill64<<<148 * 16, 256>>>((uint64_t *)o2, (size_t)n_tiles * tile_bytes / 8);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (it >= 1 && ms < best) best = ms;
        }
        printf("%-42s %.3f ms  %.0f GB/s\n", "plain 64-bit fill of the same bytes", best, bytes / best / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
