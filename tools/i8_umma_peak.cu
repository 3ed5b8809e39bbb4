// Measured int8 tensor-core rate on sm_100a (tcgen05.mma kind::i8, S8 x S8 -> S32 in TMEM), the denominator of the "Ozaki
// scheme" question in DESIGN.md §8: could the FP64 quadratic form of K1b run faster as exact int8 slice products than on the
// DMMA pipe (36.97 TFLOP/s)?  One CTA per SM; one elected thread issues `iters` back-to-back 128 x N x 32 MMAs on operand
// tiles resident in shared memory (all ones, so every accumulator must read 32 * iters whatever the core-matrix layout),
// commits to an mbarrier, and the CTA checks the accumulators through tcgen05.ld.  Tiles: N = 256, 128, 64, 32 -- a slice
// scheme that keeps its 2 s - 1 equal-weight accumulator groups in TMEM (512 columns) is confined to narrow tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/i8_umma_peak tools/i8_umma_peak.cu && tools/i8_umma_peak
// Prints one JSON line.  Development probe: nothing in the library depends on it.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static __device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, no swizzle: core matrices of 8 rows x 16 bytes, `lbo` bytes between the core
// matrices along K, `sbo` bytes between 8-row groups (cute::UMMA::SmemDescriptor: address / offsets in 16-byte units,
// version 1 at bit 46, layout type 0 at bits 61-63)
static __device__ __forceinline__ uint64_t smem_desc(unsigned addr, unsigned lbo, unsigned sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(128, 1) i8_umma_kernel(int n_tile, int iters, int* __restrict__ bad, long long* __restrict__ clocks) {
    __shared__ __align__(128) uint8_t smA[128 * 32];
    __shared__ __align__(128) uint8_t smB[256 * 32];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 32; i += 128) smA[i] = 1;
    for (int i = tid; i < 256 * 32; i += 128) smB[i] = 1;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core's reads
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), K-major
    // operands, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_tile >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = smem_desc(smem_u32(smA), 128, 256), db = smem_desc(smem_u32(smB), 128, 256);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t acc = i > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everybody waits for the commit (bounded: a broken probe must not hang the box)
    unsigned ok = 0;
    const long long w0 = clock64();
    while (!ok && clock64() - w0 < (1ll << 33))          // ~4 s at 2 GHz
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    if (tid == 0) { t1 = clock64(); clocks[blockIdx.x] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) {
        if (tid == 0) atomicAdd(bad, 1 << 20);
    } else {
        // warp w reads TMEM lanes 32 w .. 32 w + 31 (thread t: lane 32 w + t), first and last column of the tile
        uint32_t v0, v1;
        const uint32_t a0 = tm + ((uint32_t)(32 * warp) << 16), a1 = a0 + (uint32_t)(n_tile - 1);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v0) : "r"(a0) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v1) : "r"(a1) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if ((int)v0 != 32 * iters || (int)v1 != 32 * iters) atomicAdd(bad, 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s\"}\n", cudaGetErrorString(e), #x); return 1; } } while (0)

int main() {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    int* d_bad; long long* d_clk;
    CK(cudaMalloc(&d_bad, sizeof(int)));
    CK(cudaMalloc(&d_clk, sizeof(long long) * sms));
    const int tiles[4] = {256, 128, 64, 32};
    printf("{\"sms\": %d, \"sm_clock_mhz\": %.0f, \"kind\": \"tcgen05.mma.cta_group::1.kind::i8, M = 128, K = 32, S8 x S8 -> S32\", \"tiles\": [", sms, khz / 1e3);
    for (int t = 0; t < 4; ++t) {
        const int n = tiles[t], iters = 100000;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e30f; int bad = 0;
        for (int rep = 0; rep < 4; ++rep) {            // first repetition is the warm-up
            CK(cudaMemset(d_bad, 0, sizeof(int)));
            CK(cudaEventRecord(e0));
            i8_umma_kernel<<<sms, 128>>>(n, iters, d_bad, d_clk);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
            int b; CK(cudaMemcpy(&b, d_bad, sizeof(int), cudaMemcpyDeviceToHost)); bad += b;
        }
        long long clk0; CK(cudaMemcpy(&clk0, d_clk, sizeof(long long), cudaMemcpyDeviceToHost));
        const double ops = 2.0 * 128 * n * 32 * (double)iters * sms;
        printf("%s{\"n\": %d, \"mmas_per_cta\": %d, \"ms\": %.4f, \"tera_ops\": %.1f, \"cycles_per_mma\": %.1f, \"accumulators_wrong\": %d}", t ? ", " : "", n, iters,
               best, ops / best * 1e-9, (double)clk0 / iters, bad);
    }
    printf("]}\n");
    return 0;
}
