// FP64 pipe microbenchmark for B200 (sm_100a): establishes the roofline denominator for the
// causal-prior quadratic form ("FP64 peak must be measured first thing on the box").
// Three issue styles are timed with CUDA events, each alone on the device:
//   dfma      : independent DFMA chains, 8 per thread                 (vector FP64 pipe)
//   dmma884   : mma.sync.m8n8k4.f64   , 8 independent accumulators   (legacy DMMA shape, sm_80+)
//   dmma16816 : mma.sync.m16n8k16.f64 , 4 independent accumulators   (sm_90+ shape)
// Output: one JSON object on stdout.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;  // keep the chains alive
}

__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double seed) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = seed; c[i][1] = seed + i; }
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double seed) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = seed + i + j;
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (i + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                    "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                    : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                      "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    const int threads = 256;
    const int blocks_per_sm[3] = {1, 2, 4};
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [", p.name, sms);
    bool first = true;
    for (int bi = 0; bi < 3; ++bi) {
        int grid = sms * blocks_per_sm[bi] * 4;  // 4 waves
        int warps = grid * threads / 32;
        {
            int iters = 4096;
            double ms = time_ms([&] { k_dfma<<<grid, threads>>>(out, iters, 1.0); }, 10);
            double flops = 2.0 * 64.0 * iters * (double)grid * threads;
            printf("%s{\"kind\": \"dfma\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}", first ? "" : ", ",
                   blocks_per_sm[bi], ms, flops / ms * 1e-9);
            first = false;
        }
        {
            int iters = 1024;
            double ms = time_ms([&] { k_dmma884<<<grid, threads>>>(out, iters, 1.0); }, 10);
            double flops = 2.0 * 256.0 * 32.0 * iters * (double)warps;
            printf(", {\"kind\": \"dmma884\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}",
                   blocks_per_sm[bi], ms, flops / ms * 1e-9);
        }
        {
            int iters = 512;
            double ms = time_ms([&] { k_dmma16816<<<grid, threads>>>(out, iters, 1.0); }, 10);
            double flops = 2.0 * 2048.0 * 8.0 * iters * (double)warps;
            printf(", {\"kind\": \"dmma16816\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}",
                   blocks_per_sm[bi], ms, flops / ms * 1e-9);
        }
    }
    // sustained (about 3 s each) for the two best styles at 2 CTAs/SM
    {
        int grid = sms * 8;
        int warps = grid * threads / 32;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int kind = 0; kind < 2; ++kind) {
            int n = 0; float ms = 0;
            CK(cudaEventRecord(e0));
            do {
                for (int r = 0; r < 20; ++r) {
                    if (kind == 0) k_dfma<<<grid, threads>>>(out, 4096, 1.0);
                    else k_dmma884<<<grid, threads>>>(out, 1024, 1.0);
                }
                n += 20;
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
            } while (ms < 3000.f);
            double flops_per = kind == 0 ? 2.0 * 64.0 * 4096 * (double)grid * threads
                                         : 2.0 * 256.0 * 32.0 * 1024 * (double)warps;
            printf(", {\"kind\": \"%s_sustained\", \"seconds\": %.2f, \"tflops\": %.3f}",
                   kind == 0 ? "dfma" : "dmma884", ms * 1e-3, flops_per * n / ms * 1e-9);
        }
    }
    printf("]}\n");
    return 0;
}
