// Shared device/host helpers for libcbo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/cbo_b200.h"

namespace cbo {

// ---- error plumbing (thread-local string, no exceptions across the ABI) ---------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define CBO_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) {                                 \
            ::cbo::set_error(__VA_ARGS__);             \
            return -1;                                 \
        }                                              \
    } while (0)

#define CBO_CUDA(call)                                                \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return ::cbo::cuda_fail(e__, #call);  \
    } while (0)

int validate_sets(const cbo_set_desc* h_sets, int num_sets);

// Every kernel launch of the library is tallied per calling thread (cbo_launch_count): a diagnostic that lets the host
// report how many of the library's kernels ran inside a timed region without re-deriving the launch logic.
void note_launch(int n = 1);

// does the library compute the causal prior of this set (as opposed to non-causal sets / caller-supplied priors)?
__host__ __device__ inline bool computes_prior(const cbo_set_desc& S) { return S.causal && !S.prior_external; }

// Work items of the batched sweep kernel.  A batched launch is a flat list of items (set 0 tiles, set 1 tiles, ...);
// host and device count them with the same function.
enum { kItemsSweep = 2 };
__host__ __device__ inline long long host_items(const cbo_set_desc& S, int /*kind*/) {
    return (S.g_count + CBO_SWEEP_TILE - 1) / CBO_SWEEP_TILE;
}

// K3 evaluates |L^-1 k*|^2 three ways, chosen by the set's own n_int (so a set's arithmetic never depends on which sets
// share the call): n <= kSweepFmaMaxN forward substitution in registers; n <= kSweepMmaMaxN a DMMA product with L^-1 (K2
// leaves L^-T in the strict upper triangle of `L` for these sets); beyond that forward substitution in shared memory.
constexpr int kSweepFmaMaxN = 16;
constexpr int kSweepMmaMaxN = 48;

// Opt a kernel into `bytes` of dynamic shared memory.  Called before every launch that needs more than 48 KB: the
// attribute is per device and per context, the call costs about a microsecond, and keeping no "already configured" flag
// keeps the library free of hidden state (several devices or threads in one process stay correct).
template <class K>
inline cudaError_t allow_dynamic_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// ---- device helpers --------------------------------------------------------------------------------
#ifdef __CUDACC__

// D(8x8) += A(8x4) * B(4x8), FP64 tensor pipe (SASS: DMMA.8x8x4).  Fragment ownership (lane = 4*g + t):
//   a = A[g][t]     b = B[t][g]     c0,c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// 16-byte read-only global load (SASS: LDG.E.128.CONSTANT).  Table pointers come out of a descriptor in
// memory, so the compiler cannot prove the state space on its own and would emit generic LD.E.
__device__ __forceinline__ double2 ldg_nc_d2(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(__cvta_generic_to_global(p)));
    return r;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- mbarrier (shared::cta) producer/consumer primitives ------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// one arrival + the number of bytes the async (TMA) proxy will deliver to this phase
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy (no tensor map): `bytes` contiguous global bytes -> shared, completion counted on `bar`
// (SASS: UBLKCP).  16-byte aligned addresses, size a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Order this thread's generic-proxy writes (st.global / st.shared) before later async-proxy (TMA) accesses.
__device__ __forceinline__ void fence_proxy_async() {
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
}
// arrive (without incrementing the pending count) once all prior cp.async of this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait suspends the warp in hardware for a bounded time, so this is not a hot spin.  A pipeline bug must
// not hang a shared GPU: after 2^24 failed probes (seconds) the kernel traps and the launch reports an error.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// (value, index) ordering used everywhere an argmax is reduced: larger value wins, ties go to the
// smaller index; NaN never wins (callers map NaN to -inf first).  Matches np.argmax (first maximum).
__device__ __forceinline__ bool better(double v, long long i, double bv, long long bi) {
    return (v > bv) || (v == bv && i < bi);
}

// Map a flat work-item id to (set, tile inside the set) by scanning the descriptors (<= a few dozen sets).
__device__ __forceinline__ int find_item(const cbo_set_desc* __restrict__ sets, int num_sets, int kind, int item,
                                         int& tile) {
    int base = 0, s = 0;
    for (; s < num_sets - 1; ++s) {
        const int c = (int)host_items(sets[s], kind);
        if (item < base + c) break;
        base += c;
    }
    tile = item - base;
    return s;
}

#endif  // __CUDACC__

}  // namespace cbo
