// FP64 tensor-pipe (DMMA.8x8x4) tile machinery shared by the prior-precompute SYRK (K1a) and the
// prior quadratic form (K1b).
//
// Both operands of C[m][n] += sum_k A[m][k] * B[n][k] are "k-contiguous row" matrices.  A BK=16 deep
// slab of ROWS rows is kept in shared memory in FRAGMENT ORDER  [kb = k/4][row][k%4]  so that the
// fragment of one 8-row block for one k4-step is 32 consecutive doubles: every warp-wide LDS.64 is
// conflict-free (2 wavefronts for 256 B, the minimum), and a global row segment of 4 doubles lands as one
// 32-byte unit -- cp.async 16-byte chunks map straight onto it without a transposing pass.
#pragma once
#include "cbo_common.cuh"

namespace cbo {

constexpr int kBK = 16;  // k depth of one pipeline stage

__device__ __forceinline__ int frag_off(int rows, int kb, int r, int kk) { return ((kb * rows + r) << 2) + kk; }

// Blocked global layout of the prior matrix M (written by K1a, streamed by K1b): the 128-row x 16-column slab
// (row block jb, k block kt) is stored contiguously and already in fragment order, so one pipeline stage of
// K1b is a single contiguous 16 KB TMA bulk copy.  Element (n, k) of the Npad x Npad matrix:
constexpr int kMBlkRows = 128;
constexpr int kMBlkDoubles = kMBlkRows * kBK;
__host__ __device__ __forceinline__ size_t mblk_base(int jb, int kt, int npad) {
    return ((size_t)jb * (npad / kBK) + kt) * kMBlkDoubles;
}
__host__ __device__ __forceinline__ size_t mblk_off(int n, int k, int npad) {
    return mblk_base(n / kMBlkRows, k / kBK, npad) + ((((k % kBK) >> 2) * kMBlkRows + (n % kMBlkRows)) << 2) + (k & 3);
}

// Asynchronously copy ROWS x 16 doubles (global row pitch `ld`, 16-byte aligned) into a fragment-order slab.
template <int ROWS, int NT>
__device__ __forceinline__ void load_rows_async(double* slab, const double* __restrict__ g, size_t ld, int tid) {
    constexpr int CH = kBK / 2;  // 16-byte chunks per row
    static_assert((ROWS * CH) % NT == 0, "slab must divide evenly over the CTA");
#pragma unroll
    // 128-bit shared accesses are served per quarter-warp: 8 consecutive lanes take the same k4-group of four
    // consecutive rows (128 contiguous bytes of the slab -> all 32 banks), the warp still reads 4 full 128-byte lines.
    for (int it = 0; it < ROWS * CH / NT; ++it) {
        const int idx = tid + it * NT;
        const int l = idx & 31;
        const int r = (idx >> 5) * 4 + ((l & 7) >> 1), ch = (l >> 3) * 2 + (l & 1);
        cp_async16(slab + frag_off(ROWS, ch >> 1, r, (ch & 1) * 2), g + (size_t)r * ld + ch * 2);
    }
}

// One pipeline stage of the warp-tiled product: warp (wm, wn) owns MA x NB blocks of 8x8.
// LIVE = how many of the warp's MA row blocks hold live rows (compile time, so that dead blocks cost neither issue slots
// nor pipe time -- predicated-off DMMAs still flow through the pipe).  LIVE == MA is the full tile; partially filled tiles
// (the last tile of a grid slice, the few interventional rows) use the smaller instantiations.
template <int BM, int BN, int MA, int NB, int LIVE = MA>
__device__ __forceinline__ void mma_stage(const double* __restrict__ sA, const double* __restrict__ sB,
                                          double (&acc)[MA][NB][2], int row0, int col0, int lane) {
    static_assert(LIVE >= 1 && LIVE <= MA, "LIVE row blocks");
#pragma unroll
    for (int kb = 0; kb < kBK / 4; ++kb) {
        double a[LIVE], b[NB];
#pragma unroll
        for (int mi = 0; mi < LIVE; ++mi) a[mi] = sA[((kb * BM + row0 + mi * 8) << 2) + lane];
#pragma unroll
        for (int ni = 0; ni < NB; ++ni) b[ni] = sB[((kb * BN + col0 + ni * 8) << 2) + lane];
#pragma unroll
        for (int mi = 0; mi < LIVE; ++mi)
#pragma unroll
            for (int ni = 0; ni < NB; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
}

// Mainloop of a warp-tiled 128 x 128 product  acc += sum over nk 16-deep slabs of A[m][k] * B[n][k]  for two
// "k-contiguous row" operands in global memory (row pitches lda / ldb in doubles, 16-byte aligned rows): a STAGES-deep
// cp.async ring of fragment-order slabs, one __syncthreads per slab.  Used by the prior-precompute SYRK (K1a) and by the
// blocked Cholesky / triangular inverse / Ky^-1 kernels of the observational GP fit (K5).  On return every cp.async has
// landed and every warp has passed a barrier: the caller may reuse the shared memory and may overwrite the operands.
template <int WM, int WN, int MA, int NB, int STAGES>
__device__ __forceinline__ void abt_mainloop(const double* __restrict__ gA, size_t lda, const double* __restrict__ gB, size_t ldb,
                                             int nk, double* sA, double* sB, double (&acc)[MA][NB][2], int tid) {
    constexpr int BM = WM * MA * 8, BN = WN * NB * 8, NT = WM * WN * 32;
    constexpr int TA = BM * kBK, TB = BN * kBK;
    const int lane = tid & 31, warp = tid >> 5;
    const int row0 = (warp / WN) * MA * 8, col0 = (warp % WN) * NB * 8;
#pragma unroll 1
    for (int i = 0; i < STAGES - 1; ++i) {
        if (i < nk) {
            load_rows_async<BM, NT>(sA + i * TA, gA + (size_t)i * kBK, lda, tid);
            load_rows_async<BN, NT>(sB + i * TB, gB + (size_t)i * kBK, ldb, tid);
        }
        cp_async_commit();
    }
#pragma unroll 1
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nx = kt + STAGES - 1;
        if (nx < nk) {
            const int ps = nx % STAGES;
            load_rows_async<BM, NT>(sA + ps * TA, gA + (size_t)nx * kBK, lda, tid);
            load_rows_async<BN, NT>(sB + ps * TB, gB + (size_t)nx * kBK, ldb, tid);
        }
        cp_async_commit();
        const int cs = kt % STAGES;
        mma_stage<BM, BN, MA, NB>(sA + cs * TA, sB + cs * TB, acc, row0, col0, lane);
    }
    cp_async_wait<0>();
    __syncthreads();
}

// lower-triangular tile pair (bi >= bj) from a linear index t = bi (bi + 1) / 2 + bj
__device__ __forceinline__ void tri_tile(int t, int& bi, int& bj) {
    bi = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long long)bi * (bi + 1) / 2 > t) --bi;
    while ((long long)(bi + 1) * (bi + 2) / 2 <= t) ++bi;
    bj = t - bi * (bi + 1) / 2;
}

}  // namespace cbo
