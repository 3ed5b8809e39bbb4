// K5 -- exact-inference state of an observational GP on the device:  alpha = Ky^-1 y  and  Ky^-1.
//
// Replaces the GPRegression(...) inside fit_gaussian_process (reference utils.py:40-45; GPy ExactGaussianInference:
// Ky = K + (noise + 1e-8) I, L = jitchol(Ky), alpha = dpotrs(L, y), Ky^-1 = dpotri(L)) for FROZEN hyper-parameters;
// the hyper-parameter search around it stays on the host (it calls this once per evaluation at large N).
// Its outputs are the inputs of K1a (alpha_obs, kyinv): SURVEY.md §8(f).2.
//
// Everything O(N^3) runs on the FP64 tensor pipe through the 128 x 128 DMMA tile mainloop of dmma_tile.cuh
// (C += A B^T for two k-contiguous row operands), on an Npad x Npad workspace padded with an identity block:
//   gram      A = s2 exp(-.5 r^2) + (noise + 1e-8 + jitter) I           r^2 from coordinate differences; N^2 exp
//   potrf     right-looking blocked Cholesky, 128-column panels: per panel
//               diag   one CTA: Cholesky of the 128 x 128 diagonal block in shared memory, then its in-place
//                      triangular inverse Linv_p (also written, transposed, as the diagonal block of Wt)
//               trsm   A[I,p] <- A[I,p] Linv_p^T   for the row blocks below (a 128-deep tile product)
//               syrk   A[I,J] -= A[I,p] A[J,p]^T    for p < J <= I        (N^3/3 flops in total)
//   winv      Wt = L^-T (upper triangular, row-major): right-looking blocked solve of L W = I, carried out on the
//             transposes so that every operand is k-contiguous and every store lands in natural layout.  Wt starts as
//             zero with the diagonal blocks Linv_K^T; for K = 0 .. nb-1
//               scale   Wt[J,K] <- Wt[J,K] Linv_K^T            for J < K   (block row K of W is final)
//               update  Wt[J,I] -= Wt[J,K] L[I,K]^T            for J <= K < I   (128-deep tile products, (nb-K-1)(K+1)
//                                                                               independent tiles per step; N^3/3 flops)
//   kyinv     Ky^-1[i][j] = sum_{k >= max(i,j)} Wt[i][k] Wt[j][k], both triangles written (N^3/3 flops)
//   alpha     z = Wt^T y (partial sums over row chunks, added in order), alpha = Wt z     (two HBM-bound passes over Wt)
// A non-positive pivot is reported through `info` (1 + the panel index); the caller retries with GPy's jitter rule.
// Launch-latency bound for small N (3 launches per panel); at N = 1e4 the tile products dominate.
#include <stdlib.h>

#include "dmma_tma_tile.cuh"

namespace cbo {

constexpr int kFB = 128;              // block size of the factorisation = tile size of the DMMA mainloop
constexpr int kFitStages = 4;
constexpr int kDiagLd = kFB + 1;      // padded pitch of the diagonal block in shared memory (column walks hit all banks)

struct ObsX {
    const double* col[CBO_MAX_D + CBO_MAX_C];   // one device pointer per GP input column (n_obs doubles each)
    double il[CBO_MAX_D + CBO_MAX_C];           // 1 / lengthscale
    int D;
};

__global__ void __launch_bounds__(256)
gram_kernel(ObsX X, int n, int npad, double s2, double diag_add, double* __restrict__ A) {
    const int i = blockIdx.y;
    double xi[CBO_MAX_D + CBO_MAX_C];
#pragma unroll
    for (int k = 0; k < CBO_MAX_D + CBO_MAX_C; ++k) xi[k] = (k < X.D && i < n) ? X.col[k][i] : 0.0;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < npad; j += gridDim.x * 256) {
        double v;
        if (i < n && j < n) {
            double r2 = 0.0;
#pragma unroll
            for (int k = 0; k < CBO_MAX_D + CBO_MAX_C; ++k) {
                if (k < X.D) {
                    const double t = (xi[k] - X.col[k][j]) * X.il[k];
                    r2 = fma(t, t, r2);
                }
            }
            v = s2 * exp(-0.5 * r2) + (i == j ? diag_add : 0.0);
        } else {
            v = i == j ? 1.0 : 0.0;     // identity padding: the padded factor, inverse and Ky^-1 stay block diagonal
        }
        A[(size_t)i * npad + j] = v;
    }
}

// ---- diagonal block: Cholesky + triangular inverse of one 128 x 128 block, one CTA ---------------------------------------
// Blocked over 32 x 32 sub-blocks so that the serial part runs inside ONE warp on registers and shuffles (no CTA barrier
// per column: the column-by-column version spent 42 % of its 257 us in barrier stalls, profiles/r02_k5_potrf_diag_*):
//   phase A, for b = 0..3:   warp 0   Cholesky of sub-block (b, b): lane i owns row i, pivots and columns travel by shuffle
//                            warp 4   inverse of that sub-block (lane c owns column c of L_bb^-1: forward substitution)
//                            warps 1-3  rows below: X L_bb^T = A_rb by forward substitution, one thread per row
//                            all      trailing update B[i][k] -= sum_m X[i][m] X[k][m]   (register-tiled, 16 x 16 threads)
//   phase B:  W = L^-1 by block recursion  W21 = -W22 L21 W11  on 32- and then 64-wide halves (four small products).
constexpr int kSB = 32;
constexpr int kSubLd = kSB + 1;
constexpr int kHalfLd = 2 * kSB + 1;
constexpr size_t kDiagSmemDoubles = (size_t)kFB * kDiagLd + 4 * kSB * kSubLd + 2 * kSB * kHalfLd + kSB;

// acc[a][c] = sum_{k < K} A[(ty + 16 a)][k] * Bm[k][tx + 16 c]   (row-major shared-memory operands, 16 x 16 threads)
template <int MI, int NJ>
__device__ __forceinline__ void smem_product(double (&acc)[MI][NJ], const double* A, int lda, const double* Bm, int ldb, int K,
                                             int ty, int tx) {
#pragma unroll
    for (int a = 0; a < MI; ++a)
#pragma unroll
        for (int c = 0; c < NJ; ++c) acc[a][c] = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        double av[MI], bv[NJ];
#pragma unroll
        for (int a = 0; a < MI; ++a) av[a] = A[(ty + 16 * a) * lda + k];
#pragma unroll
        for (int c = 0; c < NJ; ++c) bv[c] = Bm[k * ldb + tx + 16 * c];
#pragma unroll
        for (int a = 0; a < MI; ++a)
#pragma unroll
            for (int c = 0; c < NJ; ++c) acc[a][c] = fma(av[a], bv[c], acc[a][c]);
    }
}

// B[i][k] -= sum_{m < 32} B[i][o + m] B[k][o + m]  for i, k in [o + 32, 128): NA = (96 - o) / 16 row / column groups per thread
template <int NA>
__device__ __forceinline__ void diag_trailing_update(double* B, int o, int ty, int tx) {
    double acc[NA][NA];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int c = 0; c < NA; ++c) acc[a][c] = 0.0;
    const double* ri = B + (o + kSB + ty) * kDiagLd + o;
    const double* rk = B + (o + kSB + tx) * kDiagLd + o;
#pragma unroll 4
    for (int m = 0; m < kSB; ++m) {
        double xi[NA], xk[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) xi[a] = ri[16 * a * kDiagLd + m];
#pragma unroll
        for (int c = 0; c < NA; ++c) xk[c] = rk[16 * c * kDiagLd + m];
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int c = 0; c < NA; ++c) acc[a][c] = fma(xi[a], xk[c], acc[a][c]);
    }
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int c = 0; c < NA; ++c) B[(o + kSB + ty + 16 * a) * kDiagLd + o + kSB + tx + 16 * c] -= acc[a][c];
}

__global__ void __launch_bounds__(256, 1)
potrf_diag_kernel(double* __restrict__ A, int ld, int p, double* __restrict__ Linv, double* __restrict__ Wt, int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* B = reinterpret_cast<double*>(smem_raw);          // kFB x kDiagLd: the block; L after phase A; W = L^-1 after phase B
    double* Wd = B + kFB * kDiagLd;                             // 4 x (32 x 33): inverses of the diagonal sub-blocks
    double* T = Wd + 4 * kSB * kSubLd;                          // 64 x 65: scratch product of phase B
    double* invd = T + 2 * kSB * kHalfLd;                       // 32: reciprocal diagonal of the current sub-block
    __shared__ int fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    double* __restrict__ blk = A + (size_t)p * kFB * ld + (size_t)p * kFB;
    for (int e = tid; e < kFB * kFB; e += 256) B[(e / kFB) * kDiagLd + (e % kFB)] = blk[(size_t)(e / kFB) * ld + (e % kFB)];
    if (tid == 0) fail = 0;
    __syncthreads();

    // ---- phase A
#pragma unroll 1
    for (int b = 0; b < kFB / kSB; ++b) {
        const int o = b * kSB;
        if (warp == 0) {
            double a[kSB];
            double* row = B + (o + lane) * kDiagLd + o;
#pragma unroll
            for (int k = 0; k < kSB; ++k) a[k] = row[k];
            bool bad = false;
#pragma unroll
            for (int j = 0; j < kSB; ++j) {
                const double piv = __shfl_sync(0xffffffffu, a[j], j);
                bad |= !(piv > 0.0);
                const double inv = rsqrt(piv);
                a[j] = lane == j ? piv * inv : a[j] * inv;        // lanes < j hold the (unused) upper part
                if (lane == j) invd[j] = inv;
#pragma unroll
                for (int k = j + 1; k < kSB; ++k) a[k] = fma(-a[j], __shfl_sync(0xffffffffu, a[j], k), a[k]);
            }
#pragma unroll
            for (int k = 0; k < kSB; ++k)
                if (k <= lane) row[k] = a[k];
            if (bad && lane == 0) fail = 1;
        }
        __syncthreads();
        if (warp == 4) {
            // column `lane` of L_bb^-1:  x[i] = (delta_{i,lane} - sum_{k < i} L[i][k] x[k]) / L[i][i]
            double x[kSB];
            const double* Lb = B + o * kDiagLd + o;
#pragma unroll
            for (int i = 0; i < kSB; ++i) {
                double s0 = i == lane ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k + 1 < i; k += 2) {
                    s0 = fma(-Lb[i * kDiagLd + k], x[k], s0);
                    s1 = fma(-Lb[i * kDiagLd + k + 1], x[k + 1], s1);
                }
                if (i & 1) s0 = fma(-Lb[i * kDiagLd + i - 1], x[i - 1], s0);
                x[i] = (s0 + s1) * invd[i];
                Wd[(b * kSB + i) * kSubLd + lane] = x[i];
            }
        } else if (warp >= 1 && warp <= 3) {
            const int r = o + kSB + (tid - 32);
            if (r < kFB) {
                double x[kSB];
                double* row = B + r * kDiagLd + o;
                const double* Lb = B + o * kDiagLd + o;
#pragma unroll
                for (int k = 0; k < kSB; ++k) {
                    double s0 = row[k], s1 = 0.0;
#pragma unroll
                    for (int m = 0; m + 1 < k; m += 2) {
                        s0 = fma(-x[m], Lb[k * kDiagLd + m], s0);
                        s1 = fma(-x[m + 1], Lb[k * kDiagLd + m + 1], s1);
                    }
                    if (k & 1) s0 = fma(-x[k - 1], Lb[k * kDiagLd + k - 1], s0);
                    x[k] = (s0 + s1) * invd[k];
                }
#pragma unroll
                for (int k = 0; k < kSB; ++k) row[k] = x[k];
            }
        }
        __syncthreads();
        if (b == 0) diag_trailing_update<6>(B, o, ty, tx);
        else if (b == 1) diag_trailing_update<4>(B, o, ty, tx);
        else if (b == 2) diag_trailing_update<2>(B, o, ty, tx);
        __syncthreads();
    }
    if (fail) {
        if (tid == 0) atomicCAS(info, 0, 1 + p);
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int e = tid; e < kFB * kFB; e += 256) {
            const int i = e / kFB, j = e % kFB;
            blk[(size_t)i * ld + j] = nan;
            Linv[e] = nan;
            Wt[((size_t)p * kFB + i) * ld + (size_t)p * kFB + j] = nan;
        }
        return;
    }
    for (int e = tid; e < kFB * kFB; e += 256) {   // L_pp (upper part zero) back to the workspace
        const int i = e / kFB, j = e % kFB;
        blk[(size_t)i * ld + j] = j <= i ? B[i * kDiagLd + j] : 0.0;
    }
    __syncthreads();

    // ---- phase B: B's diagonal sub-blocks <- Wd, the upper sub-blocks (0,1) and (2,3) <- 0, then the recursion
    for (int e = tid; e < 4 * kSB * kSB; e += 256) {
        const int b = e / (kSB * kSB), i = (e / kSB) % kSB, j = e % kSB;
        B[(b * kSB + i) * kDiagLd + b * kSB + j] = Wd[(b * kSB + i) * kSubLd + j];
        if (!(b & 1)) B[(b * kSB + i) * kDiagLd + (b + 1) * kSB + j] = 0.0;
    }
    {   // level 1: W[b1,b0] = -Wd[b1] (L[b1,b0] Wd[b0]) for (b0, b1) = (0, 1), (2, 3)
        double acc[2][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b0 = 2 * h, b1 = b0 + 1;
            smem_product<2, 2>(acc, B + (b1 * kSB) * kDiagLd + b0 * kSB, kDiagLd, Wd + b0 * kSB * kSubLd, kSubLd, kSB, ty, tx);
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int c = 0; c < 2; ++c) T[(h * kSB + ty + 16 * a) * kSubLd + tx + 16 * c] = acc[a][c];
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b0 = 2 * h, b1 = b0 + 1;
            smem_product<2, 2>(acc, Wd + b1 * kSB * kSubLd, kSubLd, T + h * kSB * kSubLd, kSubLd, kSB, ty, tx);
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int c = 0; c < 2; ++c) B[(b1 * kSB + ty + 16 * a) * kDiagLd + b0 * kSB + tx + 16 * c] = -acc[a][c];
        }
        __syncthreads();
    }
    {   // level 2: W21 = -W22 (L21 W11) on the 64-wide halves
        double acc[4][4];
        smem_product<4, 4>(acc, B + 2 * kSB * kDiagLd, kDiagLd, B, kDiagLd, 2 * kSB, ty, tx);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) T[(ty + 16 * a) * kHalfLd + tx + 16 * c] = acc[a][c];
        __syncthreads();
        smem_product<4, 4>(acc, B + 2 * kSB * kDiagLd + 2 * kSB, kDiagLd, T, kHalfLd, 2 * kSB, ty, tx);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) B[(2 * kSB + ty + 16 * a) * kDiagLd + tx + 16 * c] = -acc[a][c];
        __syncthreads();
    }
    for (int e = tid; e < kFB * kFB; e += 256) {
        const int i = e / kFB, j = e % kFB;
        const double w = j <= i ? B[i * kDiagLd + j] : 0.0;    // W[i][j]
        Linv[e] = w;
        Wt[((size_t)p * kFB + j) * ld + (size_t)p * kFB + i] = w;   // Wt = W^T: diagonal block of L^-T
    }
}

// ---- the one-wave tile products inside a panel ----------------------------------------------------------------------------
// Four updates run on at most nb - 1 <= 78 output tiles of 128 x 128 at a time, fewer than the GPU has SMs, and sit on the
// critical path between the diagonal-block kernels:
//   trsm      A[I,q]  <- A[I,q] Linv_q^T                       I > q                (the rows below the diagonal block)
//   colupdate A[I,q]  -= A[I,P] A[q,P]^T                        I >= q, P = p0..q-1  (left-looking, inside the panel)
//   scale     Wt[J,K] <- Wt[J,K] Linv_K^T                       J < K                (block row K of W = L^-1 becomes final)
//   wcolupdate Wt[J,K] -= Wt[J,P] L[K,P]^T                      J < K, P = max(J,K0)..K-1
// One kernel serves all four: tile t reads the rows a0 + t * a_step (pitch lda) against the rows b0 (pitch ldb) and writes or
// subtracts at c0 + t * c_step.  The 128 output rows of a tile are SPLIT over 128 / (16 MA) CTAs (MA = 8, 4, 2: 128, 64 or 32
// rows each) so that a wave of few tiles still covers the SMs; rows are independent, so the in-place forms stay race free.
struct PanelOp {
    const double* a0;
    const double* b0;
    double* c0;
    size_t a_step, c_step, lda, ldb, ldc;
    int nk;          // depth in 16-deep slabs
    int subtract;    // C -= product (else C = product)
    int tri_base;    // >= 0: tile t skips the leading max(t - tri_base, 0) blocks of the product (Wt is upper triangular)
};

template <int MA>
__global__ void __launch_bounds__(256, 1)
panel_tile_kernel(const PanelOp op) {
    constexpr int ROWS = 16 * MA, SPLIT = kFB / ROWS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sB = sA + kFitStages * ROWS * kBK;
    const int tid = threadIdx.x;
    const int t = blockIdx.x / SPLIT, sub = blockIdx.x % SPLIT;
    double acc[MA][4][2];
#pragma unroll
    for (int mi = 0; mi < MA; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const int skip = op.tri_base >= 0 && t > op.tri_base ? t - op.tri_base : 0;
    const double* gA = op.a0 + (size_t)t * op.a_step + (size_t)sub * ROWS * op.lda + (size_t)skip * kFB;
    const double* gB = op.b0 + (size_t)skip * kFB;
    abt_mainloop<2, 4, MA, 4, kFitStages>(gA, op.lda, gB, op.ldb, op.nk - skip * (kFB / kBK), sA, sB, acc, tid);
    double* __restrict__ blk = op.c0 + (size_t)t * op.c_step + (size_t)sub * ROWS * op.ldc;
    const int lane = tid & 31, warp = tid >> 5;
    const int row0 = (warp / 4) * MA * 8, col0 = (warp % 4) * 32;
#pragma unroll
    for (int mi = 0; mi < MA; ++mi) {
        const int r = row0 + mi * 8 + (lane >> 2);
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            double2* q = reinterpret_cast<double2*>(blk + (size_t)r * op.ldc + col0 + ni * 8 + (lane & 3) * 2);
            double2 o = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            if (op.subtract) {
                const double2 old = *q;
                o.x = old.x - o.x, o.y = old.y - o.y;
            }
            *q = o;
        }
    }
}

static cudaError_t launch_panel_op(const PanelOp& op, int tiles, int sms, cudaStream_t st) {
    constexpr size_t SMEM = (size_t)kFitStages * 2 * kFB * kBK * sizeof(double);
    if (tiles <= 0) return cudaSuccess;
    cudaError_t e;
    if (4 * tiles <= sms) {
        if ((e = allow_dynamic_smem(panel_tile_kernel<2>, SMEM)) != cudaSuccess) return e;
        panel_tile_kernel<2><<<4 * tiles, 256, SMEM, st>>>(op);
    } else if (2 * tiles <= sms) {
        if ((e = allow_dynamic_smem(panel_tile_kernel<4>, SMEM)) != cudaSuccess) return e;
        panel_tile_kernel<4><<<2 * tiles, 256, SMEM, st>>>(op);
    } else {
        if ((e = allow_dynamic_smem(panel_tile_kernel<8>, SMEM)) != cudaSuccess) return e;
        panel_tile_kernel<8><<<tiles, 256, SMEM, st>>>(op);
    }
    note_launch();
    return cudaGetLastError();
}

// ---- the three large tile products on the TMA pipeline (dmma_tma_tile.cuh) ------------------------------------------------
__device__ __forceinline__ void tile_subtract(double* __restrict__ blk, size_t ld, const double (&acc)[8][4][2], int tid) {
    tma_for_each_acc(acc, tid, [&](int r, int c, double v0, double v1) {
        double2* q = reinterpret_cast<double2*>(blk + (size_t)r * ld + c);
        double2 o = *q;
        o.x -= v0, o.y -= v1;
        *q = o;
    });
}

// A[I,J] -= A[I,P] A[J,P]^T, P = block columns p0 .. p0 + w - 1, lower block triangle behind the panel (T block rows)
struct SyrkPlan {
    double* A;
    int ld, p0, w, T;
    __device__ int count() const { return T * (T + 1) / 2; }
    __device__ TileJob job(int t) const {
        int bi, bj;
        tri_tile(t, bi, bj);
        return TileJob{(p0 + w + bi) * kFB, (p0 + w + bj) * kFB, p0 * kFB, w * (kFB / kBK), p0 + w + bi, p0 + w + bj};
    }
    __device__ void store(const TileJob& j, const double (&acc)[8][4][2], int tid) const {
        tile_subtract(A + (size_t)j.ti * kFB * ld + (size_t)j.tj * kFB, ld, acc, tid);
    }
};

// Wt[J,I] -= Wt[J,P] L[I,P]^T for the T block columns I behind the panel P = K0 .. K0 + w - 1 and every J < K0 + w; the product
// starts at block max(J, K0) (Wt is upper triangular).  Rows J of the panel itself (the short products) come last.
struct WinvPlan {
    double* Wt;
    int ld, K0, w, T;
    __device__ int count() const { return T * (K0 + w); }
    __device__ TileJob job(int t) const {
        const int J = t / T, I = K0 + w + t % T;
        const int k0 = J > K0 ? J : K0;
        return TileJob{J * kFB, I * kFB, k0 * kFB, (K0 + w - k0) * (kFB / kBK), J, I};
    }
    __device__ void store(const TileJob& j, const double (&acc)[8][4][2], int tid) const {
        tile_subtract(Wt + (size_t)j.ti * kFB * ld + (size_t)j.tj * kFB, ld, acc, tid);
    }
};

// Ky^-1[i][j] = sum_{k >= max(i,j)} Wt[i][k] Wt[j][k]; live N x N corner, row pitch n, both triangles.  Tiles in the order of
// tri_tile: block row I = 0 (the deepest products) first.
struct KyinvPlan {
    double* kyinv;
    int n, nb;
    __device__ int count() const { return nb * (nb + 1) / 2; }
    __device__ TileJob job(int t) const {
        int I, J;
        tri_tile(t, I, J);
        return TileJob{I * kFB, J * kFB, I * kFB, (nb - I) * (kFB / kBK), I, J};
    }
    __device__ void store(const TileJob& jb, const double (&acc)[8][4][2], int tid) const {
        const int I = jb.ti, J = jb.tj;
        tma_for_each_acc(acc, tid, [&](int r, int c, double v0, double v1) {
            const int i = I * kFB + r, j = J * kFB + c;
            if (i < n) {
                if (j < n) kyinv[(size_t)i * n + j] = v0;
                if (j + 1 < n) kyinv[(size_t)i * n + j + 1] = v1;
                if (I != J) {
                    if (j < n) kyinv[(size_t)j * n + i] = v0;
                    if (j + 1 < n) kyinv[(size_t)(j + 1) * n + i] = v1;
                }
            }
        });
    }
};

// z[k] = sum_{i <= k} Wt[i][k] y[i]: grid (column tile, row chunk); a thread owns one column of one chunk of 256 rows and
// writes one partial; wty_reduce_kernel adds the chunks in order (deterministic)
constexpr int kWtyRows = 256;
__global__ void __launch_bounds__(128)
wty_kernel(const double* __restrict__ Wt, int ld, int n, const double* __restrict__ y, double* __restrict__ part) {
    const int k = blockIdx.x * 128 + threadIdx.x, i0 = blockIdx.y * kWtyRows;
    double acc = 0.0;
    if (k < n) {
        const int top = k < i0 + kWtyRows - 1 ? k : i0 + kWtyRows - 1;   // rows i0 .. top (i <= k: Wt is upper triangular)
        for (int i = i0; i <= top; ++i) acc = fma(Wt[(size_t)i * ld + k], y[i], acc);
    }
    part[(size_t)blockIdx.y * ld + k] = acc;
}
__global__ void __launch_bounds__(128)
wty_reduce_kernel(const double* __restrict__ part, int ld, int chunks, double* __restrict__ z) {
    const int k = blockIdx.x * 128 + threadIdx.x;
    double acc = 0.0;
    for (int c = 0; c < chunks; ++c) acc += part[(size_t)c * ld + k];
    z[k] = acc;
}

// alpha[i] = sum_{k >= i} Wt[i][k] z[k]   (one warp per row, fixed lane assignment + butterfly: deterministic)
__global__ void __launch_bounds__(256)
wz_kernel(const double* __restrict__ Wt, int ld, int n, const double* __restrict__ z, double* __restrict__ alpha) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    double acc = 0.0;
    for (int k = (i & ~31) + lane; k < n; k += 32)
        if (k >= i) acc = fma(Wt[(size_t)i * ld + k], z[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) alpha[i] = acc;
}

// ---- negative log marginal likelihood and its gradient (the objective of gp.optimize(), utils.py:44) ---------------------
//   nll = 0.5 y.alpha + sum_i log L_ii + 0.5 N log(2 pi)
//   d nll / d log s2  = -0.5 sum_ij W_ij K_ij ,  d nll / d log l_k = -0.5 sum_ij W_ij K_ij ((x_ik - x_jk) / l_k)^2 ,
//   W = alpha alpha^T - Ky^-1 ,  K = s2 exp(-.5 r^2)  (the noise is fixed, utils.py:43)
// One CTA per row i recomputes K_ij on the fly and reduces its 1 + D sums in a fixed order; a second kernel adds the rows.
constexpr int kNllTerms = 1 + CBO_MAX_D + CBO_MAX_C;

__global__ void __launch_bounds__(256)
nll_rows_kernel(ObsX X, int n, double s2, const double* __restrict__ alpha, const double* __restrict__ kyinv, double* __restrict__ part) {
    __shared__ double red[8][kNllTerms];
    const int i = blockIdx.x, tid = threadIdx.x;
    double xi[CBO_MAX_D + CBO_MAX_C], acc[kNllTerms];
#pragma unroll
    for (int k = 0; k < CBO_MAX_D + CBO_MAX_C; ++k) xi[k] = k < X.D ? X.col[k][i] : 0.0;
#pragma unroll
    for (int t = 0; t < kNllTerms; ++t) acc[t] = 0.0;
    const double ai = alpha[i];
    for (int j = tid; j < n; j += 256) {
        double dk[CBO_MAX_D + CBO_MAX_C], r2 = 0.0;
#pragma unroll
        for (int k = 0; k < CBO_MAX_D + CBO_MAX_C; ++k) {
            dk[k] = 0.0;
            if (k < X.D) {
                const double t = (xi[k] - X.col[k][j]) * X.il[k];
                dk[k] = t * t;
                r2 += dk[k];
            }
        }
        const double wk = (ai * alpha[j] - kyinv[(size_t)i * n + j]) * (s2 * exp(-0.5 * r2));
        acc[0] += wk;
#pragma unroll
        for (int k = 0; k < CBO_MAX_D + CBO_MAX_C; ++k) acc[1 + k] = fma(wk, dk[k], acc[1 + k]);
    }
#pragma unroll
    for (int t = 0; t < kNllTerms; ++t) {
        const double v = warp_sum(acc[t]);
        if ((tid & 31) == 0) red[tid >> 5][t] = v;
    }
    __syncthreads();
    if (tid < kNllTerms) {
        double v = 0.0;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) v += red[wq][tid];
        part[(size_t)i * kNllTerms + tid] = v;
    }
}

// out[0] = nll, out[1] = d/dlog s2, out[2 + k] = d/dlog l_k; one CTA, fixed summation order
__global__ void __launch_bounds__(256)
nll_reduce_kernel(const double* __restrict__ part, int n, int D, const double* __restrict__ L, int ld, const double* __restrict__ y,
                  const double* __restrict__ alpha, double* __restrict__ out) {
    __shared__ double red[8];
    const int tid = threadIdx.x;
    for (int t = 0; t < 2 + kNllTerms; ++t) {          // t = 0: y.alpha, 1: sum log L_ii, 2..: the gradient sums
        double v = 0.0;
        for (int i = tid; i < n; i += 256)
            v += t == 0 ? y[i] * alpha[i] : (t == 1 ? log(L[(size_t)i * ld + i]) : part[(size_t)i * kNllTerms + (t - 2)]);
        v = warp_sum(v);
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int wq = 0; wq < 8; ++wq) tot += red[wq];
            if (t == 0) out[0] = 0.5 * tot + 0.5 * (double)n * 1.8378770664093454836;   // log(2 pi)
            else if (t == 1) out[0] += tot;
            else if (t - 2 < 1 + D) out[t - 1] = -0.5 * tot;
        }
        __syncthreads();
    }
}

// Panel width of the blocked factorisation, in 128-column blocks (CBO_FIT_PANEL overrides the default for tuning runs).
static int fit_panel_blocks() {
    static const int w = [] {
        const char* e = getenv("CBO_FIT_PANEL");
        const int v = e ? atoi(e) : 4;
        return v < 1 ? 1 : (v > 16 ? 16 : v);
    }();
    return w;
}

static size_t fit_ws_doubles(int npad) {
    const size_t nb = npad / kFB, chunks = (npad + kWtyRows - 1) / kWtyRows;
    const size_t zpart = (size_t)npad * chunks, nllpart = (size_t)npad * kNllTerms;        // the two never live together
    return 2 * (size_t)npad * npad + nb * kFB * kFB + (size_t)npad + (zpart > nllpart ? zpart : nllpart);   // A, Wt, Linv, z, partials
}

size_t obs_gp_workspace_bytes_impl(const cbo_set_desc* h_sets, int num_sets) {
    size_t m = 0;
    for (int s = 0; s < num_sets; ++s)
        if (computes_prior(h_sets[s]) && h_sets[s].y_obs) {
            const size_t b = fit_ws_doubles(h_sets[s].n_obs_pad) * sizeof(double);
            if (b > m) m = b;
        }
    return m;
}

int obs_gp_fit_impl(const cbo_set_desc* h_sets, int num_sets, double jitter, void* d_ws, size_t ws_bytes, int32_t* d_info,
                    cudaStream_t st) {
    constexpr size_t DIAG_SMEM = kDiagSmemDoubles * sizeof(double);
    CBO_CUDA(allow_dynamic_smem(potrf_diag_kernel, DIAG_SMEM));
    CBO_REQUIRE(d_info != nullptr, "cbo_obs_gp_fit: d_info is NULL");
    int dev = 0, sms = 0;
    CBO_CUDA(cudaGetDevice(&dev));
    CBO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    for (int s = 0; s < num_sets; ++s) {
        const cbo_set_desc& S = h_sets[s];
        if (!computes_prior(S) || !S.y_obs) continue;
        const int n = S.n_obs, npad = S.n_obs_pad, nb = npad / kFB;
        CBO_REQUIRE(S.alpha_obs && S.kyinv && S.x_obs_int && (S.c == 0 || S.x_obs_cond), "cbo_obs_gp_fit: set %d has a NULL pointer", s);
        CBO_REQUIRE(d_ws != nullptr && ws_bytes >= fit_ws_doubles(npad) * sizeof(double),
                    "cbo_obs_gp_fit: workspace of %zu bytes too small for set %d (%zu needed); see cbo_obs_gp_workspace_bytes",
                    ws_bytes, s, fit_ws_doubles(npad) * sizeof(double));
        double* A = reinterpret_cast<double*>(d_ws);
        double* Wt = A + (size_t)npad * npad;
        double* Linv = Wt + (size_t)npad * npad;
        double* z = Linv + (size_t)nb * kFB * kFB;
        double* zpart = z + npad;
        const int chunks = (npad + kWtyRows - 1) / kWtyRows;
        int* info = d_info + s;
        CBO_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
        CUtensorMap mapA, mapW;
        if (make_f64_rowmajor_map(&mapA, A, npad, npad, npad) || make_f64_rowmajor_map(&mapW, Wt, npad, npad, npad)) return -1;
        ObsX X;
        X.D = S.d + S.c;
        for (int k = 0; k < CBO_MAX_D + CBO_MAX_C; ++k) {
            X.col[k] = k < S.d ? S.x_obs_int + (size_t)k * n : (k < X.D ? S.x_obs_cond + (size_t)(k - S.d) * n : nullptr);
            X.il[k] = k < S.d ? 1.0 / S.ls_int[k] : (k < X.D ? 1.0 / S.ls_cond[k - S.d] : 0.0);
        }
        CBO_CUDA(cudaMemsetAsync(Wt, 0, (size_t)npad * npad * sizeof(double), st));   // right-hand side of L W = I, off-diagonal part
        gram_kernel<<<dim3((npad + 255) / 256 < 8 ? (npad + 255) / 256 : 8, npad), 256, 0, st>>>(X, n, npad, S.s2,
                                                                                               S.noise + 1e-8 + jitter, A);
        note_launch();
        // Cholesky: panels of W block columns.  Inside a panel each block column is first brought up to date against the panel's
        // earlier columns (left-looking, one wave of tiles), then factored; the trailing matrix takes ONE 128 W deep update per
        // panel, so its tiles run a W times longer mainloop per read-modify-write of C than a 128-deep right-looking update.
        const int W = fit_panel_blocks();
        for (int p0 = 0; p0 < nb; p0 += W) {
            const int we = nb - p0 < W ? nb - p0 : W;
            for (int q = p0; q < p0 + we; ++q) {
                const size_t ld = npad, step = (size_t)kFB * ld;
                double* Aq = A + (size_t)q * step;              // block row q
                if (q > p0)
                    CBO_CUDA(launch_panel_op(PanelOp{Aq + (size_t)p0 * kFB, Aq + (size_t)p0 * kFB, Aq + (size_t)q * kFB, step, step, ld, ld, ld,
                                                     (q - p0) * (kFB / kBK), 1, -1}, nb - q, sms, st));
                potrf_diag_kernel<<<1, 256, DIAG_SMEM, st>>>(A, npad, q, Linv + (size_t)q * kFB * kFB, Wt, info);
                note_launch();
                CBO_CUDA(launch_panel_op(PanelOp{Aq + step + (size_t)q * kFB, Linv + (size_t)q * kFB * kFB, Aq + step + (size_t)q * kFB, step, step,
                                                 ld, (size_t)kFB, ld, kFB / kBK, 0, -1}, nb - q - 1, sms, st));
            }
            const int T = nb - p0 - we;
            if (T > 0) CBO_CUDA(launch_tma_tiles(mapA, mapA, SyrkPlan{A, npad, p0, we, T}, T * (T + 1) / 2, sms, st));
        }
        // triangular inverse, same panel scheme on the transposes
        for (int K0 = 0; K0 < nb; K0 += W) {
            const int we = nb - K0 < W ? nb - K0 : W;
            for (int K = K0; K < K0 + we; ++K) {
                const size_t ld = npad, step = (size_t)kFB * ld;
                if (K > K0)
                    CBO_CUDA(launch_panel_op(PanelOp{Wt + (size_t)K0 * kFB, A + (size_t)K * step + (size_t)K0 * kFB, Wt + (size_t)K * kFB, step, step,
                                                     ld, ld, ld, (K - K0) * (kFB / kBK), 1, K0}, K, sms, st));
                CBO_CUDA(launch_panel_op(PanelOp{Wt + (size_t)K * kFB, Linv + (size_t)K * kFB * kFB, Wt + (size_t)K * kFB, step, step, ld,
                                                 (size_t)kFB, ld, kFB / kBK, 0, -1}, K, sms, st));
            }
            const int T = nb - K0 - we;
            if (T > 0) CBO_CUDA(launch_tma_tiles(mapW, mapA, WinvPlan{Wt, npad, K0, we, T}, T * (K0 + we), sms, st));
        }
        CBO_CUDA(launch_tma_tiles(mapW, mapW, KyinvPlan{const_cast<double*>(S.kyinv), n, nb}, nb * (nb + 1) / 2, sms, st));
        wty_kernel<<<dim3(npad / 128, chunks), 128, 0, st>>>(Wt, npad, n, S.y_obs, zpart);
        note_launch();
        wty_reduce_kernel<<<npad / 128, 128, 0, st>>>(zpart, npad, chunks, z);
        note_launch();
        wz_kernel<<<(n + 7) / 8, 256, 0, st>>>(Wt, npad, n, z, const_cast<double*>(S.alpha_obs));
        note_launch();
        CBO_CUDA(cudaGetLastError());
    }
    return 0;
}

// nll and gradient of ONE set from the state cbo_obs_gp_fit left behind: L in the workspace, alpha_obs / kyinv in the descriptor
int obs_gp_nll_impl(const cbo_set_desc* h_set, void* d_ws, size_t ws_bytes, double* d_out, cudaStream_t st) {
    const cbo_set_desc& S = *h_set;
    CBO_REQUIRE(computes_prior(S) && S.y_obs && S.alpha_obs && S.kyinv, "cbo_obs_gp_nll: the set has no device-fitted observational GP");
    const int n = S.n_obs, npad = S.n_obs_pad, nb = npad / kFB;
    CBO_REQUIRE(d_ws != nullptr && ws_bytes >= fit_ws_doubles(npad) * sizeof(double) && d_out != nullptr,
                "cbo_obs_gp_nll: needs the workspace of the preceding cbo_obs_gp_fit (%zu bytes) and an output buffer",
                fit_ws_doubles(npad) * sizeof(double));
    double* A = reinterpret_cast<double*>(d_ws);
    double* part = A + 2 * (size_t)npad * npad + (size_t)nb * kFB * kFB + npad;
    ObsX X;
    X.D = S.d + S.c;
    for (int k = 0; k < CBO_MAX_D + CBO_MAX_C; ++k) {
        X.col[k] = k < S.d ? S.x_obs_int + (size_t)k * n : (k < X.D ? S.x_obs_cond + (size_t)(k - S.d) * n : nullptr);
        X.il[k] = k < S.d ? 1.0 / S.ls_int[k] : (k < X.D ? 1.0 / S.ls_cond[k - S.d] : 0.0);
    }
    nll_rows_kernel<<<n, 256, 0, st>>>(X, n, S.s2, S.alpha_obs, S.kyinv, part);
    note_launch();
    nll_reduce_kernel<<<1, 256, 0, st>>>(part, n, X.D, A, npad, S.y_obs, S.alpha_obs, d_out);
    note_launch();
    CBO_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cbo
