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
#include "dmma_tile.cuh"

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

// Cholesky + in-place triangular inverse of diagonal block p.  One CTA.
__global__ void __launch_bounds__(256, 1)
potrf_diag_kernel(double* __restrict__ A, int ld, int p, double* __restrict__ Linv, double* __restrict__ Wt, int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* B = reinterpret_cast<double*>(smem_raw);          // kFB x kDiagLd
    double* col = B + kFB * kDiagLd;                            // kFB: the column being eliminated by the inverse
    __shared__ int fail;
    const int tid = threadIdx.x;
    double* __restrict__ blk = A + (size_t)p * kFB * ld + (size_t)p * kFB;
    for (int e = tid; e < kFB * kFB; e += 256) B[(e / kFB) * kDiagLd + (e % kFB)] = blk[(size_t)(e / kFB) * ld + (e % kFB)];
    if (tid == 0) fail = 0;
    __syncthreads();
    // right-looking Cholesky, lower
    for (int j = 0; j < kFB; ++j) {
        if (tid == 0) {
            const double piv = B[j * kDiagLd + j];
            if (!(piv > 0.0)) fail = 1;
            B[j * kDiagLd + j] = sqrt(piv);
        }
        __syncthreads();
        if (fail) break;
        const double inv = 1.0 / B[j * kDiagLd + j];
        for (int i = j + 1 + tid; i < kFB; i += 256) B[i * kDiagLd + j] *= inv;
        __syncthreads();
        for (int i = j + 1 + (tid >> 4); i < kFB; i += 16) {      // 16 x 16 threads over (row i, column k <= i)
            const double lij = B[i * kDiagLd + j];
            for (int k = j + 1 + (tid & 15); k <= i; k += 16) B[i * kDiagLd + k] = fma(-lij, B[k * kDiagLd + j], B[i * kDiagLd + k]);
        }
        __syncthreads();
    }
    if (fail) {
        if (tid == 0 && atomicCAS(info, 0, 1 + p) == 0) {}
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
    // in-place inverse of the lower-triangular block, last column first (LAPACK dtrti2, lower):
    //   W[j][j] = 1 / L[j][j] ;  W[j+1:, j] = -W[j+1:, j+1:] L[j+1:, j] W[j][j]
    for (int j = kFB - 1; j >= 0; --j) {
        if (tid == 0) B[j * kDiagLd + j] = 1.0 / B[j * kDiagLd + j];
        for (int i = j + 1 + tid; i < kFB; i += 256) col[i] = B[i * kDiagLd + j];
        __syncthreads();
        const double wjj = B[j * kDiagLd + j];
        for (int i = j + 1 + tid; i < kFB; i += 256) {
            double acc = 0.0;
            for (int k = j + 1; k <= i; ++k) acc = fma(B[i * kDiagLd + k], col[k], acc);
            B[i * kDiagLd + j] = -acc * wjj;
        }
        __syncthreads();
    }
    for (int e = tid; e < kFB * kFB; e += 256) {
        const int i = e / kFB, j = e % kFB;
        const double w = j <= i ? B[i * kDiagLd + j] : 0.0;    // W[i][j]
        Linv[e] = w;
        Wt[((size_t)p * kFB + j) * ld + (size_t)p * kFB + i] = w;   // Wt = W^T: diagonal block of L^-T
    }
}

// acc fragment -> (row, column) of the 128 x 128 tile: acc[mi][ni][e] is element (row0 + mi*8 + lane/4, col0 + ni*8 + (lane%4)*2 + e)
template <class F>
__device__ __forceinline__ void for_each_acc(const double (&acc)[8][4][2], int tid, F&& f) {
    const int lane = tid & 31, warp = tid >> 5;
    const int row0 = (warp / 4) * 64, col0 = (warp % 4) * 32;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        const int r = row0 + mi * 8 + (lane >> 2);
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int c = col0 + ni * 8 + (lane & 3) * 2;
            f(r, c, acc[mi][ni][0], acc[mi][ni][1]);
        }
    }
}

#define CBO_FIT_TILE_PROLOGUE()                                                       \
    extern __shared__ __align__(16) unsigned char smem_raw[];                         \
    double* sA = reinterpret_cast<double*>(smem_raw);                                 \
    double* sB = sA + kFitStages * kFB * kBK;                                         \
    const int tid = threadIdx.x;                                                      \
    double acc[8][4][2];                                                              \
    _Pragma("unroll") for (int mi = 0; mi < 8; ++mi)                                  \
        _Pragma("unroll") for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

// A[I,p] <- A[I,p] Linv_p^T for I = p + 1 + blockIdx.x
__global__ void __launch_bounds__(256, 1)
trsm_panel_kernel(double* __restrict__ A, int ld, int p, const double* __restrict__ Linv) {
    CBO_FIT_TILE_PROLOGUE();
    const int I = p + 1 + blockIdx.x;
    double* __restrict__ blk = A + (size_t)I * kFB * ld + (size_t)p * kFB;
    abt_mainloop<2, 4, 8, 4, kFitStages>(blk, ld, Linv, kFB, kFB / kBK, sA, sB, acc, tid);
    for_each_acc(acc, tid, [&](int r, int c, double v0, double v1) {
        *reinterpret_cast<double2*>(blk + (size_t)r * ld + c) = make_double2(v0, v1);
    });
}

// A[I,J] -= A[I,p] A[J,p]^T for the lower block triangle behind panel p
__global__ void __launch_bounds__(256, 1)
syrk_update_kernel(double* __restrict__ A, int ld, int p) {
    CBO_FIT_TILE_PROLOGUE();
    int bi, bj;
    tri_tile(blockIdx.x, bi, bj);
    const int I = p + 1 + bi, J = p + 1 + bj;
    abt_mainloop<2, 4, 8, 4, kFitStages>(A + (size_t)I * kFB * ld + (size_t)p * kFB, ld, A + (size_t)J * kFB * ld + (size_t)p * kFB, ld,
                                         kFB / kBK, sA, sB, acc, tid);
    double* __restrict__ blk = A + (size_t)I * kFB * ld + (size_t)J * kFB;
    for_each_acc(acc, tid, [&](int r, int c, double v0, double v1) {
        double2* q = reinterpret_cast<double2*>(blk + (size_t)r * ld + c);
        double2 o = *q;
        o.x -= v0, o.y -= v1;
        *q = o;
    });
}

// Wt[J,K] <- Wt[J,K] Linv_K^T for J = blockIdx.x < K: block row K of W = L^-1 becomes final (held transposed)
__global__ void __launch_bounds__(256, 1)
winv_scale_kernel(double* __restrict__ Wt, int ld, int K, const double* __restrict__ Linv) {
    CBO_FIT_TILE_PROLOGUE();
    const int J = blockIdx.x;
    double* __restrict__ blk = Wt + (size_t)J * kFB * ld + (size_t)K * kFB;
    abt_mainloop<2, 4, 8, 4, kFitStages>(blk, ld, Linv, kFB, kFB / kBK, sA, sB, acc, tid);   // acc[n][m] = sum_k R^T[n][k] Linv[m][k]
    for_each_acc(acc, tid, [&](int r, int c, double v0, double v1) {
        *reinterpret_cast<double2*>(blk + (size_t)r * ld + c) = make_double2(v0, v1);
    });
}

// Wt[J,I] -= Wt[J,K] L[I,K]^T for I = K + 1 + blockIdx.x, J = blockIdx.y <= K: the running right-hand side of the rows below K
__global__ void __launch_bounds__(256, 1)
winv_update_kernel(const double* __restrict__ L, double* __restrict__ Wt, int ld, int K) {
    CBO_FIT_TILE_PROLOGUE();
    const int I = K + 1 + blockIdx.x, J = blockIdx.y;
    abt_mainloop<2, 4, 8, 4, kFitStages>(Wt + (size_t)J * kFB * ld + (size_t)K * kFB, ld, L + (size_t)I * kFB * ld + (size_t)K * kFB, ld,
                                         kFB / kBK, sA, sB, acc, tid);
    double* __restrict__ blk = Wt + (size_t)J * kFB * ld + (size_t)I * kFB;
    for_each_acc(acc, tid, [&](int r, int c, double v0, double v1) {
        double2* q = reinterpret_cast<double2*>(blk + (size_t)r * ld + c);
        double2 o = *q;
        o.x -= v0, o.y -= v1;
        *q = o;
    });
}

// Ky^-1[i][j] = sum_{k >= max(i,j)} Wt[i][k] Wt[j][k]; live N x N corner, row pitch n, both triangles
__global__ void __launch_bounds__(256, 1)
kyinv_kernel(const double* __restrict__ Wt, int ld, int n, double* __restrict__ kyinv) {
    CBO_FIT_TILE_PROLOGUE();
    int I, J;
    tri_tile(blockIdx.x, I, J);
    const int nb = ld / kFB;
    abt_mainloop<2, 4, 8, 4, kFitStages>(Wt + (size_t)I * kFB * ld + (size_t)I * kFB, ld, Wt + (size_t)J * kFB * ld + (size_t)I * kFB, ld,
                                         (nb - I) * (kFB / kBK), sA, sB, acc, tid);
    for_each_acc(acc, tid, [&](int r, int c, double v0, double v1) {
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
    constexpr size_t TILE_SMEM = (size_t)kFitStages * 2 * kFB * kBK * sizeof(double);
    constexpr size_t DIAG_SMEM = ((size_t)kFB * kDiagLd + kFB) * sizeof(double);
    CBO_CUDA(allow_dynamic_smem(potrf_diag_kernel, DIAG_SMEM));
    CBO_CUDA(allow_dynamic_smem(trsm_panel_kernel, TILE_SMEM));
    CBO_CUDA(allow_dynamic_smem(syrk_update_kernel, TILE_SMEM));
    CBO_CUDA(allow_dynamic_smem(winv_scale_kernel, TILE_SMEM));
    CBO_CUDA(allow_dynamic_smem(winv_update_kernel, TILE_SMEM));
    CBO_CUDA(allow_dynamic_smem(kyinv_kernel, TILE_SMEM));
    CBO_REQUIRE(d_info != nullptr, "cbo_obs_gp_fit: d_info is NULL");
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
        for (int p = 0; p < nb; ++p) {
            potrf_diag_kernel<<<1, 256, DIAG_SMEM, st>>>(A, npad, p, Linv + (size_t)p * kFB * kFB, Wt, info);
            note_launch();
            const int T = nb - p - 1;
            if (T > 0) {
                trsm_panel_kernel<<<T, 256, TILE_SMEM, st>>>(A, npad, p, Linv + (size_t)p * kFB * kFB);
                note_launch();
                syrk_update_kernel<<<T * (T + 1) / 2, 256, TILE_SMEM, st>>>(A, npad, p);
                note_launch();
            }
        }
        for (int K = 0; K + 1 < nb; ++K) {
            if (K > 0) {
                winv_scale_kernel<<<K, 256, TILE_SMEM, st>>>(Wt, npad, K, Linv + (size_t)K * kFB * kFB);
                note_launch();
            }
            winv_update_kernel<<<dim3(nb - K - 1, K + 1), 256, TILE_SMEM, st>>>(A, Wt, npad, K);
            note_launch();
        }
        if (nb > 1) {
            winv_scale_kernel<<<nb - 1, 256, TILE_SMEM, st>>>(Wt, npad, nb - 1, Linv + (size_t)(nb - 1) * kFB * kFB);
            note_launch();
        }
        kyinv_kernel<<<nb * (nb + 1) / 2, 256, TILE_SMEM, st>>>(Wt, npad, n, const_cast<double*>(S.kyinv));
        note_launch();
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
