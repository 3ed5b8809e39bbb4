// K1b -- causal prior on a tensor grid:  m(x) = u(x).w ,  v(x) = s2 + noise - u(x)^T M u(x).
//
// Replaces DoCalculus.update_do_function / compute_do (reference DoCalculus.py:34-89): the reference
// rebuilds an (N x D) design per candidate and runs a full GP predict (O(N^2 D + N^3)); here the
// Monte-Carlo average has been folded into w and M once per observation (K1a) and each candidate costs a
// dense symmetric quadratic form in its N-vector u(x) = prod_k tab_k[i_k(x)][:].
//
// Shape of the work: T = U (G x N) * M (N x N) followed by a row-wise dot with U.  Only M's lower
// block triangle is visited: for column block J the k-blocks below the diagonal block are accumulated,
// doubled, and the (full, symmetric) diagonal block is added.  U is never materialised: a 128 x 16 slab is
// regenerated per pipeline stage from the L2-resident exp tables (d loads + (d-1) multiplies per
// element), M streams through a cp.async ring, the product runs on the FP64 tensor pipe (DMMA.8x8x4).
// Roofline: FP64 pipe.  Executed flops per candidate = N^2 (+ lower order); the dense-counted figure of
// SURVEY.md §8(d) is 2 N^2 + 2 N + d N.
#include "dmma_tile.cuh"

namespace cbo {

template <int WM_, int WN_, int MA_, int NB_, int STAGES_>
struct PriorCfg {
    static constexpr int WM = WM_, WN = WN_, MA = MA_, NB = NB_, STAGES = STAGES_;
    static constexpr int BM = WM * MA * 8;  // grid points per CTA
    static constexpr int BN = WN * NB * 8;  // columns of M per J block
    static constexpr int NT = WM * WN * 32;
    static constexpr int A_TILE = BM * kBK;
    static constexpr int B_TILE = BN * kBK;
    static constexpr int KPER = BM * kBK / NT;  // U elements generated per thread per stage
    static constexpr int TPR = kBK / KPER;      // threads per U row
    static constexpr size_t SMEM =
        (size_t)STAGES * (A_TILE + B_TILE) * sizeof(double) + 2 * WN * BM * sizeof(double) + CBO_MAX_D * BM * sizeof(int32_t);
    static_assert(BM == CBO_PRIOR_TILE, "host item count assumes CBO_PRIOR_TILE points per CTA");
    static_assert(KPER >= 2 && KPER % 2 == 0 && kBK % KPER == 0, "U generation mapping");
    static_assert(CBO_NPAD % BN == 0, "n_obs_pad must be a whole number of J blocks");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, 1)
prior_eval_kernel(const cbo_set_desc* __restrict__ sets, int num_sets, int which) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, NT = Cfg::NT, MA = Cfg::MA, NB = Cfg::NB, WN = Cfg::WN;
    constexpr int STAGES = Cfg::STAGES, KPER = Cfg::KPER, TPR = Cfg::TPR;
    constexpr int KB_PER_J = BN / kBK;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sB = sA + STAGES * Cfg::A_TILE;
    double* sRed = sB + STAGES * Cfg::B_TILE;                       // [2][WN][BM]
    int32_t* sRow = reinterpret_cast<int32_t*>(sRed + 2 * WN * BM);  // [CBO_MAX_D][BM], -1 = row outside the slice

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int row0 = wm * MA * 8, col0 = wn * NB * 8;

    int tile;
    const int s = find_item(sets, num_sets, which == 0 ? kItemsPriorGrid : kItemsPriorTrain, blockIdx.x, tile);
    const cbo_set_desc& S = sets[s];

    // effective problem: the tensor grid (which == 0) or the n_int interventional rows (which == 1)
    int d;
    const double* tab[CBO_MAX_D];
    int p[CBO_MAX_D];
    long long gbeg, gcnt;
    double *out_m, *out_v;
    if (which == 0) {
        d = S.d;
#pragma unroll
        for (int k = 0; k < CBO_MAX_D; ++k) { tab[k] = S.tab[k]; p[k] = S.p[k]; }
        gbeg = S.g_begin; gcnt = S.g_count; out_m = S.m; out_v = S.v;
    } else {
        d = 1;
#pragma unroll
        for (int k = 0; k < CBO_MAX_D; ++k) { tab[k] = S.u_int; p[k] = S.n_int; }
        gbeg = 0; gcnt = S.n_int; out_m = S.m_int; out_v = S.v_int;
    }
    const int Npad = S.n_obs_pad;
    const double* __restrict__ M = S.M;
    const double* __restrict__ w = S.w;
    const int nJ = (S.n_obs + BN - 1) / BN;

    // per-row table offsets (C-order decomposition of the flat grid index, last dim fastest)
    for (int r = tid; r < BM; r += NT) {
        const long long loc = (long long)tile * BM + r;
        if (loc < gcnt) {
            long long gg = gbeg + loc;
            for (int k = d - 1; k >= 0; --k) {
                const int i = (int)(gg % p[k]);
                gg /= p[k];
                sRow[k * BM + r] = i * Npad;
            }
        } else {
            for (int k = 0; k < d; ++k) sRow[k * BM + r] = -1;
        }
    }
    __syncthreads();

    // U-generation assignment of this thread: one row, KPER consecutive k's
    const int urow = tid / TPR, ukseg = (tid % TPR) * KPER;
    double ureg[KPER];

    auto ldg_u = [&](int kt) {  // table loads + product into registers (stored to smem later)
        const int j = kt * kBK + ukseg;
        const int o0 = sRow[urow];
        if (o0 < 0) {
#pragma unroll
            for (int e = 0; e < KPER; ++e) ureg[e] = 0.0;
            return;
        }
#pragma unroll
        for (int e = 0; e < KPER; e += 2) {
            const double2 t = *reinterpret_cast<const double2*>(tab[0] + o0 + j + e);
            ureg[e] = t.x; ureg[e + 1] = t.y;
        }
#pragma unroll
        for (int k = 1; k < CBO_MAX_D; ++k) {
            if (k < d) {
                const double* __restrict__ tp = tab[k] + sRow[k * BM + urow] + j;
#pragma unroll
                for (int e = 0; e < KPER; e += 2) {
                    const double2 t = *reinterpret_cast<const double2*>(tp + e);
                    ureg[e] *= t.x; ureg[e + 1] *= t.y;
                }
            }
        }
    };
    auto sts_u = [&](int stage) {
        double* slab = sA + stage * Cfg::A_TILE;
#pragma unroll
        for (int e = 0; e < KPER; e += 2) {
            const int k = ukseg + e;
            *reinterpret_cast<double2*>(slab + frag_off(BM, k >> 2, urow, k & 3)) = make_double2(ureg[e], ureg[e + 1]);
        }
    };
    auto load_b = [&](int jb, int kt, int stage) {
        load_rows_async<BN, NT>(sB + stage * Cfg::B_TILE, M + (size_t)jb * BN * Npad + (size_t)kt * kBK, Npad, tid);
    };

    double acc[MA][NB][2];
    double q[MA], mm[MA];
#pragma unroll
    for (int mi = 0; mi < MA; ++mi) {
        q[mi] = 0.0; mm[mi] = 0.0;
#pragma unroll
        for (int ni = 0; ni < NB; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    }

    // flattened (J block, k block) schedule; the producer runs STAGES-1 steps ahead of the consumer
    int pj = 0, pk = 0;  // producer position
    bool pvalid = nJ > 0;
    auto advance_p = [&]() {
        if (++pk == (pj + 1) * KB_PER_J) { pk = 0; if (++pj == nJ) pvalid = false; }
    };

#pragma unroll 1
    for (int i = 0; i < STAGES - 1; ++i) {
        if (pvalid) {
            load_b(pj, pk, i);
            ldg_u(pk);
            sts_u(i);
            advance_p();
        }
        cp_async_commit();
    }

    int cstage = 0;
#pragma unroll 1
    for (int jb = 0; jb < nJ; ++jb) {
        const int nk = (jb + 1) * KB_PER_J, noff = jb * KB_PER_J;
#pragma unroll 1
        for (int kt = 0; kt < nk; ++kt) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            const int pstage = (cstage + STAGES - 1) % STAGES;
            const bool produce = pvalid;
            if (produce) {
                load_b(pj, pk, pstage);
                ldg_u(pk);
            }
            cp_async_commit();
            if (kt == noff) {  // strictly-lower blocks appear twice in u^T M u
#pragma unroll
                for (int mi = 0; mi < MA; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NB; ++ni) { acc[mi][ni][0] *= 2.0; acc[mi][ni][1] *= 2.0; }
            }
            mma_stage<BM, BN, MA, NB>(sA + cstage * Cfg::A_TILE, sB + cstage * Cfg::B_TILE, acc, row0, col0, lane);
            if (produce) {
                sts_u(pstage);
                advance_p();
            }
            cstage = (cstage + 1) % STAGES;
        }
        // J-block epilogue: q_g += sum_{j in J} T[g][j] u[g][j] ; m_g += sum_{j in J} u[g][j] w[j]
#pragma unroll
        for (int mi = 0; mi < MA; ++mi) {
            const int r = row0 + mi * 8 + (lane >> 2);
            const int o0 = sRow[r];
#pragma unroll
            for (int ni = 0; ni < NB; ++ni) {
                const int j = jb * BN + col0 + ni * 8 + (lane & 3) * 2;
                double2 u = make_double2(0.0, 0.0);
                if (o0 >= 0) {
                    u = *reinterpret_cast<const double2*>(tab[0] + o0 + j);
#pragma unroll
                    for (int k = 1; k < CBO_MAX_D; ++k) {
                        if (k < d) {
                            const double2 t = *reinterpret_cast<const double2*>(tab[k] + sRow[k * BM + r] + j);
                            u.x *= t.x; u.y *= t.y;
                        }
                    }
                }
                const double2 ww = *reinterpret_cast<const double2*>(w + j);
                q[mi] = fma(acc[mi][ni][0], u.x, q[mi]);
                q[mi] = fma(acc[mi][ni][1], u.y, q[mi]);
                mm[mi] = fma(u.x, ww.x, mm[mi]);
                mm[mi] = fma(u.y, ww.y, mm[mi]);
                acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            }
        }
    }
    cp_async_wait<0>();

    // reduce the row sums: 4 lanes share a row, then WN warps (fixed order -> deterministic)
#pragma unroll
    for (int mi = 0; mi < MA; ++mi) {
        q[mi] += __shfl_xor_sync(0xffffffffu, q[mi], 1);
        q[mi] += __shfl_xor_sync(0xffffffffu, q[mi], 2);
        mm[mi] += __shfl_xor_sync(0xffffffffu, mm[mi], 1);
        mm[mi] += __shfl_xor_sync(0xffffffffu, mm[mi], 2);
        if ((lane & 3) == 0) {
            const int r = row0 + mi * 8 + (lane >> 2);
            sRed[wn * BM + r] = q[mi];
            sRed[(WN + wn) * BM + r] = mm[mi];
        }
    }
    __syncthreads();
    for (int r = tid; r < BM; r += NT) {
        const long long loc = (long long)tile * BM + r;
        if (loc < gcnt) {
            double qs = 0.0, ms = 0.0;
#pragma unroll
            for (int x = 0; x < WN; ++x) { qs += sRed[x * BM + r]; ms += sRed[(WN + x) * BM + r]; }
            out_m[loc] = ms;
            out_v[loc] = (S.s2 + S.noise) - qs;
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------
using PriorCfg256 = PriorCfg<2, 4, 8, 4, 4>;   // 256 threads, 128 x 128 tile, 64 accumulators/thread
using PriorCfg512 = PriorCfg<4, 4, 4, 4, 4>;   // 512 threads, 128 x 128 tile, 32 accumulators/thread
using PriorCfg64 = PriorCfg<4, 2, 4, 4, 5>;    // 256 threads, 128 x 64 tile

int prior_variant();  // api.cu: CBO_PRIOR_VARIANT env (0 = default)

template <class Cfg>
static int launch_prior(const cbo_set_desc* d_sets, int num_sets, int which, int total, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        CBO_CUDA(cudaFuncSetAttribute(prior_eval_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        configured = true;
    }
    prior_eval_kernel<Cfg><<<total, Cfg::NT, Cfg::SMEM, st>>>(d_sets, num_sets, which);
    CBO_CUDA(cudaGetLastError());
    return 0;
}

int prior_eval_impl(const cbo_set_desc* h_sets, const cbo_set_desc* d_sets, int num_sets, int which, cudaStream_t st) {
    long long total = 0;
    for (int s = 0; s < num_sets; ++s) {
        const cbo_set_desc& S = h_sets[s];
        if (!S.causal) continue;
        for (int k = 0; k < (which == 0 ? S.d : 1); ++k) {
            const long long pk = which == 0 ? S.p[k] : S.n_int;
            CBO_REQUIRE(pk * (long long)S.n_obs_pad < 2147483647LL, "cbo_prior_eval: table %d of set %d too large", k, s);
        }
        total += host_items(S, which == 0 ? kItemsPriorGrid : kItemsPriorTrain);
    }
    CBO_REQUIRE(total < 2147483647LL, "cbo_prior_eval: too many work items");
    if (total == 0) return 0;
    switch (prior_variant()) {
        case 1: return launch_prior<PriorCfg512>(d_sets, num_sets, which, (int)total, st);
        case 2: return launch_prior<PriorCfg64>(d_sets, num_sets, which, (int)total, st);
        default: return launch_prior<PriorCfg256>(d_sets, num_sets, which, (int)total, st);
    }
}

}  // namespace cbo
