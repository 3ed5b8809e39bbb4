// K1b -- causal prior on a tensor grid:  m(x) = u(x).w ,  v(x) = s2 + noise - u(x)^T M u(x).
//
// Replaces DoCalculus.update_do_function / compute_do (reference DoCalculus.py:34-89): the reference
// rebuilds an (N x D) design per candidate and runs a full GP predict (O(N^2 D + N^3)); here the
// Monte-Carlo average has been folded into w and M once per observation (K1a) and each candidate costs a
// dense symmetric quadratic form in its N-vector u(x) = prod_k tab_k[i_k(x)][:].
//
// Shape of the work: T = U (G x N) * M (N x N) followed by a row-wise dot with U.  Only M's lower
// block triangle is visited: for column block J the k-blocks below the diagonal block are accumulated,
// doubled, and the (full, symmetric) diagonal block is added.  One CTA owns 128 candidates.
//
// Warp-specialised, mbarrier-synchronised pipeline (one CTA per SM, persistent over its k loop):
//   producer warps (4)  : per stage, one thread issues a single 16 KB TMA bulk copy (cp.async.bulk, SASS UBLKCP) of
//                         the M slab -- K1a stores M blocked in fragment order, so the slab is contiguous -- whose
//                         bytes complete on the stage's "full" mbarrier; all producer threads REGENERATE the
//                         128 x 16 slab of U from the L2-resident exp tables (d loads + (d-1) multiplies per
//                         element, conflict-free STS.128 in DMMA fragment order); all global-load latency is
//                         absorbed here;
//   consumer warps (8)  : wait "full", 4 x (12 LDS.64 + 32 DMMA.8x8x4), arrive on "empty"; they never touch
//                         global memory inside the k loop, so the FP64 tensor pipe is the only thing they wait for.
// Registers are re-balanced with setmaxnreg (producers shrink, consumers hold the 128 x 128 accumulator tile).
// Roofline: FP64 pipe.  Executed flops per candidate = N^2 (+ lower order); the dense-counted figure of
// SURVEY.md §8(d) is 2 N^2 + 2 N + d N.
#include "dmma_tile.cuh"

namespace cbo {

// work items of one set for the launch that handles sets with D tables
__host__ __device__ inline long long prior_items(const cbo_set_desc& S, int which, int D) {
    if (!S.causal) return 0;
    if (which == 0) return S.d == D ? (S.g_count + CBO_PRIOR_TILE - 1) / CBO_PRIOR_TILE : 0;
    return D == 1 ? (S.n_int + CBO_PRIOR_TILE - 1) / CBO_PRIOR_TILE : 0;
}

template <int WM_, int WN_, int MA_, int NB_, int STAGES_, int PW_>
struct PriorCfg {
    static constexpr int WM = WM_, WN = WN_, MA = MA_, NB = NB_, STAGES = STAGES_, PW = PW_;
    static constexpr int BM = WM * MA * 8;  // candidates per CTA
    static constexpr int BN = WN * NB * 8;  // columns of M per J block
    static constexpr int NPROD = PW * 32, NCONS = WM * WN * 32, NT = NPROD + NCONS;
    static constexpr int A_TILE = BM * kBK;
    static constexpr int B_TILE = BN * kBK;
    static constexpr int ROWS_PER_PASS = NPROD / 8;        // 8 lanes (16 B each) cover one 128-byte row segment
    static constexpr int PASSES = BM / ROWS_PER_PASS;
    static constexpr int PASS_GROUP = 4;                   // passes whose loads are in flight together
    static constexpr size_t SMEM = (size_t)STAGES * (A_TILE + B_TILE) * sizeof(double) + 2 * WN * BM * sizeof(double) +
                                   (CBO_MAX_D + 1) * BM * sizeof(int32_t) + 2 * STAGES * sizeof(uint64_t);
    static_assert(BM == CBO_PRIOR_TILE, "host item count assumes CBO_PRIOR_TILE points per CTA");
    static_assert(PW % 4 == 0 && (WM * WN) % 4 == 0, "setmaxnreg works on whole warpgroups");
    static_assert(BM % ROWS_PER_PASS == 0 && PASSES % PASS_GROUP == 0, "U generation mapping");
    static_assert(BN == kMBlkRows, "the J block must match the row block of M's blocked layout");
    static_assert(NPROD == 128, "producer lane mapping assumes 4 warps x 4 rows per pass");
};

template <class Cfg, int D, int PROD_REGS, int CONS_REGS>
__global__ void __launch_bounds__(Cfg::NT, 1)
prior_eval_kernel(const cbo_set_desc* __restrict__ sets, int num_sets, int which) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, MA = Cfg::MA, NB = Cfg::NB, WN = Cfg::WN;
    constexpr int STAGES = Cfg::STAGES, NPROD = Cfg::NPROD;
    constexpr int KB_PER_J = BN / kBK;
    // setmaxnreg can only redistribute the CTA's own allocation: NT x (registers per thread at launch, which
    // ptxas pins to the __launch_bounds__ ceiling, a multiple of 8).  Asking for more deadlocks the consumers.
    constexpr int LAUNCH_REGS = (65536 / Cfg::NT) / 8 * 8;
    static_assert(NPROD * PROD_REGS + Cfg::NCONS * CONS_REGS <= Cfg::NT * LAUNCH_REGS, "setmaxnreg budget exceeds the CTA's register pool");
    static_assert(PROD_REGS % 8 == 0 && CONS_REGS % 8 == 0, "setmaxnreg takes multiples of 8");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sB = sA + STAGES * Cfg::A_TILE;
    double* sRed = sB + STAGES * Cfg::B_TILE;                        // [2][WN][BM] running q and m partial sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 2 * WN * BM);  // full[STAGES], empty[STAGES]
    int32_t* sRow = reinterpret_cast<int32_t*>(bars + 2 * STAGES);     // [D + 1][BM]: table row offsets, then a live flag
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // flat work item -> (set, tile): sets are scanned in order, only those with D tables count
    int tile = blockIdx.x, s = 0;
    for (; s < num_sets - 1; ++s) {
        const int cnt = (int)prior_items(sets[s], which, D);
        if (tile < cnt) break;
        tile -= cnt;
    }
    const cbo_set_desc& S = sets[s];

    // effective problem: the tensor grid (which == 0) or the n_int interventional rows (which == 1)
    // (the kernel is instantiated per number of tables D: d for the grid, 1 for the interventional rows)
    const double* tab[D];
    int p[D];
    long long gbeg, gcnt;
    double *out_m, *out_v;
    if (which == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) { tab[k] = S.tab[k]; p[k] = S.p[k]; }
        gbeg = S.g_begin; gcnt = S.g_count; out_m = S.m; out_v = S.v;
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) { tab[k] = S.u_int; p[k] = S.n_int; }
        gbeg = 0; gcnt = S.n_int; out_m = S.m_int; out_v = S.v_int;
    }
    const int Npad = S.n_obs_pad;
    const double* __restrict__ M = S.M;
    const double* __restrict__ w = S.w;
    const int nJ = (S.n_obs + BN - 1) / BN;

    // per-row table offsets (C-order decomposition of the flat grid index, last dim fastest)
    for (int r = tid; r < BM; r += Cfg::NT) {
        const long long loc = (long long)tile * BM + r;
        if (loc < gcnt) {
            long long gg = gbeg + loc;
#pragma unroll
            for (int k = D - 1; k >= 0; --k) {
                const int i = (int)(gg % p[k]);
                gg /= p[k];
                sRow[k * BM + r] = i * Npad;
            }
            sRow[D * BM + r] = 1;
        } else {  // rows past the slice read table row 0 and are masked to zero
#pragma unroll
            for (int k = 0; k < D; ++k) sRow[k * BM + r] = 0;
            sRow[D * BM + r] = 0;
        }
    }
    for (int i = tid; i < 2 * WN * BM; i += Cfg::NT) sRed[i] = 0.0;
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], NPROD + 1);        // every producer thread after its U rows + the expect_tx arrive of the TMA issuer
            mbar_init(&empty[i], Cfg::WM * WN);  // one arrive per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp < Cfg::PW) {
        // =============================== PRODUCER ===============================
        setmaxnreg_dec<PROD_REGS>();
        // lane -> (k4 group, row, half): a quarter-warp writes 128 contiguous bytes of the slab (conflict-free
        // STS.128) and the warp reads four full 128-byte table lines per load instruction
        const int pkb = lane >> 3, phalf = lane & 1;
        const int prow0 = warp * 4 + ((lane & 7) >> 1);
        const int jlane = pkb * 4 + phalf * 2;
        constexpr int PG = Cfg::PASS_GROUP;
        double2 t[PG][D];
        auto issue_loads = [&](int g0, int kt) {  // every load of the group is issued before the first use
            const int j = kt * kBK + jlane;
#pragma unroll
            for (int q = 0; q < PG; ++q) {
                const int row = prow0 + (g0 + q) * Cfg::ROWS_PER_PASS;
#pragma unroll
                for (int k = 0; k < D; ++k) t[q][k] = ldg_nc_d2(tab[k] + sRow[k * BM + row] + j);
            }
        };
        auto finish_group = [&](int g0, double* slab) {
#pragma unroll
            for (int q = 0; q < PG; ++q) {
                const int row = prow0 + (g0 + q) * Cfg::ROWS_PER_PASS;
                const double live = (double)sRow[D * BM + row];
                double2 v = make_double2(t[q][0].x * live, t[q][0].y * live);
#pragma unroll
                for (int k = 1; k < D; ++k) { v.x *= t[q][k].x; v.y *= t[q][k].y; }
                *reinterpret_cast<double2*>(slab + frag_off(BM, pkb, row, phalf * 2)) = v;
            }
        };
        int stage = 0;
        unsigned phase = 0;
#pragma unroll 1
        for (int jb = 0; jb < nJ; ++jb) {
            const int nk = (jb + 1) * KB_PER_J;
#pragma unroll 1
            for (int kt = 0; kt < nk; ++kt) {
                issue_loads(0, kt);                       // table loads do not need the smem slot: start them first
                mbar_wait(&empty[stage], phase ^ 1u);
                if (tid == 0) {                           // M slab: one contiguous 16 KB TMA bulk copy
                    mbar_arrive_expect_tx(&full[stage], Cfg::B_TILE * sizeof(double));
                    bulk_g2s(sB + stage * Cfg::B_TILE, M + mblk_base(jb, kt, Npad), Cfg::B_TILE * sizeof(double), &full[stage]);
                }
                double* slab = sA + stage * Cfg::A_TILE;
                finish_group(0, slab);
#pragma unroll
                for (int g0 = PG; g0 < Cfg::PASSES; g0 += PG) {
                    issue_loads(g0, kt);
                    finish_group(g0, slab);
                }
                mbar_arrive(&full[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // =============================== CONSUMER ===============================
        setmaxnreg_inc<CONS_REGS>();
        const int cwarp = warp - Cfg::PW;
        const int wm = cwarp / WN, wn = cwarp % WN;
        const int row0 = wm * MA * 8, col0 = wn * NB * 8;
        double acc[MA][NB][2];
#pragma unroll
        for (int mi = 0; mi < MA; ++mi)
#pragma unroll
            for (int ni = 0; ni < NB; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

        int stage = 0;
        unsigned phase = 0;
#pragma unroll 1
        for (int jb = 0; jb < nJ; ++jb) {
            const int nk = (jb + 1) * KB_PER_J, noff = jb * KB_PER_J;
#pragma unroll 1
            for (int kt = 0; kt < nk; ++kt) {
                if (kt == noff) {  // strictly-lower blocks appear twice in u^T M u
#pragma unroll
                    for (int mi = 0; mi < MA; ++mi)
#pragma unroll
                        for (int ni = 0; ni < NB; ++ni) { acc[mi][ni][0] *= 2.0; acc[mi][ni][1] *= 2.0; }
                }
                mbar_wait(&full[stage], phase);
                mma_stage<BM, BN, MA, NB>(sA + stage * Cfg::A_TILE, sB + stage * Cfg::B_TILE, acc, row0, col0, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            // J-block epilogue: q_g += sum_{j in J} T[g][j] u[g][j] ; m_g += sum_{j in J} u[g][j] w[j]
#pragma unroll
            for (int mi = 0; mi < MA; ++mi) {
                const int r = row0 + mi * 8 + (lane >> 2);
                const double live = (double)sRow[D * BM + r];
                double q = 0.0, mm = 0.0;
#pragma unroll
                for (int ni = 0; ni < NB; ++ni) {
                    const int j = jb * BN + col0 + ni * 8 + (lane & 3) * 2;
                    double2 u = ldg_nc_d2(tab[0] + sRow[r] + j);
                    u.x *= live; u.y *= live;
#pragma unroll
                    for (int k = 1; k < D; ++k) {
                        const double2 t = ldg_nc_d2(tab[k] + sRow[k * BM + r] + j);
                        u.x *= t.x; u.y *= t.y;
                    }
                    const double2 ww = ldg_nc_d2(w + j);
                    q = fma(acc[mi][ni][0], u.x, q);
                    q = fma(acc[mi][ni][1], u.y, q);
                    mm = fma(u.x, ww.x, mm);
                    mm = fma(u.y, ww.y, mm);
                    acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                }
                q += __shfl_xor_sync(0xffffffffu, q, 1);
                q += __shfl_xor_sync(0xffffffffu, q, 2);
                mm += __shfl_xor_sync(0xffffffffu, mm, 1);
                mm += __shfl_xor_sync(0xffffffffu, mm, 2);
                if ((lane & 3) == 0) {  // (wn, r) has exactly one owner: no race, fixed order -> deterministic
                    sRed[wn * BM + r] += q;
                    sRed[(WN + wn) * BM + r] += mm;
                }
            }
        }
    }
    __syncthreads();
    for (int r = tid; r < BM; r += Cfg::NT) {
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
using PriorCfgA = PriorCfg<2, 4, 8, 4, 4, 4>;   // 4 producer + 8 consumer warps, 128 x 128 tile, 64 accumulators/thread

int prior_variant();  // api.cu: CBO_PRIOR_VARIANT env (0 = default)

template <class Cfg, int D, int PR, int CR>
static int launch_prior(const cbo_set_desc* d_sets, int num_sets, int which, int total, cudaStream_t st) {
    static bool configured = false;
    auto kern = prior_eval_kernel<Cfg, D, PR, CR>;
    if (!configured) {
        CBO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        configured = true;
    }
    kern<<<total, Cfg::NT, Cfg::SMEM, st>>>(d_sets, num_sets, which);
    CBO_CUDA(cudaGetLastError());
    return 0;
}

// One launch per distinct number of tables D among the causal sets: the descriptor list handed to the kernel is
// the caller's, the kernel skips sets whose d differs (they contribute zero work items, see items_for()).
int prior_eval_impl(const cbo_set_desc* h_sets, const cbo_set_desc* d_sets, int num_sets, int which, cudaStream_t st) {
    for (int D = 1; D <= CBO_MAX_D; ++D) {
        long long total = 0;
        for (int s = 0; s < num_sets; ++s) {
            const cbo_set_desc& S = h_sets[s];
            if (!S.causal) continue;
            for (int k = 0; k < (which == 0 ? S.d : 1); ++k) {
                const long long pk = which == 0 ? S.p[k] : S.n_int;
                CBO_REQUIRE(pk * (long long)S.n_obs_pad < 2147483647LL, "cbo_prior_eval: table %d of set %d too large", k, s);
            }
            total += prior_items(S, which, D);
        }
        CBO_REQUIRE(total < 2147483647LL, "cbo_prior_eval: too many work items");
        if (total == 0) continue;
        int rc = 0;
        switch (D) {
            case 1: rc = launch_prior<PriorCfgA, 1, 96, 200>(d_sets, num_sets, which, (int)total, st); break;
            case 2: rc = launch_prior<PriorCfgA, 2, 96, 200>(d_sets, num_sets, which, (int)total, st); break;
            case 3: rc = launch_prior<PriorCfgA, 3, 96, 200>(d_sets, num_sets, which, (int)total, st); break;
            default: rc = launch_prior<PriorCfgA, 4, 96, 200>(d_sets, num_sets, which, (int)total, st); break;
        }
        if (rc) return rc;
    }
    return 0;
}

}  // namespace cbo
