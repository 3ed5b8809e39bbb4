// K1p -- the causal prior on a tensor grid when the observational GP is SMALL (n_obs <= 256: the reference's shipped
// data sizes, BASELINE.json configs 2-4).  Shared declarations; the kernels are in prior_pair.cu.
//
// At N = 100..200 the quadratic form u^T M u of one candidate is only 1e4..4e4 multiply-adds: a 128-candidate work item of
// the general kernel (prior_eval.cu) lasts 8-24 pipeline stages and its per-item fixed costs dominate.  On a TENSOR grid the
// work can be regrouped.  With u(x) = t0[i0] o t1[i1] o t2[i2] (o = elementwise, t_k = rows of the exp tables) and
// r = (j, k), j <= k, running over the N (N + 1) / 2 index pairs of the symmetric M:
//     u^T M u = sum_r  C0[i0][r] * C1[i1][r] * C2[i2][r]
//     C0[i][r] = (j == k ? 1 : 2) M_jk t0[i]_j t0[i]_k ,   C1[i][r] = t1[i]_j t1[i]_k ,   C2[i][r] = t2[i]_j t2[i]_k
// so for a fixed i0 the whole (i1, i2) plane is ONE GEMM  Q = (C1 scaled column-wise by C0[i0]) . C2^T  with a reduction
// length of R = N (N + 1) / 2 = 5050 .. 32896: a work item is a p1 x p2 output tile with hundreds of pipeline stages, both
// operands are small "pair tables" (p x R doubles, L2 resident, the same for every CTA) and nothing is materialised per
// candidate.  The mean m = u . w is the same product over N extra columns (C0 carries w).  A d = 2 grid is the same GEMM
// without the scale row.  Executed flops per candidate: N^2 + 3 N (the symmetric half of M, no block padding).
#pragma once
#include "cbo_common.cuh"

namespace cbo {

constexpr int kPairMaxN = 256;      // largest n_obs that takes this path (pair tables: p x N(N+1)/2 doubles per dimension)
constexpr int kPairChunkRows = 104; // rows of one operand chunk (one GEMM tile side) at most: 13 x 13 blocks of 8 x 8, whose
                                    // second-level accumulators (prior_pair.cu) still fit in shared memory next to the ring
constexpr int kMaxPairSets = 32;    // exploration sets per launch (kernel-parameter table)

__host__ __device__ inline bool pair_eligible(const cbo_set_desc& S) {
    return computes_prior(S) && !S.points && (S.d == 2 || S.d == 3) && S.n_obs <= kPairMaxN && S.g_count > 0;
}

// Geometry of one set on this path.  Operand A is dimension d-2 (rows of the output tile), operand B dimension d-1 (the
// fastest grid dimension: columns of the tile), the scale row dimension 0 of a d = 3 grid.  An operand table is stored per
// chunk of CR rows and per 16-deep slab in DMMA fragment order, so that one pipeline stage is one contiguous TMA bulk copy:
//   element (row i, column q) of the table -> chunk c = i / CR, lr = i % CR, slab = q / 16:
//   ((c * Kslabs + slab) * CR * 16) + ((((q % 16) / 4) * CR + lr) * 4) + q % 4
// The scale table is plain row-major (p0, Kslabs * 16).  Columns [0, R) are the pairs r = k (k + 1) / 2 + j (j <= k), zero
// padded to Rslabs * 16; then N columns for the mean, zero padded to a multiple of 16.
struct PairGeom {
    int d, N, R, Rslabs, Kslabs;
    int pa, pb, ps;
    int CRa, nchA, CRb, nchB;
    int s_begin, s_count;          // scale rows that intersect the rank's slice of the grid (d = 2: 0, 1)
    long long szA, szB, szS;       // doubles
};

__host__ __device__ inline PairGeom pair_geom(const cbo_set_desc& S) {
    PairGeom g;
    g.d = S.d;
    g.N = S.n_obs;
    g.R = g.N * (g.N + 1) / 2;
    g.Rslabs = (g.R + 15) / 16;
    g.Kslabs = g.Rslabs + (g.N + 15) / 16;
    g.pa = S.p[S.d - 2];
    g.pb = S.p[S.d - 1];
    g.ps = S.d == 3 ? S.p[0] : 1;
    g.nchA = (g.pa + kPairChunkRows - 1) / kPairChunkRows;
    g.CRa = (((g.pa + g.nchA - 1) / g.nchA) + 7) / 8 * 8;
    g.nchB = (g.pb + kPairChunkRows - 1) / kPairChunkRows;
    g.CRb = (((g.pb + g.nchB - 1) / g.nchB) + 7) / 8 * 8;
    const long long plane = (long long)g.pa * g.pb;
    g.s_begin = S.d == 3 ? (int)(S.g_begin / plane) : 0;
    g.s_count = S.d == 3 ? (int)((S.g_begin + S.g_count - 1) / plane) - g.s_begin + 1 : 1;
    g.szA = (long long)g.nchA * g.CRa * g.Kslabs * 16;
    g.szB = (long long)g.nchB * g.CRb * g.Kslabs * 16;
    g.szS = S.d == 3 ? (long long)g.ps * g.Kslabs * 16 : 0;
    return g;
}

__host__ __device__ inline long long pair_items(const cbo_set_desc& S) {
    if (!pair_eligible(S)) return 0;
    const PairGeom g = pair_geom(S);
    return (long long)g.s_count * g.nchA * g.nchB;
}

// doubles of workspace the pair tables of every eligible set need (plus the 16 ones of the d = 2 scale row)
size_t pair_area_doubles(const cbo_set_desc* h_sets, int num_sets);
long long pair_items_total(const cbo_set_desc* h_sets, int num_sets);
// builds the pair tables of every eligible set in `area` and evaluates m, v on the rank's slice of their grids
int prior_pair_impl(const cbo_set_desc* h_sets, const cbo_set_desc* d_sets, int num_sets, double* area, int num_ctas,
                    cudaStream_t st);

}  // namespace cbo
