// K1p -- causal prior m(x) = u.w, v(x) = s2 + noise - u^T M u on a tensor grid for small observational sets, as a GEMM
// over index pairs (see prior_pair.cuh for the algebra).  Replaces DoCalculus.update_do_function / compute_do
// (reference DoCalculus.py:34-89) on the shipped data sizes (CoralGraph.py:163-184: 25 sets, 100-point grids, N = 100..1000).
//
//   pair_tables_kernel : the per-dimension pair tables C_k (p_k x (R + N) doubles each) from the exp tables, M and w.
//                        Written once per call, 13 MB per set at N = 100, L2 resident while the set is swept.
//   prior_pair_kernel  : persistent, one CTA per SM, work items dealt round-robin (all items of a set cost the same).
//                        item = (scale row i0, A chunk, B chunk): a (<= 104) x (<= 104) tile of the (i1, i2) plane.
//                        producer : one elected thread; per 16-deep stage three TMA bulk copies (A slab, B slab, 16 scale
//                                   values) completing on the stage's mbarrier; 5-stage ring carried across items, so the
//                                   next item's first slabs arrive while the consumers store the current item's results.
//                        consumers: 8 warps in a 2 x 4 arrangement over the LIVE 8 x 8 blocks of the tile (13 x 13 at
//                                   p = 100: 7|6 row blocks x 4|3|3|3 column blocks, mirrored in the second row half so
//                                   the four schedulers carry 46/39/39/45 blocks); per k4 step NR + NC conflict-free
//                                   LDS.64, NC DMULs (the scale row enters through the B fragment) and NR x NC DMMA.8x8x4.
//                                   After the R pair columns the accumulators ARE u^T M u: v is stored straight from the
//                                   fragments; the N mean columns follow in the same ring and m is stored the same way.
// Roofline: FP64 tensor pipe; executed flops per candidate N^2 + 3 N (+ tile padding).  No per-candidate global traffic
// besides the 16 B of output.
#include "dmma_tile.cuh"
#include "prior_pair.cuh"

namespace cbo {

struct PairSet {
    int set, item_base, d;
    int N, R, Rslabs, Kslabs;
    int pa, pb, ps;
    int CRa, nchA, CRb, nchB;
    int s_begin, s_count;
    long long offA, offB, offS;    // doubles from the start of the pair area (offS unused when d == 2)
};
struct PairLaunch {
    int n, total_items;
    PairSet e[kMaxPairSets];
};

constexpr int kPairStages = 5;
constexpr int kPairTile = kPairChunkRows * kBK;              // doubles of one operand slab buffer (13 KB)
constexpr int kPairStageDoubles = 2 * kPairTile + kBK;       // A slab, B slab, 16 scale values
constexpr int kPairNCons = 256, kPairNT = kPairNCons + 128;  // 8 consumer warps + one producer warpgroup
// Second-level accumulators: the reduction runs over 5e3 .. 3.3e4 pair columns whose partial sums are orders of magnitude
// larger than the result (u^T M u cancels to 1e-3 .. 1e-8 of its terms on the shipped data), and a single FP64
// accumulator chain of that length loses ~sqrt(length) more than the general kernel's N-long chains do (measured on the
// coral fixture: 2.5e-7 in v against 1e-8).  Every kPairFlush slabs the register accumulators are added into a
// per-thread slot in shared memory and restart from zero (blocked summation).
constexpr int kPairFlush = 32;
constexpr int kPairBlocks = (kPairChunkRows / 8) * (kPairChunkRows / 8);
constexpr int kPairAcc2Doubles = kPairBlocks * 64;
constexpr size_t kPairSmem = ((size_t)kPairStages * kPairStageDoubles + kPairAcc2Doubles) * sizeof(double) +
                             2 * kPairStages * sizeof(uint64_t) + 16;
static_assert(kPairSmem <= 232448, "shared memory of one CTA");
constexpr int kPairProdRegs = 40, kPairConsRegs = 232;
static_assert(128 * kPairProdRegs + kPairNCons * kPairConsRegs <= kPairNT * ((65536 / kPairNT) / 8 * 8), "setmaxnreg budget");

// ---------------------------------------------------------------------------------------------------------------------
// pair tables.  grid = (slab groups, sets of the launch, 3 tables: A, B, scale); a warp owns one 16-deep slab at a time:
// lane (g, t) = (row inside an 8-row group, column inside a k4 group); the column's (j, k) pair and its coefficient are
// decoded once per k4 group and reused for every row group.
__global__ void __launch_bounds__(256)
pair_tables_kernel(const cbo_set_desc* __restrict__ sets, const __grid_constant__ PairLaunch L, double* __restrict__ area) {
    const PairSet& E = L.e[blockIdx.y];
    const cbo_set_desc& S = sets[E.set];
    const int table = blockIdx.z;
    if (blockIdx.x == 0 && blockIdx.y == 0 && table == 0 && threadIdx.x < kBK) area[threadIdx.x] = 1.0;   // scale row of d = 2 grids
    if (table == 2 && E.d != 3) return;
    const int dim = table == 0 ? E.d - 2 : (table == 1 ? E.d - 1 : 0);
    const bool with_m = dim == 0;               // dimension 0's table carries M (pairs) and w (mean columns)
    const int p = S.p[dim], npad = S.n_obs_pad, N = E.N;
    const int CR = table == 0 ? E.CRa : E.CRb, nch = table == 0 ? E.nchA : E.nchB;
    const int rows = table == 2 ? E.ps : nch * CR;
    double* __restrict__ out = area + (table == 0 ? E.offA : (table == 1 ? E.offB : E.offS));
    const double* __restrict__ tab = S.tab[dim];
    const double* __restrict__ M = S.M;
    const double* __restrict__ w = S.w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int slab = blockIdx.x * 8 + warp; slab < E.Kslabs; slab += gridDim.x * 8) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
            const int q = slab * kBK + kb * 4 + t;
            int j = 0, k = 0;
            double coef = 0.0;               // 0: a padding column
            if (q < E.R) {
                k = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
                while (k * (k + 1) / 2 > q) --k;
                while ((k + 1) * (k + 2) / 2 <= q) ++k;
                j = q - k * (k + 1) / 2;
                coef = with_m ? M[mblk_off(k, j, npad)] * (j == k ? 1.0 : 2.0) : 1.0;
            } else if (slab >= E.Rslabs) {
                j = k = q - E.Rslabs * kBK;
                if (j < N) coef = with_m ? w[j] : 1.0;
                else j = k = 0;
            }
            const bool pair = q < E.R;
            for (int i = g; i < rows; i += 8) {
                double val = 0.0;
                if (i < p && coef != 0.0) {
                    const double tj = tab[(size_t)i * npad + j];
                    val = pair ? (tj * tab[(size_t)i * npad + k]) * coef : tj * coef;
                }
                if (table == 2) out[(size_t)i * E.Kslabs * kBK + q] = val;
                else {
                    const int c = i / CR, lr = i - c * CR;
                    out[((size_t)c * E.Kslabs + slab) * CR * kBK + (((kb * CR + lr) << 2) + t)] = val;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
struct PairItem {
    const double *A, *B, *Sc;      // first slab of the item's A chunk / B chunk / scale row (Sc stride 0 when d == 2)
    int sc_stride;                 // doubles between consecutive scale slabs (16, or 0: the constant ones)
    int CRa, CRb, Rslabs, Kslabs;
    int pa, pb, ca, cb, s;
    unsigned bytes;                // TMA bytes per stage
    const cbo_set_desc* S;
};

__device__ __forceinline__ PairItem decode_pair_item(const cbo_set_desc* __restrict__ sets, const PairLaunch& L,
                                                     const double* __restrict__ area, int item) {
    int e = 0;
#pragma unroll 1
    for (; e < L.n - 1; ++e)
        if (item < L.e[e + 1].item_base) break;
    const PairSet& E = L.e[e];
    int local = item - E.item_base;
    PairItem it;
    it.cb = local % E.nchB; local /= E.nchB;
    it.ca = local % E.nchA; local /= E.nchA;
    it.s = E.s_begin + local;
    it.CRa = E.CRa, it.CRb = E.CRb, it.Rslabs = E.Rslabs, it.Kslabs = E.Kslabs, it.pa = E.pa, it.pb = E.pb;
    it.A = area + E.offA + (size_t)it.ca * E.Kslabs * E.CRa * kBK;
    it.B = area + E.offB + (size_t)it.cb * E.Kslabs * E.CRb * kBK;
    if (E.d == 3) { it.Sc = area + E.offS + (size_t)it.s * E.Kslabs * kBK; it.sc_stride = kBK; }
    else { it.Sc = area; it.sc_stride = 0; }
    it.bytes = (unsigned)((E.CRa + E.CRb + 1) * kBK * sizeof(double));
    it.S = sets + E.set;
    return it;
}

// Consumer side of one item for a warp that owns NR x NC live 8 x 8 blocks starting at (row0, col0) of the tile.
template <int NR, int NC>
__device__ __forceinline__ void pair_consume(const PairItem& it, const double* __restrict__ smem, double* __restrict__ acc2,
                                             uint64_t* full, uint64_t* empty, int& stage, unsigned& phase, int lane, int row0,
                                             int col0, int CB) {
    if constexpr (NR == 0 || NC == 0) {   // no live block: walk the ring so that the arrival counts match
#pragma unroll 1
        for (int s = 0; s < it.Kslabs; ++s) {
            mbar_wait(&full[stage], phase);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
        return;
    } else {
        const cbo_set_desc& S = *it.S;
        const int CRa = it.CRa, CRb = it.CRb;
        const int aoff = (row0 << 2) + lane, boff = (col0 << 2) + lane;
        // this thread's second-level slots: block (mi, ni) of the warp -> tile block ((row0/8 + mi) * CB + col0/8 + ni)
        double2* const my2 = reinterpret_cast<double2*>(acc2) + ((row0 >> 3) * CB + (col0 >> 3)) * 32 + lane;
        double acc[NR][NC][2];
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {     // 0: the pair columns -> v ; 1: the mean columns -> m
#pragma unroll
            for (int mi = 0; mi < NR; ++mi)
#pragma unroll
                for (int ni = 0; ni < NC; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            const int nslab = part == 0 ? it.Rslabs : it.Kslabs - it.Rslabs;
            int since = 0;
            bool have2 = false;
#pragma unroll 1
            for (int s = 0; s < nslab; ++s) {
                mbar_wait(&full[stage], phase);
                const double* __restrict__ sA = smem + (size_t)stage * kPairStageDoubles;
                const double* __restrict__ sB = sA + kPairTile;
                const double* __restrict__ sS = sB + kPairTile;
#pragma unroll
                for (int kb = 0; kb < kBK / 4; ++kb) {
                    const double sc = sS[kb * 4 + (lane & 3)];
                    double a[NR], b[NC];
#pragma unroll
                    for (int mi = 0; mi < NR; ++mi) a[mi] = sA[((kb * CRa + mi * 8) << 2) + aoff];
#pragma unroll
                    for (int ni = 0; ni < NC; ++ni) b[ni] = sB[((kb * CRb + ni * 8) << 2) + boff] * sc;
#pragma unroll
                    for (int mi = 0; mi < NR; ++mi)
#pragma unroll
                        for (int ni = 0; ni < NC; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
                if (++since == kPairFlush && s + 1 < nslab) {
#pragma unroll
                    for (int mi = 0; mi < NR; ++mi)
#pragma unroll
                        for (int ni = 0; ni < NC; ++ni) {
                            double2* p2 = my2 + (mi * CB + ni) * 32;
                            double2 o = have2 ? *p2 : make_double2(0.0, 0.0);
                            o.x += acc[mi][ni][0], o.y += acc[mi][ni][1];
                            *p2 = o;
                            acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                        }
                    have2 = true, since = 0;
                }
            }
            if (have2) {
#pragma unroll
                for (int mi = 0; mi < NR; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NC; ++ni) {
                        const double2 o = my2[(mi * CB + ni) * 32];
                        acc[mi][ni][0] += o.x, acc[mi][ni][1] += o.y;
                    }
            }
            // straight from the fragments: lane (g, t) holds C[g][2t], C[g][2t+1] of every block
            double* __restrict__ out = part == 0 ? S.v : S.m;
            const double sn = S.s2 + S.noise;
            const long long gb = S.g_begin, gc = S.g_count;
#pragma unroll
            for (int mi = 0; mi < NR; ++mi) {
                const int ia = it.ca * CRa + row0 + mi * 8 + (lane >> 2);
                if (ia >= it.pa) continue;
                const long long base = ((long long)it.s * it.pa + ia) * it.pb - gb;
#pragma unroll
                for (int ni = 0; ni < NC; ++ni) {
                    const int ib = it.cb * CRb + col0 + ni * 8 + 2 * (lane & 3);
#pragma unroll
                    for (int x = 0; x < 2; ++x) {
                        const long long loc = base + ib + x;
                        if (ib + x < it.pb && loc >= 0 && loc < gc) out[loc] = part == 0 ? sn - acc[mi][ni][x] : acc[mi][ni][x];
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kPairNT, 1)
prior_pair_kernel(const cbo_set_desc* __restrict__ sets, const __grid_constant__ PairLaunch L, const double* __restrict__ area) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    double* acc2 = smem + (size_t)kPairStages * kPairStageDoubles;
    uint64_t* full = reinterpret_cast<uint64_t*>(acc2 + kPairAcc2Doubles);
    uint64_t* empty = full + kPairStages;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < kPairStages; ++i) {
            mbar_init(&full[i], 1);                   // the producer's arrive.expect_tx; the bytes come from the TMA
            mbar_init(&empty[i], kPairNCons / 32);    // one arrive per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();
    int stage = 0;          // ring position, carried across items: producer and consumers advance identically
    unsigned phase = 0;
    if (warp >= kPairNCons / 32) {
        setmaxnreg_dec<kPairProdRegs>();
        if (warp == kPairNCons / 32 && lane == 0) {
#pragma unroll 1
            for (int item = blockIdx.x; item < L.total_items; item += gridDim.x) {
                const PairItem it = decode_pair_item(sets, L, area, item);
                const unsigned ab = (unsigned)(it.CRa * kBK * sizeof(double)), bb = (unsigned)(it.CRb * kBK * sizeof(double));
#pragma unroll 1
                for (int s = 0; s < it.Kslabs; ++s) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&full[stage], it.bytes);
                    double* dst = smem + (size_t)stage * kPairStageDoubles;
                    bulk_g2s(dst, it.A + (size_t)s * it.CRa * kBK, ab, &full[stage]);
                    bulk_g2s(dst + kPairTile, it.B + (size_t)s * it.CRb * kBK, bb, &full[stage]);
                    bulk_g2s(dst + 2 * kPairTile, it.Sc + (size_t)s * it.sc_stride, (unsigned)(kBK * sizeof(double)), &full[stage]);
                    if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        setmaxnreg_inc<kPairConsRegs>();
        int prev_geom = -1;
#pragma unroll 1
        for (int item = blockIdx.x; item < L.total_items; item += gridDim.x) {
            const PairItem it = decode_pair_item(sets, L, area, item);
            // live 8 x 8 blocks of this tile, dealt to 2 x 4 warps: row halves (ceil | floor), column quarters with the
            // larger ones first; the second row half takes the quarters in reverse so that warps w and w + 4 (same
            // scheduler) pair a large share with a small one
            const int ra = it.pa - it.ca * it.CRa, rb = it.pb - it.cb * it.CRb;
            const int RB = ((ra < it.CRa ? ra : it.CRa) + 7) >> 3, CB = ((rb < it.CRb ? rb : it.CRb) + 7) >> 3;
            // The second-level slots are indexed by tile block and the block -> warp map depends on (RB, CB): when the tile
            // geometry changes between two items of this CTA (chunked operands, sets with different grids), a warp that is
            // already in the new item could touch a slot a slower warp still owns under the old map.  The consumer warps
            // (they walk the same item sequence) meet at a named barrier in that case; the shipped configurations (every
            // tile 13 x 13) never do.
            if (prev_geom >= 0 && prev_geom != RB * 16 + CB) asm volatile("bar.sync 1, %0;" ::"n"(kPairNCons) : "memory");
            prev_geom = RB * 16 + CB;
            const int wm = warp >> 2, wn = warp & 3;
            const int rtop = (RB + 1) >> 1;
            const int nr = wm == 0 ? rtop : RB - rtop, rblk0 = wm == 0 ? 0 : rtop;
            const int gq = wm == 0 ? wn : 3 - wn, cbase = CB >> 2, crem = CB & 3;
            const int nc = cbase + (gq < crem ? 1 : 0), cblk0 = gq * cbase + (gq < crem ? gq : crem);
            const int row0 = rblk0 * 8, col0 = cblk0 * 8;
#define CBO_PC(NR_, NC_) pair_consume<NR_, NC_>(it, smem, acc2, full, empty, stage, phase, lane, row0, col0, CB)
#define CBO_PC_ROW(NR_)                                                        \
    switch (nc) {                                                              \
        case 1: CBO_PC(NR_, 1); break;                                         \
        case 2: CBO_PC(NR_, 2); break;                                         \
        case 3: CBO_PC(NR_, 3); break;                                         \
        default: CBO_PC(NR_, 4); break;                                        \
    }
            if (nr <= 0 || nc <= 0) CBO_PC(0, 0);
            else switch (nr) {
                case 1: CBO_PC_ROW(1); break;
                case 2: CBO_PC_ROW(2); break;
                case 3: CBO_PC_ROW(3); break;
                case 4: CBO_PC_ROW(4); break;
                case 5: CBO_PC_ROW(5); break;
                case 6: CBO_PC_ROW(6); break;
                case 7: CBO_PC_ROW(7); break;
                default: CBO_PC_ROW(8); break;
            }
#undef CBO_PC_ROW
#undef CBO_PC
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
size_t pair_area_doubles(const cbo_set_desc* h_sets, int num_sets) {
    size_t t = kBK;   // the ones
    for (int s = 0; s < num_sets; ++s) {
        if (!pair_eligible(h_sets[s])) continue;
        const PairGeom g = pair_geom(h_sets[s]);
        t += (size_t)(g.szA + g.szB + g.szS);
    }
    return t;
}

long long pair_items_total(const cbo_set_desc* h_sets, int num_sets) {
    long long t = 0;
    for (int s = 0; s < num_sets; ++s) t += pair_items(h_sets[s]);
    return t;
}

int prior_pair_impl(const cbo_set_desc* h_sets, const cbo_set_desc* d_sets, int num_sets, double* area, int num_ctas,
                    cudaStream_t st) {
    CBO_CUDA(allow_dynamic_smem(prior_pair_kernel, kPairSmem));
    long long off = kBK;
    int s = 0;
    while (s < num_sets) {
        PairLaunch L;
        L.n = 0, L.total_items = 0;
        int kslabs_max = 0;
        for (; s < num_sets && L.n < kMaxPairSets; ++s) {
            const cbo_set_desc& S = h_sets[s];
            if (!pair_eligible(S)) continue;
            const PairGeom g = pair_geom(S);
            for (int k = 0; k < S.d; ++k)
                CBO_REQUIRE((long long)S.p[k] * S.n_obs_pad < 2147483647LL, "cbo_prior_eval: table %d of set %d too large", k, s);
            PairSet& E = L.e[L.n++];
            E.set = s, E.item_base = L.total_items, E.d = g.d;
            E.N = g.N, E.R = g.R, E.Rslabs = g.Rslabs, E.Kslabs = g.Kslabs;
            E.pa = g.pa, E.pb = g.pb, E.ps = g.ps;
            E.CRa = g.CRa, E.nchA = g.nchA, E.CRb = g.CRb, E.nchB = g.nchB;
            E.s_begin = g.s_begin, E.s_count = g.s_count;
            E.offA = off, off += g.szA;
            E.offB = off, off += g.szB;
            E.offS = off, off += g.szS;
            const long long items = (long long)g.s_count * g.nchA * g.nchB;
            CBO_REQUIRE(L.total_items + items < 2147483647LL, "cbo_prior_eval: too many work items");
            L.total_items += (int)items;
            if (g.Kslabs > kslabs_max) kslabs_max = g.Kslabs;
        }
        if (L.n == 0) break;
        for (int i = L.n; i < kMaxPairSets; ++i) L.e[i] = L.e[L.n - 1];
        const unsigned gx = (unsigned)((kslabs_max + 7) / 8 < 64 ? (kslabs_max + 7) / 8 : 64);
        pair_tables_kernel<<<dim3(gx, (unsigned)L.n, 3), 256, 0, st>>>(d_sets, L, area);
        note_launch();
        CBO_CUDA(cudaGetLastError());
        const int grid = L.total_items < num_ctas ? L.total_items : num_ctas;
        prior_pair_kernel<<<(unsigned)grid, kPairNT, kPairSmem, st>>>(d_sets, L, area);
        note_launch();
        CBO_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace cbo
