// Persistent, warp-specialised 128 x 128 FP64 tile products  acc = sum_k A[m][k] B[n][k]  on row-major operands, fed by the
// tensor-map TMA (cp.async.bulk.tensor.2d, SASS UTMALDG) -- the K1b pipeline (prior_eval.cu) for operands that are NOT stored
// in the blocked fragment order: a tensor map over the plain row-major matrix with a 4-column x 128-row box lands one k4-group
// of a slab as [row][4] doubles, which IS the fragment order of dmma_tile.cuh.  One pipeline stage (16 deep) = 4 boxes per
// operand = 32 KB, six stages in flight.  Roles: 8 consumer warps (2 x 4, 64 x 32 each: 128 accumulator registers, 232 after
// setmaxnreg) and a producer warpgroup whose elected thread walks the same tile list one ring ahead, so that the epilogue of a
// tile (read-modify-write of C from the fragments) overlaps the loads of the next one.  Tiles are dealt round-robin over a
// grid of one CTA per SM (no counter: both roles derive the same sequence from blockIdx).
//
// Used by the trailing updates of the blocked Cholesky / triangular inverse and by Ky^-1 (K5, obs_gp_fit.cu) and by the
// prior-precompute SYRK (K1a, prior_precompute.cu).
#pragma once
#include <cuda.h>

#include "dmma_tile.cuh"

namespace cbo {

constexpr int kTmaStages = 6;
constexpr int kTmaTile = 128;
constexpr int kTmaSlab = kTmaTile * kBK;                       // doubles per operand per stage
constexpr unsigned kTmaStageBytes = 2 * kTmaSlab * sizeof(double);
constexpr int kTmaCons = 256, kTmaThreads = kTmaCons + 128;
constexpr size_t kTmaSmem = (size_t)kTmaStages * kTmaStageBytes + 2 * kTmaStages * sizeof(uint64_t) + 128;
constexpr int kTmaProdRegs = 40, kTmaConsRegs = 232;
static_assert(128 * kTmaProdRegs + kTmaCons * kTmaConsRegs <= kTmaThreads * ((65536 / kTmaThreads) / 8 * 8), "setmaxnreg budget");

// One tile of work: rows of the two operands, first column and depth (in 16-deep slabs) of the product.
struct TileJob {
    int rowA, rowB, col0, nk;
    int ti, tj;               // what the epilogue needs to find its output
};

// ---- host: tensor map of a row-major FP64 matrix (rows x cols, row pitch ld doubles) with the 4 x 128 box -----------------
inline int make_f64_rowmajor_map(CUtensorMap* map, const double* base, size_t rows, size_t cols, size_t ld) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return reinterpret_cast<encode_fn>(fn);
    }();
    CBO_REQUIRE(encode != nullptr, "the driver does not export cuTensorMapEncodeTiled");
    CBO_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld % 2 == 0 && rows > 0 && cols > 0,
                "tensor map: base must be 16-byte aligned and the row pitch even");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {4, (cuuint32_t)kTmaTile};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CBO_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rows %zu, cols %zu, pitch %zu)", (int)r, rows, cols, ld);
    return 0;
}

#ifdef __CUDACC__

__device__ __forceinline__ void tma_box_g2s(void* smem_dst, const CUtensorMap* map, int col, int row, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(row), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// acc fragment -> (row, column) of the 128 x 128 tile (consumer thread tid < 256)
template <class F>
__device__ __forceinline__ void tma_for_each_acc(const double (&acc)[8][4][2], int tid, F&& f) {
    const int lane = tid & 31, warp = tid >> 5;
    const int row0 = (warp / 4) * 64, col0 = (warp % 4) * 32;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        const int r = row0 + mi * 8 + (lane >> 2);
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) f(r, col0 + ni * 8 + (lane & 3) * 2, acc[mi][ni][0], acc[mi][ni][1]);
    }
}

// PLAN:  int count() const;  TileJob job(int t) const;  void store(const TileJob&, const double (&acc)[8][4][2], int tid) const;
template <class PLAN>
__global__ void __launch_bounds__(kTmaThreads, 1)
tma_tile_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const PLAN plan) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    // [stage][A slab | B slab]; the TMA wants 128-byte aligned destinations (the allocation carries 128 bytes of slack)
    double* ring = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(tma_smem) + 127) & ~(uintptr_t)127);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)kTmaStages * 2 * kTmaSlab);
    uint64_t* empty = full + kTmaStages;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < kTmaStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kTmaCons / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int count = plan.count();
    int stage = 0;
    unsigned phase = 0;
    if (warp >= kTmaCons / 32) {
        setmaxnreg_dec<kTmaProdRegs>();
        if (warp == kTmaCons / 32 && lane == 0) {
            tma_prefetch_map(&mapA);
            tma_prefetch_map(&mapB);
#pragma unroll 1
            for (int t = blockIdx.x; t < count; t += gridDim.x) {
                const TileJob j = plan.job(t);
#pragma unroll 1
                for (int kt = 0; kt < j.nk; ++kt) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&full[stage], kTmaStageBytes);
                    double* sA = ring + (size_t)stage * 2 * kTmaSlab;
                    const int c = j.col0 + kt * kBK;
#pragma unroll
                    for (int kb = 0; kb < kBK / 4; ++kb) {
                        tma_box_g2s(sA + kb * kTmaTile * 4, &mapA, c + kb * 4, j.rowA, &full[stage]);
                        tma_box_g2s(sA + kTmaSlab + kb * kTmaTile * 4, &mapB, c + kb * 4, j.rowB, &full[stage]);
                    }
                    if (++stage == kTmaStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        setmaxnreg_inc<kTmaConsRegs>();
        const int row0 = (warp / 4) * 64, col0 = (warp % 4) * 32;
#pragma unroll 1
        for (int t = blockIdx.x; t < count; t += gridDim.x) {
            const TileJob j = plan.job(t);
            double acc[8][4][2];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll 1
            for (int kt = 0; kt < j.nk; ++kt) {
                mbar_wait(&full[stage], phase);
                const double* sA = ring + (size_t)stage * 2 * kTmaSlab;
                mma_stage<kTmaTile, kTmaTile, 8, 4>(sA, sA + kTmaSlab, acc, row0, col0, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == kTmaStages) { stage = 0; phase ^= 1u; }
            }
            plan.store(j, acc, tid);
        }
    }
}

// grid = min(tiles, SMs); dynamic shared memory opt-in included
template <class PLAN>
inline cudaError_t launch_tma_tiles(const CUtensorMap& mapA, const CUtensorMap& mapB, const PLAN& plan, int tiles, int sms, cudaStream_t st) {
    if (tiles <= 0) return cudaSuccess;
    cudaError_t e = allow_dynamic_smem(tma_tile_kernel<PLAN>, kTmaSmem);
    if (e != cudaSuccess) return e;
    tma_tile_kernel<PLAN><<<tiles < sms ? tiles : sms, kTmaThreads, kTmaSmem, st>>>(mapA, mapB, plan);
    note_launch();
    return cudaGetLastError();
}

#endif  // __CUDACC__

}  // namespace cbo
