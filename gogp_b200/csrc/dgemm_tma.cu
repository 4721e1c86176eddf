// Warp-specialised variant of the FP64 DMMA GEMM: operand tiles arrive through the
// TMA unit (cp.async.bulk.tensor.2d, 128-byte swizzle) into a 6-stage ring guarded by
// full/empty mbarriers; one producer warp issues the copies, eight consumer warps run
// mma.sync.m8n8k4.f64 and never meet at a CTA-wide barrier.
//
// Same contract as dgemm_nt_kernel (dgemm.cu): C = beta C + alpha A B^T, 128 x 128 CTA
// tile, warp tile 64 x 32, modes LOWER / KTRI / DIAG_OUT / INPLACE.
//
// Shared-memory layout of one operand stage: 128 rows x 16 doubles = 128 rows x 128 B,
// written by TMA with CU_TENSOR_MAP_SWIZZLE_128B: the 16-byte chunk c of row r sits at
// chunk position c ^ (r & 7).  A DMMA fragment thread (fr = lane / 4, fk = lane % 4)
// reads k = 8 (fk / 2) + 2 s + (fk % 2) in k-step s (any bijection of the 16 k's onto
// (fk, s) works as long as A and B agree): chunk ((fk / 2) * 4 + s) ^ fr, half fk % 2 --
// for a half-warp (4 rows x 4 fk) that is 16 distinct 8-byte bank pairs: no conflicts.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "kexpr.cuh"

namespace gogp {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int OPERAND_BYTES = BM * BK * 8;           // 16 KB
constexpr int STAGE_BYTES = 2 * OPERAND_BYTES;       // A then B
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 2 * STAGES * 8;

struct TmaArgs {
    double* C;
    double* cdiag;
    int64_t ldc;
    int tm, tn;
    int k;
    int mode;
    double alpha, beta;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}

__global__ void __launch_bounds__(THREADS, 1)
    dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                     const TmaArgs g) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B wants 1024-byte aligned tiles
    const uint32_t bars = base + STAGES * STAGE_BYTES;             // full[STAGES] then empty[STAGES]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    int ti, tj;
    if (g.mode & GEMM_LOWER) {
        lower_tile(blockIdx.x, ti, tj);
    } else {
        constexpr int GROUP_M = 16;  // grouped rasterisation, as in dgemm.cu
        const int per_group = GROUP_M * g.tn;
        const int gid = blockIdx.x / per_group, rem = blockIdx.x % per_group;
        const int first = gid * GROUP_M;
        const int gsz = (g.tm - first) < GROUP_M ? (g.tm - first) : GROUP_M;
        ti = first + rem % gsz;
        tj = rem / gsz;
    }
    const int row0 = ti * BM, col0 = tj * BN;
    const int k_lo = (g.mode & GEMM_KTRI) ? ti * BM : 0;
    const int nk = (g.k - k_lo) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8 * s), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8 * (STAGES + s)), "r"(CONSUMER_WARPS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ---- producer: one lane feeds the ring ----
        if (lane == 0) {
            for (int kt = 0; kt < nk; ++kt) {
                const int s = kt % STAGES;
                if (kt >= STAGES) mbar_wait_u32(bars + 8 * (STAGES + s), ((kt / STAGES) - 1) & 1);
                const uint32_t full = bars + 8 * s;
                mbar_expect(full, STAGE_BYTES);
                const uint32_t dst = base + s * STAGE_BYTES;
                tma_load_2d(dst, &mapA, k_lo + kt * BK, row0, full);
                tma_load_2d(dst + OPERAND_BYTES, &mapB, k_lo + kt * BK, col0, full);
            }
        }
        return;
    }

    // ---- consumers ----
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps, warp tile 64 x 32
    const int fr = lane >> 2, fk = lane & 3;
    uint32_t koff[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) koff[s] = (uint32_t)((((((fk >> 1) << 2) + s) ^ fr) << 4) | ((fk & 1) << 3));
    const uint32_t arow = (uint32_t)((wm * 64 + fr) * 128);
    const uint32_t brow = (uint32_t)(OPERAND_BYTES + (wn * 32 + fr) * 128);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % STAGES;
        mbar_wait_u32(bars + 8 * s, (kt / STAGES) & 1);
        const uint32_t st = base + s * STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[i]) : "r"(st + arow + i * 8 * 128 + koff[ks]));
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b[j]) : "r"(st + brow + j * 8 * 128 + koff[ks]));
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (STAGES + s));  // this warp is done with the stage
    }

    double* Cb;
    int64_t ldc;
    if ((g.mode & GEMM_DIAG_OUT) && ti == tj) {
        Cb = g.cdiag + (int64_t)ti * BM * BN;
        ldc = BN;
    } else {
        Cb = g.C + (int64_t)row0 * g.ldc + col0;
        ldc = g.ldc;
    }
    const int er = wm * 64 + fr, ec = wn * 32 + 2 * fk;
    if (g.mode & GEMM_INPLACE) {
        // C aliases A: every consumer warp must have read its last stage before anyone stores
        asm volatile("bar.sync 1, %0;" ::"r"(CONSUMER_WARPS * 32) : "memory");
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double2* p = reinterpret_cast<double2*>(Cb + (int64_t)(er + i * 8) * ldc + ec + j * 8);
            double2 v;
            v.x = g.alpha * acc[i][j][0];
            v.y = g.alpha * acc[i][j][1];
            if (g.beta != 0.0) {
                const double2 c = *p;
                v.x += g.beta * c.x;
                v.y += g.beta * c.y;
            }
            *p = v;
        }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
    static EncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    }
    return fn;
}

// rows x k view (row-major, leading dimension ld) -> tensor map with a 16 x 128 box
bool make_map(CUtensorMap* map, const double* ptr, int64_t ld, int64_t rows, int64_t k) {
    EncodeFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {BK, BM};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Returns false when the TMA path cannot be used (no driver entry point / encode failure):
// the caller then launches the cp.async kernel.
bool launch_dgemm_tma(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t m,
                      int64_t n, int64_t k, double alpha, double beta, int mode, double* cdiag, cudaStream_t s) {
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A, lda, m, k) || !make_map(&mapB, B, ldb, n, k)) return false;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(dgemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        configured[dev & 63] = true;
    }
    TmaArgs g;
    g.C = C;
    g.cdiag = cdiag;
    g.ldc = ldc;
    g.tm = (int)(m / BM);
    g.tn = (int)(n / BN);
    g.k = (int)k;
    g.mode = mode;
    g.alpha = alpha;
    g.beta = beta;
    const int ntiles = (mode & GEMM_LOWER) ? g.tm * (g.tm + 1) / 2 : g.tm * g.tn;
    if (ntiles <= 0) return true;
    dgemm_tma_kernel<<<ntiles, THREADS, SMEM_BYTES, s>>>(mapA, mapB, g);
    return true;
}

}  // namespace gogp
