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

#include <atomic>

#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "kexpr.cuh"

namespace gogp {

namespace {

#include "dgemm_tma_kernel.cuh"

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
                      int64_t n, int64_t k, double alpha, double beta, int mode, double* cdiag, cudaStream_t s,
                      const GemmMask* mask) {
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A, lda, m, k) || !make_map(&mapB, B, ldb, n, k)) return false;
    static std::atomic<bool> configured[64];  // the attribute is per device; handles on other threads may race here
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63].load(std::memory_order_acquire)) {
        cudaFuncSetAttribute(dgemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        configured[dev & 63].store(true, std::memory_order_release);
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
    if (mask && mask->tb > 0) {
        g.mtb = mask->tb;
        g.mr0 = mask->r0;
        g.mpr = mask->pr;
        g.mc0 = mask->c0;
        g.mpc = mask->pc;
    }
    const int ntiles = (mode & GEMM_LOWER) ? g.tm * (g.tm + 1) / 2 : g.tm * g.tn;
    if (ntiles <= 0) return true;
    dgemm_tma_kernel<<<ntiles, THREADS, SMEM_BYTES, s>>>(mapA, mapB, g);
    return true;
}

}  // namespace gogp
