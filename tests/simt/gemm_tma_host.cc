// TEST INFRASTRUCTURE: the product's TMA + mbarrier warp-specialised DMMA GEMM kernel
// (gogp_b200/csrc/dgemm_tma_kernel.cuh, unmodified source) compiled for the host under the SIMT emulator.
#define GOGP_SIMT_HOST 1
#include "simt.h"

// what cuTensorMapEncodeTiled describes for this kernel: a row-major rows x k view with a 16 x 128 box
struct CUtensorMap {
    const double* ptr;
    uint64_t k, rows;
    uint64_t row_stride_bytes;
};

namespace simt {
// cp.async.bulk.tensor.2d ... mbarrier::complete_tx with CU_TENSOR_MAP_SWIZZLE_128B: box = 16 doubles (128 B)
// x 128 rows at element coordinates (c0 along k, c1 along rows); the 16-byte chunk c of row r lands at chunk
// position c ^ (r & 7).  Done at issue time -- the earliest the data can arrive.
inline void tma_load_2d_swizzle128(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    unsigned char* base = ctx.cta->dyn_aligned + dst;
    for (int r = 0; r < 128; ++r) {
        const double* src = reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(map->ptr) +
                                                            (uint64_t)(c1 + r) * map->row_stride_bytes) + c0;
        for (int kk = 0; kk < 16; ++kk) {
            const unsigned off = (unsigned)r * 128u + ((((unsigned)kk >> 1) ^ ((unsigned)r & 7u)) << 4) + (((unsigned)kk & 1u) << 3);
            *reinterpret_cast<double*>(base + off) = src[kk];
        }
    }
    mbar_complete_tx(bar, 128u * 16u * 8u);
}
}  // namespace simt

#include "../../gogp_b200/csrc/kexpr.cuh"

namespace gogp {
constexpr int TILE = 128;
enum GemmMode : int { GEMM_FULL = 0, GEMM_LOWER = 1, GEMM_KTRI = 2, GEMM_DIAG_OUT = 4, GEMM_INPLACE = 8 };  // kernels.h
#include "../../gogp_b200/csrc/dgemm_tma_kernel.cuh"
}  // namespace gogp

extern "C" void simt_dgemm_tma(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb,
                               int64_t m, int64_t n, int64_t k, double alpha, double beta, int mode, double* cdiag) {
    using namespace gogp;
    CUtensorMap mapA{A, (uint64_t)k, (uint64_t)m, (uint64_t)lda * 8}, mapB{B, (uint64_t)k, (uint64_t)n, (uint64_t)ldb * 8};
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
    simt::launch((unsigned)ntiles, THREADS, SMEM_BYTES, [&] { dgemm_tma_kernel(mapA, mapB, g); });
}
