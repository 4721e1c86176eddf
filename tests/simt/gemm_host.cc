// TEST INFRASTRUCTURE: the product's cp.async DMMA GEMM kernel (gogp_b200/csrc/dgemm_kernels.cuh, unmodified
// source, both shipped CTA shapes, every tile-map mode) compiled for the host under the SIMT emulator.
#define GOGP_SIMT_HOST 1
#include "simt.h"

#include "../../gogp_b200/csrc/kexpr.cuh"

namespace gogp {
constexpr int TILE = 128;
enum GemmMode : int { GEMM_FULL = 0, GEMM_LOWER = 1, GEMM_KTRI = 2, GEMM_DIAG_OUT = 4, GEMM_INPLACE = 8 };  // kernels.h
#include "../../gogp_b200/csrc/dgemm_kernels.cuh"

template <int WM, int WN, int STAGES, int MINB, int FM = 8>
void run(GemmArgs g, int64_t m, int64_t n) {
    constexpr int BM = 8 * FM * WM, BN = 32 * WN;
    const size_t smem = (size_t)STAGES * (BM + BN) * PITCH * sizeof(double);
    g.tm = (int)(m / BM);
    g.tn = (int)(n / BN);
    const int ratio = BM / BN;
    const int ntiles = (g.mode & GEMM_LOWER) ? ratio * g.tm * (g.tm + 1) / 2 : g.tm * g.tn;
    simt::launch((unsigned)ntiles, WM * WN * 32, smem, [&] { dgemm_nt_kernel<WM, WN, STAGES, MINB, FM>(g); });
}
}  // namespace gogp

// C = beta C + alpha A B^T with the kernel's modes; shape 0: <2,4,4,1> (128 x 128, in-place capable), 1: <2,2,3,2>,
// 2: <2,2,3,3,4> (64 x 64, 32 x 32 warp tiles), 3: <1,4,3,2,4> (32 x 128, in-place capable).
extern "C" void simt_dgemm(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t m,
                           int64_t n, int64_t k, double alpha, double beta, int mode, double* cdiag, int shape) {
    gogp::GemmArgs g;
    g.C = C;
    g.A = A;
    g.B = B;
    g.cdiag = cdiag;
    g.ldc = ldc;
    g.lda = lda;
    g.ldb = ldb;
    g.tm = g.tn = 0;
    g.k = (int)k;
    g.mode = mode;
    g.alpha = alpha;
    g.beta = beta;
    if (shape == 0)
        gogp::run<2, 4, 4, 1>(g, m, n);
    else if (shape == 1)
        gogp::run<2, 2, 3, 2>(g, m, n);
    else if (shape == 2)
        gogp::run<2, 2, 3, 3, 4>(g, m, n);  // latency shapes: 64 x 64 ...
    else
        gogp::run<1, 4, 3, 2, 4>(g, m, n);  // ... and 32 x 128 (in place)
}
