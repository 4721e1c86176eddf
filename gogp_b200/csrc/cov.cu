// Covariance-side kernels: K(X,X)+noise build, K(X,Z) cross build, prior
// variance, the fused gradient trace and the input gradient.  The kernels themselves are in
// cov_kernels.cuh (shared with the CPU test tier); this file holds their launchers.
//
// Reference semantics: gp/gp.go:109-156 (cov closure), :220-225 (K.SetSym over
// j >= i), :270-278 and :322-332 (Produce), :434-486 (Gradient).  HBM-bound by
// design: one FP64 store (build) or one FP64 load (trace) per matrix element;
// the inputs are dimension-major ([D][Npad]) so a tile's coordinates are D
// contiguous 1 KB runs, staged into shared memory by TMA bulk copies.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "kexpr.cuh"

namespace gogp {

#include "cov_kernels.cuh"

void launch_transpose_x(const double* X, double* Xt, int64_t N, int64_t Npad, int D, cudaStream_t s) {
    int threads = 256;
    int blocks = (int)((Npad + threads - 1) / threads);
    transpose_x_kernel<<<blocks, threads, 0, s>>>(X, Xt, N, Npad, D);
}

static size_t cov_smem(int D) { return (size_t)2 * D * TILE * sizeof(double) + 16; }

// ---- specialised element loop ("fast shape", cov_tile_fast_kernel in cov_kernels.cuh): dispatch ----
static int fast_elem_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GOGP_ELEM_FAST");
        v = e ? atoi(e) : 1;
    }
    return v;
}
// NN the fast kernels are instantiated for, or -1
static int fast_shape(const DevProgram& prog) {
    if (!fast_elem_enabled() || prog.nterms != 1) return -1;
    const int nn = prog.nnorm[0], rest = prog.fbeg[1] - prog.fbeg[0] - nn;
    if (rest < 0 || rest > kFastMaxRest) return -1;
    if (nn == 0 || nn == 1 || nn == 2 || nn == 3 || nn == 4 || nn == 8) return nn;
    return -1;
}

template <bool SYM>
static bool launch_cov_fast(int ntiles, size_t smem, cudaStream_t s, const DevProgram& prog, const double* Rt,
                            int64_t ldr, int64_t nrows, const double* Ct, int64_t ldc, int64_t ncols, int D,
                            double noise, double* out, int64_t ld, int tiles_n) {
    const int nn = fast_shape(prog);
    if (nn < 0 || ntiles <= 0) return false;
#define GOGP_COV_FAST(NNV)                                                                                        \
    case NNV:                                                                                                     \
        cudaFuncSetAttribute(cov_tile_fast_kernel<NNV, SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                             (int)smem);                                                                          \
        cov_tile_fast_kernel<NNV, SYM><<<ntiles, 256, smem, s>>>(prog, Rt, ldr, nrows, Ct, ldc, ncols, D, noise,  \
                                                                 out, ld, tiles_n);                               \
        return true;
    switch (nn) {
        GOGP_COV_FAST(0)
        GOGP_COV_FAST(1)
        GOGP_COV_FAST(2)
        GOGP_COV_FAST(3)
        GOGP_COV_FAST(4)
        GOGP_COV_FAST(8)
    }
#undef GOGP_COV_FAST
    return false;
}

void launch_cov_build(const DevProgram& prog, const double* Xt, int64_t N, int64_t Npad, int D, double noise,
                      double* out, cudaStream_t s) {
    int T = (int)(Npad / TILE);
    int ntiles = T * (T + 1) / 2;
    size_t smem = cov_smem(D);
    if (launch_cov_fast<true>(ntiles, smem, s, prog, Xt, Npad, N, Xt, Npad, N, D, noise, out, Npad, T)) return;
    cudaFuncSetAttribute(cov_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cov_tile_kernel<true><<<ntiles, 256, smem, s>>>(prog, Xt, Npad, N, Xt, Npad, N, D, noise, out, Npad, T);
}

void launch_cov_sym_block(const DevProgram& prog, const double* Xt, int64_t ldx, int64_t nvalid, int tiles, int D,
                          double noise, double* out, int64_t ld, cudaStream_t s) {
    size_t smem = cov_smem(D);
    if (launch_cov_fast<true>(tiles * (tiles + 1) / 2, smem, s, prog, Xt, ldx, nvalid, Xt, ldx, nvalid, D, noise, out, ld,
                              tiles))
        return;
    cudaFuncSetAttribute(cov_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cov_tile_kernel<true><<<tiles * (tiles + 1) / 2, 256, smem, s>>>(prog, Xt, ldx, nvalid, Xt, ldx, nvalid, D, noise, out,
                                                                   ld, tiles);
}

void launch_cov_rect_block(const DevProgram& prog, const double* Rt, int64_t ldr, int64_t rows_valid, int rtiles,
                           const double* Ct, int64_t ldc, int64_t cols_valid, int ctiles, int D, double* out,
                           int64_t ld, cudaStream_t s) {
    size_t smem = cov_smem(D);
    if (launch_cov_fast<false>(rtiles * ctiles, smem, s, prog, Rt, ldr, rows_valid, Ct, ldc, cols_valid, D, 0.0, out, ld,
                               ctiles))
        return;
    cudaFuncSetAttribute(cov_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cov_tile_kernel<false><<<rtiles * ctiles, 256, smem, s>>>(prog, Rt, ldr, rows_valid, Ct, ldc, cols_valid, D, 0.0, out,
                                                             ld, ctiles);
}

void launch_cov_cross(const DevProgram& prog, const double* Xt, int64_t N, int64_t Npad, const double* Zt, int64_t M,
                      int64_t Mpad, int D, double* out, cudaStream_t s) {
    int tm = (int)(Mpad / TILE), tn = (int)(Npad / TILE);
    size_t smem = cov_smem(D);
    if (launch_cov_fast<false>(tm * tn, smem, s, prog, Zt, Mpad, M, Xt, Npad, N, D, 0.0, out, Npad, tn)) return;
    cudaFuncSetAttribute(cov_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cov_tile_kernel<false><<<tm * tn, 256, smem, s>>>(prog, Zt, Mpad, M, Xt, Npad, N, D, 0.0, out, Npad, tn);
}

void launch_cov_self(const DevProgram& prog, const double* Zt, int64_t M, int64_t Mpad, int D, double* kss,
                     cudaStream_t s) {
    (void)D;
    int threads = 128;
    int blocks = (int)((M + threads - 1) / threads);
    if (blocks > 0) cov_self_kernel<<<blocks, threads, 0, s>>>(prog, Zt, M, Mpad, kss);
}

// ---- fused gradient trace: launchers (kernel and its description in cov_kernels.cuh) ----
void launch_grad_trace(const DevProgram& prog, const double* Xt, const double* alpha, const double* kinv,
                       const double* kdiag, int64_t N, int64_t Npad, int D, double* partial, double* out,
                       cudaStream_t s) {
    int T = (int)(Npad / TILE);
    int ntiles = T * (T + 1) / 2;
    size_t smem = (size_t)2 * D * TILE * sizeof(double) + 16 + (size_t)2 * GT_E * 256 * sizeof(double) +
                  (size_t)(prog.ntheta + 1) * 256 * sizeof(double);
    cudaFuncSetAttribute(grad_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    grad_trace_kernel<<<ntiles, 256, smem, s>>>(prog, Xt, Npad, alpha, kinv, Npad, kdiag, N, D, partial, 0, 0, 0);
    grad_reduce_kernel<<<prog.ntheta + 1, 256, 0, s>>>(partial, ntiles, prog.ntheta + 1, out, 0);
}

void launch_grad_trace_block(const DevProgram& prog, const double* Xt, int64_t ldx, const double* alpha,
                             const double* kinv, int64_t ld, int64_t N, int D, int64_t grow0, int rtiles,
                             int64_t gcol0, int ctiles, double* partial, double* out, cudaStream_t s) {
    const int ntiles = rtiles * ctiles;
    if (ntiles <= 0) return;
    size_t smem = (size_t)2 * D * TILE * sizeof(double) + 16 + (size_t)2 * GT_E * 256 * sizeof(double) +
                  (size_t)(prog.ntheta + 1) * 256 * sizeof(double);
    cudaFuncSetAttribute(grad_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    grad_trace_kernel<<<ntiles, 256, smem, s>>>(prog, Xt, ldx, alpha, kinv, ld, nullptr, N, D, partial, ctiles, grow0,
                                                gcol0);
    grad_reduce_kernel<<<prog.ntheta + 1, 256, 0, s>>>(partial, ntiles, prog.ntheta + 1, out, 1);
}

void launch_grad_inputs(const DevProgram& prog, const double* Xt, const double* alpha, const double* kinv,
                        const double* kdiag, int64_t N, int64_t Npad, int D, double* gx, cudaStream_t s) {
    if (N <= 0) return;
    grad_inputs_kernel<<<(int)N, 256, 0, s>>>(prog, Xt, Npad, alpha, kinv, Npad, kdiag, N, D, gx);
}

}  // namespace gogp
