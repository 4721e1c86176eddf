// Host-callable launchers of the sm_100a kernels (one translation unit each).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "program.h"

namespace gogp {

constexpr int TILE = 128;  // all dense blocks are 128 x 128 FP64 tiles

// ---- cov.cu ----------------------------------------------------------------
// X (N x D row-major) -> Xt ([D][Npad], zero padded)
void launch_transpose_x(const double* X, double* Xt, int64_t N, int64_t Npad, int D, cudaStream_t s);
// K = k(X, X) + noise I on the lower tiles of out (Npad x Npad, ld = Npad); the
// padding block is the identity.  gp/gp.go:109-156,220-225.
void launch_cov_build(const DevProgram& prog, const double* Xt, int64_t N, int64_t Npad, int D, double noise,
                      double* out, cudaStream_t s);
// out[m][i] = k(xa = x_i, xb = z_m), Mpad x Npad (ld = Npad), zero padded.  gp/gp.go:322-332.
void launch_cov_cross(const DevProgram& prog, const double* Xt, int64_t N, int64_t Npad, const double* Zt, int64_t M,
                      int64_t Mpad, int D, double* out, cudaStream_t s);
// Block-wise builds for the multi-GPU block-cyclic layout: a diagonal block (lower tiles, noise on
// the diagonal, identity beyond nvalid) and an off-diagonal block (full rectangle, zero padding).
// Xt / Rt / Ct already point at the block's first row / column of the dimension-major inputs.
void launch_cov_sym_block(const DevProgram& prog, const double* Xt, int64_t ldx, int64_t nvalid, int tiles, int D,
                          double noise, double* out, int64_t ld, cudaStream_t s);
void launch_cov_rect_block(const DevProgram& prog, const double* Rt, int64_t ldr, int64_t rows_valid, int rtiles,
                           const double* Ct, int64_t ldc, int64_t cols_valid, int ctiles, int D, double* out,
                           int64_t ld, cudaStream_t s);
// kss[m] = k(z_m, z_m).  gp/gp.go:270-278.
void launch_cov_self(const DevProgram& prog, const double* Zt, int64_t M, int64_t Mpad, int D, double* kss,
                     cudaStream_t s);
// Fused gradient trace, gp/gp.go:434-486 without materialising dK:
//   out[q]      = sum_{i>j} W_ij dK_q,ij + 0.5 sum_i W_ii dK_q,ii  (similarity parameters, log scale)
//   out[ntheta] = sum_i W_ii,   W = alpha alpha^T - K^-1
// Kinv: strictly-lower tiles in `kinv` (ld = Npad), diagonal tiles in `kdiag` ([T][128][128]).
// partial: scratch of ntiles * (ntheta+1) doubles.
void launch_grad_trace(const DevProgram& prog, const double* Xt, const double* alpha, const double* kinv,
                       const double* kdiag, int64_t N, int64_t Npad, int D, double* partial, double* out,
                       cudaStream_t s);
// The same trace over one rows x cols block of a distributed K^-1 (block-cyclic layout): kinv
// points at the block (ld), whose origin is element (grow0, gcol0) of the global matrix; only
// elements with global row >= global column count.  out[0..ntheta] is ACCUMULATED into.
// partial: scratch of rtiles * ctiles * (ntheta+1) doubles.
void launch_grad_trace_block(const DevProgram& prog, const double* Xt, int64_t ldx, const double* alpha,
                             const double* kinv, int64_t ld, int64_t N, int D, int64_t grow0, int rtiles,
                             int64_t gcol0, int ctiles, double* partial, double* out, cudaStream_t s);
// The same trace in ONE launch over a rank's whole local matrix of a pr x pc block-cyclic distribution (grid.hpp):
// rtiles x ctiles tiles at kinv (ld), tb tiles per distribution block, the first local block row / column being
// global block r0 / c0.  out[0..ntheta] is ACCUMULATED into; partial: rtiles * ctiles * (ntheta+1) doubles.
void launch_grad_trace_bc(const DevProgram& prog, const double* Xt, int64_t ldx, const double* alpha,
                          const double* kinv, int64_t ld, int64_t N, int D, int rtiles, int ctiles, int tb, int r0, int pr,
                          int c0, int pc, double* partial, double* out, cudaStream_t s);
// dynamic shared memory the trace kernels need for a descriptor (checked against the 227 KB opt-in limit at create)
size_t grad_trace_smem_bytes(int ndim, int ntheta);
// Input gradient for Observe's with_obs layout (gp/gp.go:118-129, 488-493):
//   gx[i*D+d] = sum_{j != i} W_ij d k(x_i, x_j)/d x_{i,d}
void launch_grad_inputs(const DevProgram& prog, const double* Xt, const double* alpha, const double* kinv,
                        const double* kdiag, int64_t N, int64_t Npad, int D, double* gx, cudaStream_t s);

// ---- dgemm.cu ----------------------------------------------------------------
enum GemmMode : int {
    GEMM_FULL = 0,
    GEMM_LOWER = 1,     // only tiles ti >= tj (SYRK / LAUUM)
    GEMM_KTRI = 2,      // A is upper triangular w.r.t. its own origin: k starts at ti*128
    GEMM_DIAG_OUT = 4,  // diagonal tiles go to cdiag ([T][128][128]) instead of C
    GEMM_INPLACE = 8,   // C aliases A (n == 128): one CTA must own entire rows of the block
};
// Block-cyclic tile mask (grid.hpp: C is a window of one rank's local matrix of a Pr x Pc block-cyclic
// distribution with tb x tb tiles per distribution block).  Tile (ti, tj) of C belongs to global block
// (I, J) = (r0 + pr * (ti / tb), c0 + pc * (tj / tb)); it is computed only when it meets the lower triangle of
// the GLOBAL matrix: J < I, or J == I and the tile is on or below the block's diagonal.  tb == 0: no mask.
struct GemmMask {
    int tb = 0;
    int r0 = 0, pr = 1, c0 = 0, pc = 1;
};
// C[i][j] = beta*C[i][j] + alpha * sum_k A[i][k] B[j][k]; all of m, n, k multiples of 128.
void launch_dgemm_nt(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t m,
                     int64_t n, int64_t k, double alpha, double beta, int mode, double* cdiag, cudaStream_t s,
                     const GemmMask* mask = nullptr);
// TMA + mbarrier warp-specialised variant (dgemm_tma.cu); false: not available, use launch_dgemm_nt's kernels
bool launch_dgemm_tma(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t m,
                      int64_t n, int64_t k, double alpha, double beta, int mode, double* cdiag, cudaStream_t s,
                      const GemmMask* mask = nullptr);
void set_gemm_config(int cfg);  // 0: 128x128 tiles, 1 CTA/SM; 1: 128x64 tiles, 2 CTAs/SM
// register-only issue-rate microbenchmarks: which = 0 DMMA m8n8k4, 1 DFMA
int fp64_peak_variants();
void launch_fp64_peak(int which, int variant, int iters, double* sink, cudaStream_t s);
double fp64_peak_flops_per_launch(int which, int variant, int iters, int nsm);

// ---- leaf.cu -------------------------------------------------------------------
// Cholesky of the 128x128 diagonal tile at A (ld), in place (lower; upper zeroed),
// and its inverse into winv (128x128 row-major lower).  info: 0 or 1-based index
// of the first non-positive pivot (global index = base + local).
// variant: -1 the process default (GOGP_LEAF, else the shipped blocked kernel); 0/1/2 select a kernel for timing runs.
void launch_potrf_leaf(double* A, int64_t ld, double* winv, int* info, int base, cudaStream_t s, int variant = -1);
// dst tile (ld) = winv^T (upper triangular, strictly-lower zeroed).
void launch_trtri_leaf(const double* winv, double* dst, int64_t ld, cudaStream_t s);
// Blocked triangular solves with the tile inverses.  fwd: out = L^-1 rhs; bwd: out = L^-T rhs.
// rhs (Npad) may be destroyed; out must not alias it.  sync: Npad/128 + 1 unsigned ints of device
// scratch for the single-launch kernels (NULL: one launch per block).
void launch_trsv_lower(const double* L, int64_t ld, const double* winv, double* rhs, double* out, int64_t Npad,
                       bool transposed, cudaStream_t s, int64_t* launches, unsigned* sync);
// dst[r][c] = src[r][c] for a rows x cols block (cols even, 16-byte aligned rows)
void launch_copy_block(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols,
                       cudaStream_t s);
// y += a x
void launch_axpy(double* y, const double* x, double a, int64_t n, cudaStream_t s);
// v[i] = value for i in [0, n)
void launch_fill(double* v, int64_t n, double value, cudaStream_t s);
void launch_fill_diag(double* v, int64_t stride, int n, double value, cudaStream_t s);
// pseudo-random fill in (-0.5, 0.5) for microbenchmarks
void launch_fill_pattern(double* v, int64_t n, cudaStream_t s);
// out[0] = sum_i log L_ii (i < N), out[1] = sum_i y_i alpha_i, out[2] / out[3] = min / max L_ii
void launch_logdet_dot(const double* L, int64_t ld, const double* y, const double* alpha, int64_t N, double* out,
                       cudaStream_t s);
// out = L v (transposed: L^T v) with the lower triangle of L, first N rows / columns (condition estimate)
void launch_trmv_lower(const double* L, int64_t ld, const double* v, double* out, int64_t N, bool transposed,
                       cudaStream_t s);
// winv[t] = (diagonal tile t of L)^-1 for t < tiles (gogp_set_state)
void launch_tile_inverse(const double* L, int64_t ld, double* winv, int tiles, cudaStream_t s);
// row reductions of a Mpad x Npad matrix: out[m] = sum_i B[m][i] * (v ? v[i] : B[m][i])
void launch_row_reduce(const double* B, int64_t ld, int64_t rows, int64_t cols, const double* v, double* out,
                       cudaStream_t s);
// mirror helper for debug fetches: out (N x N) from lower tiles (+ optional diagonal tiles)
void launch_gather_sym(const double* src, int64_t ld, const double* diag, int64_t N, double* out, cudaStream_t s);

}  // namespace gogp
