// Covariance-side kernels: K(X,X)+noise build, K(X,Z) cross build, prior
// variance, the fused gradient trace and the input gradient.  The kernels themselves are in
// cov_kernels.cuh (shared with the CPU test tier); this file holds their launchers.
//
// Reference semantics: gp/gp.go:109-156 (cov closure), :220-225 (K.SetSym over
// j >= i), :270-278 and :322-332 (Produce), :434-486 (Gradient).  HBM-bound by
// design: one FP64 store (build) or one FP64 load (trace) per matrix element;
// the inputs are dimension-major ([D][Npad]) so a tile's coordinates are D
// contiguous 1 KB runs, staged into shared memory by TMA bulk copies.
#include <atomic>
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
static int fast_elem_env() {
    const char* e = getenv("GOGP_ELEM_FAST");
    return e ? atoi(e) : 1;
}
static int fast_elem_enabled() {
    static const int v = fast_elem_env();  // initialised once, thread-safely
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

// ---- fused gradient trace: launchers (kernels and their description in cov_kernels.cuh) ----
static size_t trace_smem_generic(const DevProgram& prog, int D) {
    return (size_t)2 * D * TILE * sizeof(double) + 16 + (size_t)2 * GT_E * 256 * sizeof(double) +
           (size_t)(prog.ntheta + 1) * 256 * sizeof(double);
}
static size_t trace_smem_fast(const DevProgram& prog, int D) {
    return (size_t)2 * D * TILE * sizeof(double) + 16 + (size_t)(prog.ntheta + 1) * 256 * sizeof(double);
}
// Shared memory the trace needs for this descriptor, against the 227 KB a CTA can opt into (checked once, at
// gogp_create: ndim <= 64 and ntheta <= 32 together could ask for more).
size_t grad_trace_smem_bytes(int ndim, int ntheta) {
    return (size_t)2 * ndim * TILE * sizeof(double) + 16 + (size_t)2 * GT_E * 256 * sizeof(double) +
           (size_t)(ntheta + 1) * 256 * sizeof(double);
}

// The specialised trace's instantiation for a program: NN unrolled Normal factors (one term) and room for MAXR other
// leaves per term besides bare parameters; false: the interpreter.
static bool trace_fast_shape(const DevProgram& prog, int* nn_out, int* maxr_out) {
    if (!fast_elem_enabled()) return false;
    static const int inst[] = {8, 4, 3, 2, 1, 0};
    for (int nn : inst) {
        if (nn > 0 && (prog.nterms != 1 || prog.nnorm[0] < nn)) continue;
        int worst = 0;
        for (int t = 0; t < prog.nterms; ++t) {
            int rest = 0;
            for (int fi = prog.fbeg[t] + nn; fi < prog.fbeg[t + 1]; ++fi)
                if (prog.f[fi].kind != F_PARAM) ++rest;
            if (rest > worst) worst = rest;
        }
        if (worst > kTraceMaxRest) continue;
        *nn_out = nn;
        *maxr_out = worst == 0 ? 0 : (worst == 1 ? 1 : kTraceMaxRest);
        return true;
    }
    return false;
}

template <int NN, int MAXR, int MINB>
static void launch_trace_fast_inst(int ntiles, size_t smem, cudaStream_t s, const DevProgram& prog, const double* Xt,
                                   int64_t ldx, const double* alpha, const double* kinv, int64_t ld,
                                   const double* kdiag, int64_t N, int D, double* partial, const TraceMap& map) {
    static std::atomic<size_t> configured[64];  // largest size the attribute was set to, per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured[dev & 63].load(std::memory_order_acquire) < smem) {
        cudaFuncSetAttribute(grad_trace_fast_kernel<NN, MAXR, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
        configured[dev & 63].store(smem, std::memory_order_release);
    }
    grad_trace_fast_kernel<NN, MAXR, MINB><<<ntiles, 256, smem, s>>>(prog, Xt, ldx, alpha, kinv, ld, kdiag, N, D, partial,
                                                                      map);
}

static void launch_trace_any(int ntiles, cudaStream_t s, const DevProgram& prog, const double* Xt, int64_t ldx,
                             const double* alpha, const double* kinv, int64_t ld, const double* kdiag, int64_t N, int D,
                             double* partial, const TraceMap& map) {
    int nn = -1, maxr = 0;
    const size_t fs = trace_smem_fast(prog, D);
    if (trace_fast_shape(prog, &nn, &maxr)) {
#define GOGP_TRACE_FAST(NNV, MR, MINB)                                                                                 \
    if (nn == NNV && maxr == MR) {                                                                                     \
        launch_trace_fast_inst<NNV, MR, MINB>(ntiles, fs, s, prog, Xt, ldx, alpha, kinv, ld, kdiag, N, D, partial, map); \
        return;                                                                                                        \
    }
        // CTAs per SM by register budget: 4 -> 64, 3 -> 80, 2 -> 128 registers per thread
        GOGP_TRACE_FAST(0, 0, 4) GOGP_TRACE_FAST(0, 1, 3) GOGP_TRACE_FAST(0, 4, 2)
        GOGP_TRACE_FAST(1, 0, 4) GOGP_TRACE_FAST(1, 1, 3) GOGP_TRACE_FAST(1, 4, 2)
        GOGP_TRACE_FAST(2, 0, 3) GOGP_TRACE_FAST(2, 1, 3) GOGP_TRACE_FAST(2, 4, 2)
        GOGP_TRACE_FAST(3, 0, 3) GOGP_TRACE_FAST(3, 1, 2) GOGP_TRACE_FAST(3, 4, 2)
        GOGP_TRACE_FAST(4, 0, 3) GOGP_TRACE_FAST(4, 1, 2) GOGP_TRACE_FAST(4, 4, 2)
        GOGP_TRACE_FAST(8, 0, 2) GOGP_TRACE_FAST(8, 1, 2) GOGP_TRACE_FAST(8, 4, 2)
#undef GOGP_TRACE_FAST
    }
    const size_t smem = trace_smem_generic(prog, D);
    static std::atomic<size_t> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured[dev & 63].load(std::memory_order_acquire) < smem) {
        cudaFuncSetAttribute(grad_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63].store(smem, std::memory_order_release);
    }
    grad_trace_kernel<<<ntiles, 256, smem, s>>>(prog, Xt, ldx, alpha, kinv, ld, kdiag, N, D, partial, map);
}

void launch_grad_trace(const DevProgram& prog, const double* Xt, const double* alpha, const double* kinv,
                       const double* kdiag, int64_t N, int64_t Npad, int D, double* partial, double* out,
                       cudaStream_t s) {
    int T = (int)(Npad / TILE);
    int ntiles = T * (T + 1) / 2;
    TraceMap map{};
    map.mode = 0;
    launch_trace_any(ntiles, s, prog, Xt, Npad, alpha, kinv, Npad, kdiag, N, D, partial, map);
    grad_reduce_kernel<<<prog.ntheta + 1, 256, 0, s>>>(partial, ntiles, prog.ntheta + 1, out, 0);
}

void launch_grad_trace_block(const DevProgram& prog, const double* Xt, int64_t ldx, const double* alpha,
                             const double* kinv, int64_t ld, int64_t N, int D, int64_t grow0, int rtiles,
                             int64_t gcol0, int ctiles, double* partial, double* out, cudaStream_t s) {
    const int ntiles = rtiles * ctiles;
    if (ntiles <= 0) return;
    TraceMap map{};
    map.mode = 1;
    map.ctiles = ctiles;
    map.grow0 = grow0;
    map.gcol0 = gcol0;
    launch_trace_any(ntiles, s, prog, Xt, ldx, alpha, kinv, ld, nullptr, N, D, partial, map);
    grad_reduce_kernel<<<prog.ntheta + 1, 256, 0, s>>>(partial, ntiles, prog.ntheta + 1, out, 1);
}

void launch_grad_trace_bc(const DevProgram& prog, const double* Xt, int64_t ldx, const double* alpha,
                          const double* kinv, int64_t ld, int64_t N, int D, int rtiles, int ctiles, int tb, int r0, int pr,
                          int c0, int pc, double* partial, double* out, cudaStream_t s) {
    const int ntiles = rtiles * ctiles;
    if (ntiles <= 0) return;
    TraceMap map{};
    map.mode = 2;
    map.ctiles = ctiles;
    map.tb = tb;
    map.r0 = r0;
    map.pr = pr;
    map.c0 = c0;
    map.pc = pc;
    launch_trace_any(ntiles, s, prog, Xt, ldx, alpha, kinv, ld, nullptr, N, D, partial, map);
    grad_reduce_kernel<<<prog.ntheta + 1, 256, 0, s>>>(partial, ntiles, prog.ntheta + 1, out, 1);
}

void launch_grad_inputs(const DevProgram& prog, const double* Xt, const double* alpha, const double* kinv,
                        const double* kdiag, int64_t N, int64_t Npad, int D, double* gx, cudaStream_t s) {
    if (N <= 0) return;
    grad_inputs_kernel<<<(int)N, 256, 0, s>>>(prog, Xt, Npad, alpha, kinv, Npad, kdiag, N, D, gx);
}

}  // namespace gogp
