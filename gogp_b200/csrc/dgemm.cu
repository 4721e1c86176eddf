// FP64 tensor-core GEMM  C = beta*C + alpha * A * B^T  for sm_100a.
//
// This is the dense contraction behind every O(N^3) step of the path: the
// trailing SYRK/GEMM updates of the Cholesky (reference: gp.L.Factorize(K),
// gp/gp.go:228), the triangular inverse and the K^-1 = L^-T L^-1 product that
// replace the reference's per-parameter L.SolveTo (gp/gp.go:454,480), and the
// multi-right-hand-side solve of Produce (gp/gp.go:337-340).
//
// Blackwell's tcgen05/TMEM path has no FP64 kind (ptxas rejects .kind::f64), so
// the FP64 tensor pipe is reached through mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).
// CTA tile 128x128, BK = 16, 4-stage cp.async ring (160 KB smem), 8 warps as
// 2 (M) x 4 (N), warp tile 64x32 = 8x4 DMMA fragments (64 FP64 accumulators per
// thread).  Both operands are row-major with k contiguous ("NT"), staged as
// [row][k] with a row pitch of 20 doubles: a half-warp's fragment read touches
// rows r..r+3 x k..k+3 -> word offsets (20r + k)*2, all 32 banks distinct.
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "kexpr.cuh"

namespace gogp {

namespace {

#include "dgemm_kernels.cuh"

template <int WM, int WN, int STAGES, int MINB, int FM = 8>
void launch_cfg(const GemmArgs& g0, int64_t m, int64_t n, cudaStream_t s) {
    constexpr int BM = 8 * FM * WM, BN = 32 * WN;
    constexpr size_t SMEM = (size_t)STAGES * (BM + BN) * PITCH * sizeof(double);
    static std::atomic<bool> configured[64];  // the attribute is per device; handles on other threads may race here
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63].load(std::memory_order_acquire)) {
        cudaFuncSetAttribute(dgemm_nt_kernel<WM, WN, STAGES, MINB, FM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SMEM);
        configured[dev & 63].store(true, std::memory_order_release);
    }
    GemmArgs g = g0;
    g.tm = (int)(m / BM);
    g.tn = (int)(n / BN);
    const int ratio = BM / BN;
    const int ntiles = (g.mode & GEMM_LOWER) ? ratio * g.tm * (g.tm + 1) / 2 : g.tm * g.tn;
    if (ntiles <= 0) return;
    dgemm_nt_kernel<WM, WN, STAGES, MINB, FM><<<ntiles, WM * WN * 32, SMEM, s>>>(g);
}

// ---- FP64 peak microbenchmarks (registers only) --------------------------------------
template <int NACC>
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* sink) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double* sink) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x * 1e-9 + i;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
}

}  // namespace

static int g_gemm_cfg = -1;  // 0: 128x128 1 CTA/SM, 1: 128x64 2 CTAs/SM, 2: TMA + mbarrier warp-specialised
                             // (env GOGP_GEMM_CFG, default 1)
void set_gemm_config(int cfg) { g_gemm_cfg = cfg; }

void launch_dgemm_nt(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t m,
                     int64_t n, int64_t k, double alpha, double beta, int mode, double* cdiag, cudaStream_t s,
                     const GemmMask* mask) {
    static int64_t tma_min_k = 0;
    static std::once_flag knobs;  // handles on different host threads (one per GPU) come through here concurrently
    std::call_once(knobs, [] {
        const char* mk = getenv("GOGP_TMA_MIN_K");
        tma_min_k = mk ? atoll(mk) : 512;  // below it the 2-CTA cp.async shape has the lower per-tile latency
        if (g_gemm_cfg < 0) {
            const char* e = getenv("GOGP_GEMM_CFG");
            g_gemm_cfg = e ? atoi(e) : 2;
        }
    });
    if (k <= 0) return;
    // Latency shapes: a launch of a few 128 x 128 tiles is bound by one SM's DMMA rate per tile, not by the GPU's;
    // quarter tiles with 32 x 32 warp tiles spread it over 4x the SMs (GOGP_SMALL_TILES: at most this many
    // 128-tiles, default 110: 4x the CTAs still fit in about three per SM; a sweep over 36 .. 296 at N = 4096 moved the
    // evaluation by 4 % only; 0 switches the shapes off).
    static int64_t small_tiles = 110;
    static std::once_flag small_knob;
    std::call_once(small_knob, [] {
        const char* e = getenv("GOGP_SMALL_TILES");
        if (e) small_tiles = atoll(e);
    });
    const int64_t tm128 = m / TILE, tn128 = n / TILE;
    const int64_t tiles128 = (mode & GEMM_LOWER) ? tm128 * (tm128 + 1) / 2 : tm128 * tn128;
    const bool small = tiles128 <= small_tiles && !(mode & GEMM_DIAG_OUT) && !(mask && mask->tb > 0) &&
                       (!(mode & GEMM_INPLACE) || n == TILE);
    if (!small && g_gemm_cfg == 2 && k >= tma_min_k &&
        launch_dgemm_tma(C, ldc, A, lda, B, ldb, m, n, k, alpha, beta, mode, cdiag, s, mask))
        return;
    GemmArgs g;
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
    if (mask && mask->tb > 0) {
        g.mtb = mask->tb;
        g.mr0 = mask->r0;
        g.mpr = mask->pr;
        g.mc0 = mask->c0;
        g.mpc = mask->pc;
    }
    if (small) {
        if (mode & GEMM_INPLACE)
            launch_cfg<1, 4, 3, 2, 4>(g, m, n, s);  // 32 x 128: the CTA owns entire rows of the 128-wide block
        else
            launch_cfg<2, 2, 3, 3, 4>(g, m, n, s);  // 64 x 64
        return;
    }
    // C aliasing A needs a CTA that owns entire rows of the (128-wide) block
    if (g_gemm_cfg == 0 || (mode & GEMM_INPLACE))  // (cfg 2 falls back to the 2-CTA shape below)
        launch_cfg<2, 4, 4, 1>(g, m, n, s);
    else
        launch_cfg<2, 2, 3, 2>(g, m, n, s);
}

// variant = nacc_index * 4 + ctas_index; nacc in {8, 16, 32}, CTAs per SM in {1, 2, 4, 8}
static const int kPeakAcc[3] = {8, 16, 32};
static const int kPeakCtas[4] = {1, 2, 4, 8};
int fp64_peak_variants() { return 12; }

void launch_fp64_peak(int which, int variant, int iters, double* sink, cudaStream_t s) {
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int ai = variant / 4, grid = nsm * kPeakCtas[variant % 4];
    if (which == 0) {
        if (ai == 0) dmma_peak_kernel<8><<<grid, 256, 0, s>>>(iters, sink);
        if (ai == 1) dmma_peak_kernel<16><<<grid, 256, 0, s>>>(iters, sink);
        if (ai == 2) dmma_peak_kernel<32><<<grid, 256, 0, s>>>(iters, sink);
    } else {
        if (ai == 0) dfma_peak_kernel<8><<<grid, 256, 0, s>>>(iters, sink);
        if (ai == 1) dfma_peak_kernel<16><<<grid, 256, 0, s>>>(iters, sink);
        if (ai == 2) dfma_peak_kernel<32><<<grid, 256, 0, s>>>(iters, sink);
    }
}

double fp64_peak_flops_per_launch(int which, int variant, int iters, int nsm) {
    const double warps = (double)nsm * kPeakCtas[variant % 4] * 8;
    const double nacc = kPeakAcc[variant / 4];
    if (which == 0) return warps * iters * nacc * (8 * 8 * 4 * 2);  // DMMA m8n8k4
    return warps * 32.0 * iters * nacc * 2;                        // DFMA
}

}  // namespace gogp
