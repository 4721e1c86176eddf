// The covariance-side kernels (build, fused gradient trace, input gradient) in a header of their own:
// cov.cu includes it inside namespace gogp and keeps the launchers; the CPU test tier compiles the SAME
// source for the host under the SIMT emulator of tests/simt/ (GOGP_SIMT_HOST), where the TMA staging of
// the coordinate tiles is replaced by a plain copy.  Under nvcc nothing changes (device SASS identical).
#pragma once

#if defined(GOGP_SIMT_HOST)
#define GOGP_DYN_SMEM_RAW(name) unsigned char* name = reinterpret_cast<unsigned char*>(simt::dyn_smem())
#else
#define GOGP_DYN_SMEM_RAW(name) extern __shared__ __align__(128) unsigned char name[]
#endif

constexpr int kFastMaxRest = 4;
constexpr int GT_E = 8;

__global__ void transpose_x_kernel(const double* __restrict__ X, double* __restrict__ Xt, int64_t N, int64_t Npad,
                                   int D) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= Npad) return;
    for (int d = 0; d < D; ++d) Xt[d * Npad + i] = i < N ? X[i * D + d] : 0.0;
}

// Stage the coordinates of one row tile and one column tile: [D][128] each.
__device__ __forceinline__ void stage_tiles(double* xr, double* xc, uint64_t* bar, const double* Rt, int64_t ldr,
                                            int64_t row0, const double* Ct, int64_t ldc, int64_t col0, int D) {
#if defined(GOGP_SIMT_HOST)
    (void)bar;
    if (threadIdx.x == 0)
        for (int d = 0; d < D; ++d)
            for (int t = 0; t < TILE; ++t) {
                xr[d * TILE + t] = Rt[d * ldr + row0 + t];
                xc[d * TILE + t] = Ct[d * ldc + col0 + t];
            }
    __syncthreads();
#else
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)(2 * D * TILE * sizeof(double)));
        for (int d = 0; d < D; ++d) {
            tma_load_1d(xr + d * TILE, Rt + d * ldr + row0, TILE * sizeof(double), bar);
            tma_load_1d(xc + d * TILE, Ct + d * ldc + col0, TILE * sizeof(double), bar);
        }
    }
    mbar_wait(bar, 0);
#endif
}

__device__ __forceinline__ double eval_program(const DevProgram& prog, const double* xc, int c, const double* xr,
                                               int r) {
    double k = 0.0;
    auto xa = [&](int d) { return xc[d * TILE + c]; };
    auto xb = [&](int d) { return xr[d * TILE + r]; };
    for (int t = 0; t < prog.nterms; ++t) k += term_value(prog, t, xa, xb);
    return k;
}

// out[row][col] = k(xa = C[col], xb = R[row]).  SYM: lower tiles of a square
// matrix, + noise on the diagonal, identity in the padding; otherwise a full
// rectangle with zero padding.
template <bool SYM>
__global__ void __launch_bounds__(256) cov_tile_kernel(const __grid_constant__ DevProgram prog,
                                                       const double* __restrict__ Rt, int64_t ldr, int64_t nrows,
                                                       const double* __restrict__ Ct, int64_t ldc, int64_t ncols,
                                                       int D, double noise, double* __restrict__ out, int64_t ld,
                                                       int tiles_n) {
    GOGP_DYN_SMEM_RAW(smem_raw);
    double* xr = reinterpret_cast<double*>(smem_raw);
    double* xc = xr + D * TILE;
    uint64_t* bar = reinterpret_cast<uint64_t*>(xc + D * TILE);

    int ti, tj;
    if (SYM) {
        lower_tile(blockIdx.x, ti, tj);
    } else {
        ti = blockIdx.x / tiles_n;
        tj = blockIdx.x % tiles_n;
    }
    const int64_t row0 = (int64_t)ti * TILE, col0 = (int64_t)tj * TILE;
    stage_tiles(xr, xc, bar, Rt, ldr, row0, Ct, ldc, col0, D);

    const int c0 = 2 * (threadIdx.x & 63);
    const int ir = threadIdx.x >> 6;
    for (int rr = 0; rr < TILE / 4; ++rr) {
        const int r = rr * 4 + ir;
        const int64_t gi = row0 + r, gj = col0 + c0;
        double2 v;
        v.x = eval_program(prog, xc, c0, xr, r);
        v.y = eval_program(prog, xc, c0 + 1, xr, r);
        if (SYM) {
            if (gi == gj) v.x += noise;
            if (gi == gj + 1) v.y += noise;
            if (gi >= nrows || gj >= ncols) v.x = (gi == gj) ? 1.0 : 0.0;
            if (gi >= nrows || gj + 1 >= ncols) v.y = (gi == gj + 1) ? 1.0 : 0.0;
        } else {
            if (gi >= nrows || gj >= ncols) v.x = 0.0;
            if (gi >= nrows || gj + 1 >= ncols) v.y = 0.0;
        }
        *reinterpret_cast<double2*>(out + gi * ld + gj) = v;
    }
}

// ---- specialised element loop ("fast shape") ----------------------------------------
// ncu showed the interpreted kernels above to be bound by instruction issue, not by HBM: ~185
// instructions per element for a 1-D RBF, ~40 of them the FP64 exp.  When the program is ONE
// product term whose leading factors are NN Normal leaves (the ARD kernels of every BASELINE
// config but hyperpriors), the Normal part is unrolled at compile time with the factor constants
// and the thread's own column coordinates hoisted into registers, and tiles that lie strictly
// below the diagonal and inside the data skip the per-element diagonal / padding masks.  The
// remaining factors (parameters, at most a few other leaves) keep the generic factor_value.
// Same operations in the same order as term_value: bit-identical to the interpreted kernel.
// Measured at N = 32768: 1-D RBF build 2.97 -> 1.71 ms (2.5 TB/s), C3 kernel 6.6 -> 4.7 ms.  The
// same treatment of the trace kernel (all-register, one element at a time) measured SLOWER than
// the staged interpreter (3.7 -> 4.3 ms RBF, 11.3 -> 12.1 ms C3: the 8-element batches of the
// interpreter overlap their exp latencies) and was dropped.
template <int NN, bool SYM>
__global__ void __launch_bounds__(256) cov_tile_fast_kernel(const __grid_constant__ DevProgram prog,
                                                            const double* __restrict__ Rt, int64_t ldr, int64_t nrows,
                                                            const double* __restrict__ Ct, int64_t ldc, int64_t ncols,
                                                            int D, double noise, double* __restrict__ out, int64_t ld,
                                                            int tiles_n) {
    GOGP_DYN_SMEM_RAW(smem_raw);
    double* xr = reinterpret_cast<double*>(smem_raw);
    double* xc = xr + D * TILE;
    uint64_t* bar = reinterpret_cast<uint64_t*>(xc + D * TILE);

    int ti, tj;
    if (SYM) {
        lower_tile(blockIdx.x, ti, tj);
    } else {
        ti = blockIdx.x / tiles_n;
        tj = blockIdx.x % tiles_n;
    }
    const int64_t row0 = (int64_t)ti * TILE, col0 = (int64_t)tj * TILE;
    stage_tiles(xr, xc, bar, Rt, ldr, row0, Ct, ldc, col0, D);

    const int c0 = 2 * (threadIdx.x & 63);
    const int ir = threadIdx.x >> 6;
    constexpr int NR = NN > 0 ? NN : 1;
    double inv[NR], ca0[NR], ca1[NR];
    int rofs[NR];
#pragma unroll
    for (int j = 0; j < NN; ++j) {
        const DevFactor& f = prog.f[j];
        inv[j] = f.i0;
        rofs[j] = f.dim * TILE;
        ca0[j] = xc[rofs[j] + c0];
        ca1[j] = xc[rofs[j] + c0 + 1];
    }
    const double coef = prog.coef[0];
    const int fe = prog.fbeg[1];
    const bool interior = (SYM ? ti > tj : true) && row0 + TILE <= nrows && col0 + TILE <= ncols;
    for (int rr = 0; rr < TILE / 4; ++rr) {
        const int r = rr * 4 + ir;
        double q0 = 0.0, q1 = 0.0;
#pragma unroll
        for (int j = 0; j < NN; ++j) {
            const double xb = xr[rofs[j] + r];
            const double d0 = (ca0[j] - xb) * inv[j], d1 = (ca1[j] - xb) * inv[j];
            q0 = fma(d0, d0, q0);
            q1 = fma(d1, d1, q1);
        }
        double2 v;
        v.x = coef;
        v.y = coef;
        if (NN > 0) {
            v.x *= exp(-q0 / 2);
            v.y *= exp(-q1 / 2);
        }
        for (int fi = NN; fi < fe; ++fi) {
            const DevFactor& f = prog.f[fi];
            const double xb = xr[f.dim * TILE + r], xa0 = xc[f.dim * TILE + c0], xa1 = xc[f.dim * TILE + c0 + 1];
            v.x *= (f.kind == F_EVENTS) ? events_value(prog, xa0, xb) : factor_value(f, xa0, xb);
            v.y *= (f.kind == F_EVENTS) ? events_value(prog, xa1, xb) : factor_value(f, xa1, xb);
        }
        const int64_t gi = row0 + r, gj = col0 + c0;
        if (!interior) {
            if (SYM) {
                if (gi == gj) v.x += noise;
                if (gi == gj + 1) v.y += noise;
                if (gi >= nrows || gj >= ncols) v.x = (gi == gj) ? 1.0 : 0.0;
                if (gi >= nrows || gj + 1 >= ncols) v.y = (gi == gj + 1) ? 1.0 : 0.0;
            } else {
                if (gi >= nrows || gj >= ncols) v.x = 0.0;
                if (gi >= nrows || gj + 1 >= ncols) v.y = 0.0;
            }
        }
        *reinterpret_cast<double2*>(out + gi * ld + gj) = v;
    }
}

__global__ void cov_self_kernel(const __grid_constant__ DevProgram prog, const double* __restrict__ Zt, int64_t M,
                                int64_t Mpad, double* __restrict__ kss) {
    int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    double k = 0.0;
    auto z = [&](int d) { return Zt[d * Mpad + m]; };
    for (int t = 0; t < prog.nterms; ++t) k += term_value(prog, t, z, z);
    kss[m] = k;
}

// ---- fused gradient trace ----------------------------------------------------------
// One CTA per lower tile; a thread owns a column pair and 32 rows, walked in chunks of
// 4 rows (E = 8 elements).  For each product term: phase 1 forms w_e * W_e * P_e (term
// value weighted by W = alpha alpha^T - K^-1 and by 1 below / 1/2 on the diagonal) per
// element; phase 2 runs over the term's factors and adds the P-weighted log-derivatives
// to per-thread accumulators, one shared-memory cell per (parameter slot, thread), so
// nothing is reduced across threads until the end of the tile.  dK is never
// materialised (the reference stores one dense N x N matrix per parameter,
// gp/gp.go:93-97,158-163).  ~70 KB of shared memory -> 3 CTAs per SM.
// Which tiles a trace launch walks and where they sit in the global matrix:
//   mode 0  the whole matrix: CTA b = lower tile (ti, tj), diagonal tiles in kdiag;
//   mode 1  a rows x cols block whose origin is element (grow0, gcol0): CTA b = tile (b / ctiles, b % ctiles);
//   mode 2  one rank's local matrix of a pr x pc block-cyclic distribution with tb x tb tiles per block
//           (grid.hpp): tile (ti, tj) belongs to global block (r0 + pr (ti / tb), c0 + pc (tj / tb)).
// kinv always points at the launch's own storage (tile (ti, tj) at kinv + ti 128 ld + tj 128); only elements with
// global row >= global column (< N) count.
struct TraceMap {
    int mode, ctiles;
    int64_t grow0, gcol0;
    int tb, r0, pr, c0, pc;
};
__device__ __forceinline__ void trace_tile(const TraceMap& m, int b, int& ti, int& tj, int64_t& row0, int64_t& col0) {
    if (m.mode == 0) {
        lower_tile(b, ti, tj);
        row0 = (int64_t)ti * TILE;
        col0 = (int64_t)tj * TILE;
    } else {
        ti = b / m.ctiles;
        tj = b % m.ctiles;
        if (m.mode == 1) {
            row0 = m.grow0 + (int64_t)ti * TILE;
            col0 = m.gcol0 + (int64_t)tj * TILE;
        } else {
            row0 = ((int64_t)(m.r0 + m.pr * (ti / m.tb)) * m.tb + ti % m.tb) * TILE;
            col0 = ((int64_t)(m.c0 + m.pc * (tj / m.tb)) * m.tb + tj % m.tb) * TILE;
        }
    }
}

__global__ void __launch_bounds__(256) grad_trace_kernel(const __grid_constant__ DevProgram prog,
                                                         const double* __restrict__ Xt, int64_t ldx,
                                                         const double* __restrict__ alpha,
                                                         const double* __restrict__ kinv, int64_t ld,
                                                         const double* __restrict__ kdiag, int64_t N, int D,
                                                         double* __restrict__ partial, const TraceMap map) {
    GOGP_DYN_SMEM_RAW(smem_raw);
    double* xr = reinterpret_cast<double*>(smem_raw);
    double* xc = xr + D * TILE;
    uint64_t* bar = reinterpret_cast<uint64_t*>(xc + D * TILE);
    double* pw = reinterpret_cast<double*>(bar + 2);  // [GT_E][256] weighted term values
    double* ww = pw + GT_E * 256;                     // [GT_E][256] weights w_e * W_e
    double* acc = ww + GT_E * 256;                    // [ntheta + 1][256] per-thread accumulators

    int ti, tj;
    int64_t row0, col0;
    trace_tile(map, blockIdx.x, ti, tj, row0, col0);
    const int64_t lrow0 = (int64_t)ti * TILE, lcol0 = (int64_t)tj * TILE;
    const int tid = threadIdx.x;
    const int nslot = prog.ntheta;  // slot ntheta = trace of W
    if (col0 > row0 + TILE - 1) {   // a tile above the diagonal (upper part of a diagonal block): nothing counts
        if (tid <= nslot) partial[(int64_t)blockIdx.x * (nslot + 1) + tid] = 0.0;
        return;
    }
    for (int q = 0; q <= nslot; ++q) acc[q * 256 + tid] = 0.0;
    stage_tiles(xr, xc, bar, Xt, ldx, row0, Xt, ldx, col0, D);

    const bool side_diag = kdiag != nullptr && ti == tj;
    const double* ktile = side_diag ? kdiag + (int64_t)ti * TILE * TILE : kinv + lrow0 * ld + lcol0;
    const int64_t kld = side_diag ? TILE : ld;
    const int c0 = 2 * (tid & 63);
    const int ir = tid >> 6;
    const int64_t gj = col0 + c0;
    const double aj0 = alpha[gj], aj1 = alpha[gj + 1];
    double trw = 0.0;

    for (int chunk = 0; chunk < 2 * TILE / (4 * GT_E); ++chunk) {
        // weights w_e * W_e of this chunk (1 below the diagonal, 1/2 on it, 0 elsewhere)
#pragma unroll
        for (int e = 0; e < GT_E; e += 2) {
            const int r = (chunk * (GT_E / 2) + (e >> 1)) * 4 + ir;
            const int64_t gi = row0 + r;
            const double2 kv = *reinterpret_cast<const double2*>(ktile + (int64_t)r * kld + c0);
            const double ai = alpha[gi];
            double w0 = 0.0, w1 = 0.0;
            if (gi < N && gj < N && gi >= gj) {
                const double W = ai * aj0 - kv.x;
                if (gi == gj) {
                    trw += W;
                    w0 = 0.5 * W;
                } else {
                    w0 = W;
                }
            }
            if (gi < N && gj + 1 < N && gi >= gj + 1) {
                const double W = ai * aj1 - kv.y;
                if (gi == gj + 1) {
                    trw += W;
                    w1 = 0.5 * W;
                } else {
                    w1 = W;
                }
            }
            ww[e * 256 + tid] = w0;
            ww[(e + 1) * 256 + tid] = w1;
        }
        for (int t = 0; t < prog.nterms; ++t) {
            const int fb = prog.fbeg[t], fe = prog.fbeg[t + 1];
            // phase 1: weighted term values
            for (int e = 0; e < GT_E; ++e) {
                const int r = (chunk * (GT_E / 2) + (e >> 1)) * 4 + ir;
                const int c = c0 + (e & 1);
                auto xa = [&](int d) { return xc[d * TILE + c]; };
                auto xb = [&](int d) { return xr[d * TILE + r]; };
                const double w = ww[e * 256 + tid];
                pw[e * 256 + tid] = (w != 0.0) ? w * term_value(prog, t, xa, xb) : 0.0;
            }
            // phase 2: per-factor log-derivatives
            for (int fi = fb; fi < fe; ++fi) {
                const DevFactor& f = prog.f[fi];
                if (f.p0 < 0) continue;  // parameter-free factor (events)
                double s0 = 0.0, s1 = 0.0;
                for (int e = 0; e < GT_E; ++e) {
                    const int r = (chunk * (GT_E / 2) + (e >> 1)) * 4 + ir;
                    const int c = c0 + (e & 1);
                    double g0, g1;
                    factor_dlog_theta(f, xc[f.dim * TILE + c], xr[f.dim * TILE + r], g0, g1);
                    const double p = pw[e * 256 + tid];
                    s0 = fma(p, g0, s0);
                    s1 = fma(p, g1, s1);
                }
                acc[f.p0 * 256 + tid] += s0;
                if (f.p1 >= 0) acc[f.p1 * 256 + tid] += s1;
            }
        }
    }
    acc[nslot * 256 + tid] = trw;
    __syncthreads();
    // fixed-order tree over the 256 threads of each slot (bit-repeatable)
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o)
            for (int q = 0; q <= nslot; ++q) acc[q * 256 + tid] += acc[q * 256 + tid + o];
        __syncthreads();
    }
    if (tid <= nslot) partial[(int64_t)blockIdx.x * (nslot + 1) + tid] = acc[tid * 256];
}

// ---- specialised fused trace ---------------------------------------------------------------------
// ncu showed the interpreted trace above bound by instruction issue (~600 thread instructions per element for the
// C3 kernel, FP64 pipe 41 %): the descriptor walk, the term value computed in one phase and every transcendental
// computed AGAIN for the log-derivatives in the next, shared-memory staging of both.  Here, for a product term
// whose leading NN factors are Normal leaves:
//   - the Normal factors are unrolled at compile time; d_j^2 is both the exponent's summand and the factor's
//     log-derivative, so the gradient of a Normal factor costs one FMA on top of the value;
//   - a bare parameter multiplies the term's coefficient once per CTA and its gradient slot is the plain sum of
//     the weighted term values;
//   - the (at most kTraceMaxRest) remaining leaves are evaluated once, value and log-derivatives together
//     (factor_value_dlog: one sincos + one exp per Periodic, one exp + one reciprocal per Matern);
//   - everything stays in registers: a thread walks its column pair down the tile one row at a time (2 elements,
//     independent dependency chains) and flushes its accumulators to its shared-memory cells once per term;
//   - tiles strictly below the diagonal and inside the data skip the diagonal / padding masks, and the next
//     row of K^-1 is requested before the current one is used.
// NN = 0 handles any number of terms (e.g. hyperpriors' Matern52 + Periodic sum).
// MAXR = how many such leaves the instantiation carries registers for (0, 1 or kTraceMaxRest): the occupancy of this
// latency-bound kernel is set by its register count (NN = 8: 232 registers and 8 warps per SM measured 11.0 ms at
// N = 32768 for the C3 kernel, the same code squeezed into 128 registers and 16 warps 8.0 ms).
constexpr int kTraceMaxRest = 4;

template <int NN, int MAXR, int MINB>
__global__ void __launch_bounds__(256, MINB) grad_trace_fast_kernel(const __grid_constant__ DevProgram prog,
                                                                    const double* __restrict__ Xt, int64_t ldx,
                                                                    const double* __restrict__ alpha,
                                                                    const double* __restrict__ kinv, int64_t ld,
                                                                    const double* __restrict__ kdiag, int64_t N, int D,
                                                                    double* __restrict__ partial, const TraceMap map) {
    GOGP_DYN_SMEM_RAW(smem_raw);
    double* xr = reinterpret_cast<double*>(smem_raw);
    double* xc = xr + D * TILE;
    uint64_t* bar = reinterpret_cast<uint64_t*>(xc + D * TILE);
    double* acc = reinterpret_cast<double*>(bar + 2);  // [ntheta + 1][256] per-thread cells

    int ti, tj;
    int64_t row0, col0;
    trace_tile(map, blockIdx.x, ti, tj, row0, col0);
    const int tid = threadIdx.x;
    const int nslot = prog.ntheta;
    if (col0 > row0 + TILE - 1) {
        if (tid <= nslot) partial[(int64_t)blockIdx.x * (nslot + 1) + tid] = 0.0;
        return;
    }
    for (int q = 0; q <= nslot; ++q) acc[q * 256 + tid] = 0.0;
    stage_tiles(xr, xc, bar, Xt, ldx, row0, Xt, ldx, col0, D);

    const bool side_diag = map.mode == 0 && kdiag != nullptr && ti == tj;
    const double* ktile = side_diag ? kdiag + (int64_t)ti * TILE * TILE : kinv + (int64_t)ti * TILE * ld + (int64_t)tj * TILE;
    const int64_t kld = side_diag ? TILE : ld;
    const int c0 = 2 * (tid & 63);
    const int ir = tid >> 6;
    const int64_t gj = col0 + c0;
    const double aj0 = alpha[gj], aj1 = alpha[gj + 1];
    const bool interior = row0 >= col0 + TILE && row0 + TILE <= N;
    double trw = 0.0;

    for (int t = 0; t < prog.nterms; ++t) {
        const int fb = prog.fbeg[t], fe = prog.fbeg[t + 1];
        // bare parameters fold into the coefficient; ridx[k] = the k-th other factor behind the Normal ones
        double cterm = prog.coef[t];
        for (int fi = fb + NN; fi < fe; ++fi)
            if (prog.f[fi].kind == F_PARAM) cterm *= prog.f[fi].a0;
        constexpr int MR = MAXR > 0 ? MAXR : 1;
        int ridx[MR];
        {
            int fi = fb + NN;
#pragma unroll
            for (int k = 0; k < MAXR; ++k) {
                while (fi < fe && prog.f[fi].kind == F_PARAM) ++fi;
                ridx[k] = fi < fe ? fi++ : -1;
            }
        }
        constexpr int NR = NN > 0 ? NN : 1;
        int rofs[NR];
        double accN[NR];
#pragma unroll
        for (int j = 0; j < NN; ++j) {
            rofs[j] = prog.f[fb + j].dim * TILE;
            accN[j] = 0.0;
        }
        double accR0[MR], accR1[MR];
#pragma unroll
        for (int k = 0; k < MAXR; ++k) accR0[k] = accR1[k] = 0.0;
        double sP = 0.0;

        double2 kv = *reinterpret_cast<const double2*>(ktile + (int64_t)ir * kld + c0);
        for (int rr = 0; rr < TILE / 4; ++rr) {
            const int r = rr * 4 + ir;
            const int64_t gi = row0 + r;
            const double2 kc = kv;
            if (rr + 1 < TILE / 4) kv = *reinterpret_cast<const double2*>(ktile + (int64_t)(r + 4) * kld + c0);
            const double ai = alpha[gi];
            double w0 = ai * aj0 - kc.x, w1 = ai * aj1 - kc.y;
            if (!interior) {
                const bool in0 = gi < N && gj < N && gi >= gj, in1 = gi < N && gj + 1 < N && gi >= gj + 1;
                if (!in0) w0 = 0.0;
                if (!in1) w1 = 0.0;
                if (gi == gj) {
                    if (t == 0) trw += w0;
                    w0 *= 0.5;
                }
                if (gi == gj + 1) {
                    if (t == 0) trw += w1;
                    w1 *= 0.5;
                }
            }
            double p0 = cterm, p1 = cterm;
            double e0[NR], e1[NR];
            if (NN > 0) {
                double q0 = 0.0, q1 = 0.0;
#pragma unroll
                for (int j = 0; j < NN; ++j) {
                    const double xb = xr[rofs[j] + r], inv = prog.f[fb + j].i0;
                    const double d0 = (xc[rofs[j] + c0] - xb) * inv, d1 = (xc[rofs[j] + c0 + 1] - xb) * inv;
                    e0[j] = d0 * d0;
                    e1[j] = d1 * d1;
                    q0 += e0[j];
                    q1 += e1[j];
                }
                p0 *= exp(-q0 / 2);
                p1 *= exp(-q1 / 2);
            }
            double g00[MR], g01[MR], g10[MR], g11[MR];
#pragma unroll
            for (int k = 0; k < MAXR; ++k) {
                g00[k] = g01[k] = g10[k] = g11[k] = 0.0;
                if (ridx[k] >= 0) {
                    const DevFactor& f = prog.f[ridx[k]];
                    const double xb = xr[f.dim * TILE + r];
                    double v0, v1;
                    factor_value_dlog(prog, f, xc[f.dim * TILE + c0], xb, v0, g00[k], g01[k]);
                    factor_value_dlog(prog, f, xc[f.dim * TILE + c0 + 1], xb, v1, g10[k], g11[k]);
                    p0 *= v0;
                    p1 *= v1;
                }
            }
            const double wp0 = w0 * p0, wp1 = w1 * p1;
            sP += wp0 + wp1;
#pragma unroll
            for (int j = 0; j < NN; ++j) accN[j] = fma(wp0, e0[j], fma(wp1, e1[j], accN[j]));
#pragma unroll
            for (int k = 0; k < MAXR; ++k) {
                accR0[k] = fma(wp0, g00[k], fma(wp1, g10[k], accR0[k]));
                accR1[k] = fma(wp0, g01[k], fma(wp1, g11[k], accR1[k]));
            }
        }
        // flush this term's sums into the thread's own cells
#pragma unroll
        for (int j = 0; j < NN; ++j) acc[prog.f[fb + j].p0 * 256 + tid] += accN[j];
#pragma unroll
        for (int k = 0; k < MAXR; ++k)
            if (ridx[k] >= 0) {
                const DevFactor& f = prog.f[ridx[k]];
                if (f.p0 >= 0) acc[f.p0 * 256 + tid] += accR0[k];
                if (f.p1 >= 0) acc[f.p1 * 256 + tid] += accR1[k];
            }
        for (int fi = fb + NN; fi < fe; ++fi)
            if (prog.f[fi].kind == F_PARAM) acc[prog.f[fi].p0 * 256 + tid] += sP;
    }
    acc[nslot * 256 + tid] = trw;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o)
            for (int q = 0; q <= nslot; ++q) acc[q * 256 + tid] += acc[q * 256 + tid + o];
        __syncthreads();
    }
    if (tid <= nslot) partial[(int64_t)blockIdx.x * (nslot + 1) + tid] = acc[tid * 256];
}

// Deterministic second stage: one CTA per slot, fixed summation order.
__global__ void __launch_bounds__(256) grad_reduce_kernel(const double* __restrict__ partial, int nblocks, int nslots,
                                                          double* __restrict__ out, int accumulate) {
    __shared__ double sh[256];
    const int q = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) s += partial[(int64_t)b * nslots + q];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[q] = accumulate ? out[q] + sh[0] : sh[0];
}

// ---- input gradient (with_obs) ---------------------------------------------------
// One CTA per observation i; for each coordinate d a block reduction over j != i of
// W_ij * d k(x_i, x_j) / d x_{i,d}.  O(N^2 D F); the reference needs N*D dense
// N x N matrices and 4N^3 flops each for the same numbers (gp/gp.go:93-94,118-129).
__global__ void __launch_bounds__(256) grad_inputs_kernel(const __grid_constant__ DevProgram prog,
                                                          const double* __restrict__ Xt, int64_t ldx,
                                                          const double* __restrict__ alpha,
                                                          const double* __restrict__ kinv, int64_t ld,
                                                          const double* __restrict__ kdiag, int64_t N, int D,
                                                          double* __restrict__ gx) {
    __shared__ double sh[256];
    const int64_t i = blockIdx.x;
    const double ai = alpha[i];
    for (int d = 0; d < D; ++d) {
        double s = 0.0;
        for (int64_t j = threadIdx.x; j < N; j += 256) {
            if (j == i) continue;
            const int64_t hi = i > j ? i : j, lo = i > j ? j : i;
            const int64_t th = hi / TILE, tl = lo / TILE;
            double kv = (th == tl) ? kdiag[th * TILE * TILE + (hi % TILE) * TILE + (lo % TILE)] : kinv[hi * ld + lo];
            const double W = ai * alpha[j] - kv;
            double dk = 0.0;
            auto xa = [&](int dd) { return Xt[dd * ldx + i]; };
            auto xb = [&](int dd) { return Xt[dd * ldx + j]; };
            for (int t = 0; t < prog.nterms; ++t) {
                double g = 0.0;
                for (int fi = prog.fbeg[t]; fi < prog.fbeg[t + 1]; ++fi) {
                    const DevFactor& f = prog.f[fi];
                    if (f.dim == d && f.kind != F_PARAM) g += factor_dlog_xa(f, xa(d), xb(d));
                }
                if (g != 0.0) dk += term_value(prog, t, xa, xb) * g;
            }
            s += W * dk;
        }
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) gx[i * D + d] = sh[0];
        __syncthreads();
    }
}
