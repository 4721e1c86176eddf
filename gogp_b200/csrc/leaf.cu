// Tile-level (128x128) and vector kernels around the DMMA GEMM: diagonal-block
// Cholesky + inverse, blocked triangular solves for alpha = K^-1 y
// (gp/gp.go:232-236), log-determinant and y.alpha (gp/gp.go:250-251), and small
// reductions used by Produce (gp/gp.go:335-357).
#include <cstdio>
#include <cstdlib>

#include "kernels.h"

namespace gogp {

namespace {

constexpr int LP = 132;  // smem pitch (doubles), = 4 mod 16: a half-warp = 4 rows x 4 k-phases hits 16 distinct 8-byte banks

// First version of the tile kernel (kept selectable, set_leaf_variant(1) / GOGP_LEAF=1, as the
// timing baseline): 150 us per tile, bound by one CTA-wide barrier per column.
// Cholesky (Crout, column by column) of one 128x128 tile held in shared memory,
// fused with the inverse W = L^-1 built row by row one step behind: at step j
// the threads of rows r >= j form column j of L (dot products of rows r and j
// over k < j), while the threads of rows r < j -- idle in a plain Crout sweep --
// form row j-1 of W,  W[j-1][r] = -(sum_{k=r}^{j-2} L[j-1][k] W[k][r]) / L[j-1][j-1].
// Four threads share a row and split k by k mod 4; one barrier per column.
// W[i][c] (c <= i) lives at S[c][i+1], the unused upper triangle of the array.
__global__ void __launch_bounds__(512, 1) potrf_leaf_crout_kernel(double* __restrict__ A, int64_t ld,
                                                                  double* __restrict__ winv, int* __restrict__ info,
                                                                  int base) {
    extern __shared__ double S[];  // [128][LP]
    __shared__ double adiag[TILE];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < TILE * LP; idx += 512) {
        const int r = idx / LP, c = idx - r * LP;
        S[idx] = (c <= r) ? A[(int64_t)r * ld + c] : 0.0;
    }
    if (tid < TILE) adiag[tid] = A[(int64_t)tid * ld + tid];
    __syncthreads();

    const int r = tid >> 2, h = tid & 3;
    const double* Sr = S + r * LP;
    // adiag[i] is kept right-looking: after column j, adiag[i] = A[i][i] - sum_{k<=j} L[i][k]^2, so
    // the pivot of column j is read, not summed (each row's h == 0 thread maintains its own entry)
    for (int j = 0; j <= TILE; ++j) {
        double s0 = 0.0, s1 = 0.0;  // two independent accumulator chains
        if (r >= j) {
            if (r > j && j < TILE) {
                const double* Sj = S + j * LP;
                int k = h;
#pragma unroll 2
                for (; k + 4 < j; k += 8) {
                    s0 = fma(Sr[k], Sj[k], s0);
                    s1 = fma(Sr[k + 4], Sj[k + 4], s1);
                }
                if (k < j) s0 = fma(Sr[k], Sj[k], s0);
            }
        } else if (r < j - 1) {
            const double* Sj = S + (j - 1) * LP;
            int k = r + ((h - r) & 3);  // smallest k >= r with k mod 4 == h
#pragma unroll 2
            for (; k + 4 < j - 1; k += 8) {
                s0 = fma(Sj[k], Sr[k + 1], s0);
                s1 = fma(Sj[k + 4], Sr[k + 5], s1);
            }
            if (k < j - 1) s0 = fma(Sj[k], Sr[k + 1], s0);
        }
        double s = s0 + s1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (h == 0) {
            if (r >= j) {
                if (j < TILE) {
                    double ajj = adiag[j];
                    if (!(ajj > 0.0)) {  // not positive definite (or NaN): flag, keep going
                        if (r == j) atomicCAS(info, 0, base + j + 1);
                        ajj = 1.0;
                    }
                    if (r == j) {
                        S[j * LP + j] = sqrt(ajj);
                    } else {
                        const double l = (Sr[j] - s) * rsqrt(ajj);  // one reciprocal square root, no divide
                        S[r * LP + j] = l;
                        adiag[r] = fma(-l, l, adiag[r]);
                    }
                }
            } else {
                const double dj = 1.0 / S[(j - 1) * LP + (j - 1)];
                S[r * LP + j] = (r == j - 1) ? dj : -s * dj;
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < TILE * TILE; idx += 512) {
        const int rr = idx >> 7, c = idx & 127;
        A[(int64_t)rr * ld + c] = (c <= rr) ? S[rr * LP + c] : 0.0;
        winv[idx] = (c <= rr) ? S[c * LP + rr + 1] : 0.0;
    }
}

__device__ __forceinline__ void dmma_leaf(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int MP = 100;  // pitch of the 32 x 96 side buffer: rows 32 B apart mod 128 B, like LP
constexpr int WS = 1;    // W[i][c] (c <= i) lives at S[c][i + WS], the unused upper triangle

constexpr int LEAF_THREADS = 384;  // 12 warps: 170 registers each, no spills in the register-resident phases
// Blocked tile kernel: Cholesky of one 128x128 tile and its inverse W = L^-1, by 32-column
// blocks so that the serial part never meets a CTA-wide barrier.  Per block b (o = 32 b):
//   A  warp 0, registers: lane i holds row i of the 32x32 diagonal block; right-looking,
//      column by column, pivot and column broadcast by warp shuffles (no barrier);
//   B  warp 0: Wd = D^-1, lane c solves D w = e_c forward (rows of D broadcast from smem);
//   C  panel below:  X = P Wd^T                                   (DMMA, 16-row warp tiles)
//   E1 M = -Wd L[b, 0:o]  into the side buffer                     (DMMA)
//   D  trailing update:  A22 -= X X^T, lower 16x16 warp tiles       (DMMA)
//   E2 block row b of the inverse:  W[b, j] = sum_{k=j}^{b-1} M[:, k] W[k, j]   (DMMA)
// C/E1 and D/E2 are independent pairs and share a phase; three CTA barriers per block.
// Fragment convention as in dgemm.cu: a = A[row fr + 8i][k fk], b = B[col fr + 8j][k fk],
// acc = C[row fr + 8i][col 2 fk + 8j + {0,1}], fr = lane >> 2, fk = lane & 3.
__global__ void __launch_bounds__(LEAF_THREADS, 1) potrf_leaf_kernel(double* __restrict__ A, int64_t ld,
                                                            double* __restrict__ winv, int* __restrict__ info,
                                                            int base) {
    extern __shared__ __align__(16) double S[];  // [128][LP], then the side buffer [32][MP]
    double* Mb = S + TILE * LP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    constexpr unsigned FULL = 0xffffffffu;

    for (int idx = tid; idx < TILE * (LP / 2); idx += LEAF_THREADS) {
        const int r = idx / (LP / 2), c2 = (idx - r * (LP / 2)) * 2;
        double2 v = make_double2(0.0, 0.0);
        if (c2 <= r) {
            v = *reinterpret_cast<const double2*>(A + (int64_t)r * ld + c2);
            if (c2 + 1 > r) v.y = 0.0;
        }
        *reinterpret_cast<double2*>(S + r * LP + c2) = v;
    }
    __syncthreads();

    for (int b = 0; b < 4; ++b) {
        const int o = 32 * b;
        if (warp == 0) {
            // ---- A: diagonal block in registers ----
            double a[32];
            {
                const double* row = S + (o + lane) * LP + o;
#pragma unroll
                for (int k = 0; k < 32; k += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(row + k);
                    a[k] = v.x;
                    a[k + 1] = v.y;
                }
            }
            double rinv = 0.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                double d = __shfl_sync(FULL, a[j], j);
                if (!(d > 0.0)) {  // not positive definite (or NaN): flag, keep going
                    if (lane == j) atomicCAS(info, 0, base + o + j + 1);
                    d = 1.0;
                }
                const double rs = rsqrt(d);
                const double l = (lane == j ? d : a[j]) * rs;  // one reciprocal square root, no divide
                a[j] = l;
                if (lane == j) rinv = rs;
#pragma unroll
                for (int k = j + 1; k < 32; ++k) {
                    const double lk = __shfl_sync(FULL, l, k);
                    a[k] = fma(-l, lk, a[k]);
                }
            }
            {
                double* row = S + (o + lane) * LP + o;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (k <= lane) row[k] = a[k];
            }
            __syncwarp();
            // ---- B: inverse of the diagonal block, one column per lane ----
            double t[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const double rk = __shfl_sync(FULL, rinv, k);
                t[k] *= rk;
#pragma unroll
                for (int i = k + 1; i < 32; ++i) t[i] = fma(-S[(o + i) * LP + o + k], t[k], t[i]);
            }
            {
                double* row = S + (o + lane) * LP + o + WS;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i >= lane) row[i] = t[i];
            }
        }
        __syncthreads();

        // ---- phase 2: C (panel solve) and E1 (scaled block row) ----
        const int nC = (96 - o) / 16, nE1 = 2 * b;
        for (int task = warp; task < nC + nE1; task += LEAF_THREADS / 32) {
            if (task < nC) {
                const int r0 = o + 32 + 16 * task;
                double acc[2][4][2];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const int k = 4 * ks + fk;
                    double av[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) av[i] = S[(r0 + fr + 8 * i) * LP + o + k];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (ks > 2 * j + 1) continue;  // Wd[c][k] = 0 for k > c
                        const int c = 8 * j + fr;
                        double bv = S[(o + k) * LP + o + c + WS];
                        if (ks >= 2 * j && k > c) bv = 0.0;
#pragma unroll
                        for (int i = 0; i < 2; ++i) dmma_leaf(acc[i][j][0], acc[i][j][1], av[i], bv);
                    }
                }
                __syncwarp();  // every lane has read its rows of P before X replaces them
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<double2*>(S + (r0 + fr + 8 * i) * LP + o + 8 * j + 2 * fk) =
                            make_double2(acc[i][j][0], acc[i][j][1]);
            } else {
                const int te = task - nC;
                const int i0 = 16 * (te & 1), q0 = 32 * (te >> 1);
                double acc[2][4][2];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
                for (int ks = 0; 4 * ks <= i0 + 15; ++ks) {  // Wd[i][m] = 0 for m > i
                    const int m = 4 * ks + fk;
                    double av[2], bv[4];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int ii = i0 + fr + 8 * i;
                        av[i] = (m <= ii) ? S[(o + m) * LP + o + ii + WS] : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) bv[j] = S[(o + m) * LP + q0 + fr + 8 * j];
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dmma_leaf(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<double2*>(Mb + (i0 + fr + 8 * i) * MP + q0 + 8 * j + 2 * fk) =
                            make_double2(-acc[i][j][0], -acc[i][j][1]);
            }
        }
        __syncthreads();

        // ---- phase 3: D (trailing update) and E2 (block row b of the inverse) ----
        const int nt = (96 - o) / 16, nD = nt * (nt + 1) / 2, nE2 = 4 * b;
        for (int task = warp; task < nD + nE2; task += LEAF_THREADS / 32) {
            double acc[2][2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            if (task < nD) {
                int ti = 0, p = task;
                while (p > ti) {
                    p -= ti + 1;
                    ++ti;
                }
                const int r0 = o + 32 + 16 * ti, c0 = o + 32 + 16 * p;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const int k = o + 4 * ks + fk;
                    double av[2], bv[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) av[i] = S[(r0 + fr + 8 * i) * LP + k];
#pragma unroll
                    for (int j = 0; j < 2; ++j) bv[j] = S[(c0 + fr + 8 * j) * LP + k];
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) dmma_leaf(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int row = r0 + fr + 8 * i, col = c0 + 8 * j + 2 * fk + e;
                            if (col <= row) S[row * LP + col] -= acc[i][j][e];  // above the diagonal lives W
                        }
            } else {
                const int te = task - nD;
                const int jb = te >> 2, i0 = 16 * ((te >> 1) & 1), c0 = 32 * jb + 16 * (te & 1);
                for (int q0 = c0; q0 < o; q0 += 4) {  // W[q][cc] = 0 for q < cc
                    const int q = q0 + fk;
                    double av[2], bv[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) av[i] = Mb[(i0 + fr + 8 * i) * MP + q];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int cc = c0 + fr + 8 * j;
                        bv[j] = (q >= cc) ? S[cc * LP + q + WS] : 0.0;
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) dmma_leaf(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e)
                            S[(c0 + 8 * j + 2 * fk + e) * LP + o + i0 + fr + 8 * i + WS] = acc[i][j][e];
            }
        }
        __syncthreads();
    }

    for (int idx = tid; idx < TILE * TILE; idx += LEAF_THREADS) {
        const int rr = idx >> 7, c = idx & 127;
        A[(int64_t)rr * ld + c] = (c <= rr) ? S[rr * LP + c] : 0.0;
        winv[idx] = (c <= rr) ? S[c * LP + rr + WS] : 0.0;
    }
}

__global__ void __launch_bounds__(256) trtri_leaf_kernel(const double* __restrict__ winv, double* __restrict__ dst,
                                                         int64_t ld) {
    __shared__ double T[32][33];
    // 4 x 4 grid of 32 x 32 sub-tiles, transposed through smem
    const int bi = blockIdx.x >> 2, bj = blockIdx.x & 3;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) T[r][tx] = winv[(bj * 32 + r) * TILE + bi * 32 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int gr = bi * 32 + r, gc = bj * 32 + tx;
        dst[(int64_t)gr * ld + gc] = (gc >= gr) ? T[tx][r] : 0.0;
    }
}

// 128-row block times a 128-vector held in registers (4 contiguous columns per
// lane): every lane issues all of its loads before the first reduction.
template <int ROWS>
__device__ __forceinline__ void block_matvec(const double* __restrict__ M, int64_t ldm, const double* vs, int lane,
                                             double (&out)[ROWS]) {
    const double v0 = vs[4 * lane], v1 = vs[4 * lane + 1], v2 = vs[4 * lane + 2], v3 = vs[4 * lane + 3];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
        const double2* p = reinterpret_cast<const double2*>(M + (int64_t)rr * ldm + 4 * lane);
        const double2 a = p[0], b = p[1];
        out[rr] = a.x * v0 + a.y * v1 + b.x * v2 + b.y * v3;
    }
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) out[rr] += __shfl_xor_sync(0xffffffffu, out[rr], o);
}

// out[lane] without dynamic register indexing
template <int ROWS>
__device__ __forceinline__ double pick(const double (&out)[ROWS], int lane) {
    double v = 0.0;
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr)
        if (lane == rr) v = out[rr];
    return v;
}

// z_0 = Winv_0 w_0 (forward) -- the first block of the pipelined solves below.
__global__ void __launch_bounds__(256) trsv_first_kernel(const double* __restrict__ winv,
                                                         const double* __restrict__ w, double* __restrict__ z,
                                                         int I, int transposed) {
    __shared__ double ws[TILE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < TILE) ws[tid] = w[(int64_t)I * TILE + tid];
    __syncthreads();
    const double* Wi = winv + (int64_t)I * TILE * TILE;
    if (!transposed) {
        double out[16];
        block_matvec<16>(Wi + (int64_t)warp * 16 * TILE, TILE, ws, lane, out);
        if (lane < 16) z[(int64_t)I * TILE + warp * 16 + lane] = pick(out, lane);  // every lane holds all 16 sums
    } else if (tid < TILE) {
        double s = 0.0;
        for (int r = tid; r < TILE; ++r) s += Wi[r * TILE + tid] * ws[r];
        z[(int64_t)I * TILE + tid] = s;
    }
}

// One step of the blocked forward solve  L z = y  (w is the running right-hand
// side, z_I is already known): CTA b updates w_{I+1+b} -= L[I+1+b, I] z_I; CTA 0,
// whose block is then final, also produces z_{I+1} = Winv_{I+1} w_{I+1}, so a
// step is one launch and no CTA recomputes anything.
__global__ void __launch_bounds__(256) trsv_fwd_step_kernel(const double* __restrict__ L, int64_t ld,
                                                            const double* __restrict__ winv, double* __restrict__ w,
                                                            double* __restrict__ z, int I) {
    __shared__ double zs[TILE], ws[TILE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < TILE) zs[tid] = z[(int64_t)I * TILE + tid];
    __syncthreads();
    const int64_t rb = ((int64_t)I + 1 + blockIdx.x) * TILE;
    double out[16];
    block_matvec<16>(L + (rb + warp * 16) * ld + (int64_t)I * TILE, ld, zs, lane, out);
    if (lane < 16) {
        const double v = w[rb + warp * 16 + lane] - pick(out, lane);
        w[rb + warp * 16 + lane] = v;
        ws[warp * 16 + lane] = v;
    }
    if (blockIdx.x != 0) return;
    __syncthreads();
    const double* Wn = winv + ((int64_t)I + 1) * TILE * TILE;
    block_matvec<16>(Wn + (int64_t)warp * 16 * TILE, TILE, ws, lane, out);
    if (lane < 16) z[rb + warp * 16 + lane] = pick(out, lane);
}

// One step of the blocked backward solve  L^T x = z  (x_I already known): CTA b
// updates w_J -= L[I, J]^T x_I for J = I-1-b; CTA 0 (J = I-1, final after this
// update) also produces x_{I-1} = Winv_{I-1}^T w_{I-1}.
__global__ void __launch_bounds__(256) trsv_bwd_step_kernel(const double* __restrict__ L, int64_t ld,
                                                            const double* __restrict__ winv, double* __restrict__ w,
                                                            double* __restrict__ x, int I) {
    __shared__ double xs[TILE], ws[TILE], part[TILE];
    const int tid = threadIdx.x;
    if (tid < TILE) xs[tid] = x[(int64_t)I * TILE + tid];
    __syncthreads();
    const int64_t J = (int64_t)I - 1 - blockIdx.x;
    const double* Lb = L + (int64_t)I * TILE * ld + J * TILE;
    const int c = tid & 127, half = tid >> 7;
    double s = 0.0;
#pragma unroll 16
    for (int r = half * 64; r < half * 64 + 64; ++r) s += Lb[(int64_t)r * ld + c] * xs[r];
    if (half) part[c] = s;
    __syncthreads();
    if (!half) {
        const double v = w[J * TILE + c] - (s + part[c]);
        w[J * TILE + c] = v;
        ws[c] = v;
    }
    if (blockIdx.x != 0) return;
    __syncthreads();
    const double* Wn = winv + J * TILE * TILE;
    s = 0.0;
#pragma unroll 8
    for (int r = c + half; r < TILE; r += 2) s += Wn[r * TILE + c] * ws[r];
    if (half) part[c] = s;
    __syncthreads();
    if (!half) x[J * TILE + c] = s + part[c];
}

// ---- single-launch triangular solves (default) --------------------------------------------
// The step kernels above cost one launch per 128-row block (2 x 255 dependent launches at
// N = 32768, ~21 us each).  Here ONE launch per direction: a CTA per 128-row block takes a
// ticket (so blocks start in dependency order whatever the hardware's dispatch order), streams
// its row (forward) / column (backward) of L block by block, and for each block waits on a
// ready flag that the producing CTA releases after storing its part of the solution
// (st.release.gpu after __threadfence / ld.acquire.gpu).  A CTA only ever waits for smaller
// tickets, which are running or finished, so there is no deadlock however many CTAs are
// resident.  The block of L is prefetched into registers BEFORE the wait and the 128x128
// inverse of the CTA's own diagonal tile sits in shared memory, so the dependent chain per
// block is: flag -> 1 KB of the solution from L2 -> FMAs -> one reduction -> tile matvec -> release.
// sync[0] is the ticket counter, sync[1 + b] the flag of block b; zeroed by the launcher.
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const unsigned* p, int lane) {
    if (lane == 0)
        while (ld_acquire_u32(p) == 0) {
        }
    __syncwarp();
    (void)ld_acquire_u32(p);  // every lane acquires once the flag is known to be set
}

constexpr int WP = 130;  // smem pitch of the staged tile inverse

__device__ __forceinline__ void stage_winv(double* Ws, const double* __restrict__ W, int tid) {
    for (int idx = tid; idx < TILE * (TILE / 2); idx += 256) {
        const int r = idx >> 6, c2 = (idx & 63) * 2;
        *reinterpret_cast<double2*>(Ws + r * WP + c2) = *reinterpret_cast<const double2*>(W + r * TILE + c2);
    }
}

// z = L^-1 rhs.  Warp w owns rows 16w .. 16w+15 of the CTA's block row, lane l the columns
// 4l .. 4l+3 of every block: partial dot products stay per lane until the row is complete.
__global__ void __launch_bounds__(256, 1) trsv_fwd_chain_kernel(const double* __restrict__ L, int64_t ld,
                                                                const double* __restrict__ winv,
                                                                const double* __restrict__ rhs, double* z, int T,
                                                                unsigned* sync) {
    extern __shared__ __align__(16) double Ws[];  // [128][WP]
    __shared__ double ws[TILE];
    __shared__ int s_i;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_i = (int)atomicAdd(&sync[0], 1u);
    __syncthreads();
    const int i = s_i;
    stage_winv(Ws, winv + (int64_t)i * TILE * TILE, tid);
    const double* Lrow = L + ((int64_t)i * TILE + warp * 16) * ld + 4 * lane;
    double acc[16];
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) acc[rr] = 0.0;
    for (int J = 0; J < i; ++J) {
        double2 a0[16], a1[16];
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            const double2* p = reinterpret_cast<const double2*>(Lrow + (int64_t)rr * ld + (int64_t)J * TILE);
            a0[rr] = p[0];
            a1[rr] = p[1];
        }
        wait_flag(sync + 1 + J, lane);
        const double2 z0 = __ldcg(reinterpret_cast<const double2*>(z + (int64_t)J * TILE + 4 * lane));
        const double2 z1 = __ldcg(reinterpret_cast<const double2*>(z + (int64_t)J * TILE + 4 * lane + 2));
#pragma unroll
        for (int rr = 0; rr < 16; ++rr)
            acc[rr] += a0[rr].x * z0.x + a0[rr].y * z0.y + a1[rr].x * z1.x + a1[rr].y * z1.y;
    }
#pragma unroll
    for (int rr = 0; rr < 16; ++rr)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[rr] += __shfl_xor_sync(0xffffffffu, acc[rr], o);
    if (lane < 16) ws[warp * 16 + lane] = rhs[(int64_t)i * TILE + warp * 16 + lane] - pick(acc, lane);
    __syncthreads();  // ws complete, tile inverse staged
    {
        const double v0 = ws[4 * lane], v1 = ws[4 * lane + 1], v2 = ws[4 * lane + 2], v3 = ws[4 * lane + 3];
        double out[16];
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            const double* wr = Ws + (warp * 16 + rr) * WP + 4 * lane;
            out[rr] = wr[0] * v0 + wr[1] * v1 + wr[2] * v2 + wr[3] * v3;
        }
#pragma unroll
        for (int rr = 0; rr < 16; ++rr)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) out[rr] += __shfl_xor_sync(0xffffffffu, out[rr], o);
        if (lane < 16) z[(int64_t)i * TILE + warp * 16 + lane] = pick(out, lane);
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        st_release_u32(sync + 1 + i, 1u);
    }
}

// x = L^-T zin.  Same tiling of each block L[I, J] (warp: 16 rows, lane: 4 columns), but the sums
// run over rows, so every lane keeps its 4 column sums and the warps meet once at the end.
__global__ void __launch_bounds__(256, 1) trsv_bwd_chain_kernel(const double* __restrict__ L, int64_t ld,
                                                                const double* __restrict__ winv,
                                                                const double* __restrict__ zin, double* x, int T,
                                                                unsigned* sync) {
    extern __shared__ __align__(16) double Ws[];  // [128][WP]
    __shared__ double ws[TILE];
    __shared__ double part[8][TILE];
    __shared__ int s_i;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_i = (int)atomicAdd(&sync[0], 1u);
    __syncthreads();
    const int J = T - 1 - s_i;
    stage_winv(Ws, winv + (int64_t)J * TILE * TILE, tid);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int I = T - 1; I > J; --I) {
        const double* Lblk = L + ((int64_t)I * TILE + warp * 16) * ld + (int64_t)J * TILE + 4 * lane;
        double2 a0[16], a1[16];
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            const double2* p = reinterpret_cast<const double2*>(Lblk + (int64_t)rr * ld);
            a0[rr] = p[0];
            a1[rr] = p[1];
        }
        wait_flag(sync + 1 + I, lane);
        const double* xr = x + (int64_t)I * TILE + warp * 16;
#pragma unroll
        for (int rr = 0; rr < 16; rr += 2) {
            const double2 xv = __ldcg(reinterpret_cast<const double2*>(xr + rr));
            acc[0] += a0[rr].x * xv.x + a0[rr + 1].x * xv.y;
            acc[1] += a0[rr].y * xv.x + a0[rr + 1].y * xv.y;
            acc[2] += a1[rr].x * xv.x + a1[rr + 1].x * xv.y;
            acc[3] += a1[rr].y * xv.x + a1[rr + 1].y * xv.y;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) part[warp][4 * lane + q] = acc[q];
    __syncthreads();
    if (tid < TILE) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += part[w][tid];
        ws[tid] = zin[(int64_t)J * TILE + tid] - sum;
    }
    __syncthreads();
    {
        double o4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
            const double* wr = Ws + (warp * 16 + rr) * WP + 4 * lane;
            const double v = ws[warp * 16 + rr];
            o4[0] += wr[0] * v;
            o4[1] += wr[1] * v;
            o4[2] += wr[2] * v;
            o4[3] += wr[3] * v;
        }
        __syncthreads();  // part[] is reused
#pragma unroll
        for (int q = 0; q < 4; ++q) part[warp][4 * lane + q] = o4[q];
    }
    __syncthreads();
    if (tid < TILE) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += part[w][tid];
        x[(int64_t)J * TILE + tid] = sum;
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        st_release_u32(sync + 1 + J, 1u);
    }
}

__global__ void __launch_bounds__(1024) logdet_dot_kernel(const double* __restrict__ L, int64_t ld,
                                                          const double* __restrict__ y,
                                                          const double* __restrict__ alpha, int64_t N,
                                                          double* __restrict__ out) {
    __shared__ double a[1024], b[1024];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += 1024) {
        s0 += log(L[i * ld + i]);
        if (y) s1 += y[i] * alpha[i];
    }
    a[threadIdx.x] = s0;
    b[threadIdx.x] = s1;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            a[threadIdx.x] += a[threadIdx.x + o];
            b[threadIdx.x] += b[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = a[0];
        out[1] = b[0];
    }
}

__global__ void __launch_bounds__(256) row_reduce_kernel(const double* __restrict__ B, int64_t ld, int64_t rows,
                                                         int64_t cols, const double* __restrict__ v,
                                                         double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const double* row = B + r * ld;
    double s = 0.0;
    if (v) {
        for (int64_t c = lane; c < cols; c += 32) s += row[c] * v[c];
    } else {
        for (int64_t c = lane; c < cols; c += 32) s += row[c] * row[c];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) out[r] = s;
}

__global__ void gather_sym_kernel(const double* __restrict__ src, int64_t ld, const double* __restrict__ diag,
                                  int64_t N, double* __restrict__ out) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= N * N) return;
    const int64_t i = idx / N, j = idx % N;
    const int64_t hi = i > j ? i : j, lo = i > j ? j : i;
    const int64_t th = hi / TILE, tl = lo / TILE;
    out[idx] = (diag && th == tl) ? diag[th * TILE * TILE + (hi % TILE) * TILE + (lo % TILE)] : src[hi * ld + lo];
}

__global__ void fill_kernel(double* __restrict__ v, int64_t n, double value, int pattern) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (pattern) {
        uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 32;
        v[i] = (double)(h & 0xFFFFF) / 1048576.0 - 0.5;
    } else {
        v[i] = value;
    }
}

}  // namespace

__global__ void axpy_kernel(double* __restrict__ y, const double* __restrict__ x, double a, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) y[i] += a * x[i];
}
void launch_axpy(double* y, const double* x, double a, int64_t n, cudaStream_t s) {
    if (n > 0) axpy_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(y, x, a, n);
}

__global__ void copy_block_kernel(double* __restrict__ dst, int64_t ldd, const double* __restrict__ src, int64_t lds,
                                  int64_t rows, int cols2) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= rows * cols2) return;
    const int64_t r = i / cols2;
    const int c = (int)(i - r * cols2) * 2;
    *reinterpret_cast<double2*>(dst + r * ldd + c) = *reinterpret_cast<const double2*>(src + r * lds + c);
}
void launch_copy_block(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols,
                       cudaStream_t s) {
    const int64_t n = rows * (cols / 2);
    if (n > 0) copy_block_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(dst, ldd, src, lds, rows, (int)(cols / 2));
}

void launch_fill(double* v, int64_t n, double value, cudaStream_t s) {
    if (n > 0) fill_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(v, n, value, 0);
}
__global__ void fill_diag_kernel(double* v, int64_t stride, int n, double value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[(int64_t)i * stride] = value;
}
void launch_fill_diag(double* v, int64_t stride, int n, double value, cudaStream_t s) {
    fill_diag_kernel<<<(n + 127) / 128, 128, 0, s>>>(v, stride, n, value);
}
void launch_fill_pattern(double* v, int64_t n, cudaStream_t s) {
    if (n > 0) fill_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(v, n, 0.0, 1);
}

static int g_leaf_variant = -1;  // 0: blocked kernel (shipped), 1: first version (Crout, one barrier per column)
void set_leaf_variant(int v) { g_leaf_variant = v; }

void launch_potrf_leaf(double* A, int64_t ld, double* winv, int* info, int base, cudaStream_t s) {
    if (g_leaf_variant < 0) {
        const char* e = getenv("GOGP_LEAF");
        g_leaf_variant = e ? atoi(e) : 0;
    }
    const size_t smem_crout = (size_t)TILE * LP * sizeof(double);
    const size_t smem = ((size_t)TILE * LP + 32 * MP) * sizeof(double);
    static bool configured[64] = {false};  // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(potrf_leaf_crout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_crout);
        cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = true;
    }
    if (g_leaf_variant == 1)
        potrf_leaf_crout_kernel<<<1, 512, smem_crout, s>>>(A, ld, winv, info, base);
    else
        potrf_leaf_kernel<<<1, LEAF_THREADS, smem, s>>>(A, ld, winv, info, base);
}

void launch_trtri_leaf(const double* winv, double* dst, int64_t ld, cudaStream_t s) {
    trtri_leaf_kernel<<<16, 256, 0, s>>>(winv, dst, ld);
}

void launch_trsv_lower(const double* L, int64_t ld, const double* winv, double* rhs, double* out, int64_t Npad,
                       bool transposed, cudaStream_t s, int64_t* launches, unsigned* sync) {
    // rhs may be consumed (the step kernels use it as the running right-hand side); out must not alias it.
    const int T = (int)(Npad / TILE);
    static int chain = -1;  // 1: one launch per direction (default), 0: one launch per block (GOGP_TRSV=0)
    if (chain < 0) {
        const char* e = getenv("GOGP_TRSV");
        chain = e ? atoi(e) : 1;
    }
    if (chain && sync && T > 1) {
        const size_t smem = (size_t)TILE * WP * sizeof(double);
        static bool configured[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured[dev & 63]) {
            cudaFuncSetAttribute(trsv_fwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(trsv_bwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            configured[dev & 63] = true;
        }
        cudaMemsetAsync(sync, 0, (size_t)(T + 1) * sizeof(unsigned), s);
        if (!transposed)
            trsv_fwd_chain_kernel<<<T, 256, smem, s>>>(L, ld, winv, rhs, out, T, sync);
        else
            trsv_bwd_chain_kernel<<<T, 256, smem, s>>>(L, ld, winv, rhs, out, T, sync);
        if (launches) *launches += 1;
        return;
    }
    if (!transposed) {
        trsv_first_kernel<<<1, 256, 0, s>>>(winv, rhs, out, 0, 0);
        for (int I = 0; I + 1 < T; ++I) trsv_fwd_step_kernel<<<T - 1 - I, 256, 0, s>>>(L, ld, winv, rhs, out, I);
    } else {
        trsv_first_kernel<<<1, 256, 0, s>>>(winv, rhs, out, T - 1, 1);
        for (int I = T - 1; I > 0; --I) trsv_bwd_step_kernel<<<I, 256, 0, s>>>(L, ld, winv, rhs, out, I);
    }
    if (launches) *launches += T;
}

void launch_logdet_dot(const double* L, int64_t ld, const double* y, const double* alpha, int64_t N, double* out,
                       cudaStream_t s) {
    logdet_dot_kernel<<<1, 1024, 0, s>>>(L, ld, y, alpha, N, out);
}

void launch_row_reduce(const double* B, int64_t ld, int64_t rows, int64_t cols, const double* v, double* out,
                       cudaStream_t s) {
    if (rows <= 0) return;
    row_reduce_kernel<<<(int)((rows + 7) / 8), 256, 0, s>>>(B, ld, rows, cols, v, out);
}

void launch_gather_sym(const double* src, int64_t ld, const double* diag, int64_t N, double* out, cudaStream_t s) {
    if (N <= 0) return;
    const int64_t total = N * N;
    gather_sym_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(src, ld, diag, N, out);
}

}  // namespace gogp
