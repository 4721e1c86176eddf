// The two most intricate kernels of the path -- the blocked 128x128 tile Cholesky + inverse and
// the single-launch (ticket / release-acquire) triangular solves -- in a header of their own, so that
// besides leaf.cu (which includes it inside its anonymous namespace) the CPU test tier can compile the
// SAME source for the host under the lock-step SIMT emulator of tests/simt/ (GOGP_SIMT_HOST):
// tests/test_simt_kernels.py runs them against NumPy on a machine without a GPU, optionally under
// ThreadSanitizer.  Under nvcc nothing changes: the device SASS of leaf.cu is byte-identical to the
// version that had these kernels inline.
#pragma once

#if defined(GOGP_SIMT_HOST)
#define GOGP_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(simt::dyn_smem())
#else
#define GOGP_DYN_SMEM(T, name) extern __shared__ __align__(16) T name[]
#endif

__device__ __forceinline__ void dmma_leaf(double& c0, double& c1, double a, double b) {
#if defined(GOGP_SIMT_HOST)
    simt::dmma_m8n8k4(c0, c1, a, b);
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#endif
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
// BCAST (variant 2, timing candidate): the scaled column of phase A reaches the other lanes through a
// 2 x 32-double shared-memory buffer (one store, one __syncwarp, broadcast loads) instead of 2 (31 - j)
// shuffles per column; same operations on the same values, bit-identical results.
template <bool BCAST>
__global__ void __launch_bounds__(LEAF_THREADS, 1) potrf_leaf_kernel(double* __restrict__ A, int64_t ld,
                                                            double* __restrict__ winv, int* __restrict__ info,
                                                            int base) {
    GOGP_DYN_SMEM(double, S);  // [128][LP], then the side buffer [32][MP]
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
            if (!BCAST) {
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
            } else {
                double* colbuf = Mb + 32 * MP;  // [2][32], after the side buffer
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    double d = __shfl_sync(FULL, a[j], j);
                    if (!(d > 0.0)) {
                        if (lane == j) {
                            atomicCAS(info, 0, base + o + j + 1);
                            a[j] = 1.0;
                        }
                        d = 1.0;
                    }
                    const double rs = rsqrt(d);
                    const double l = a[j] * rs;
                    a[j] = l;
                    if (lane == j) rinv = rs;
                    double* cb = colbuf + (j & 1) * 32;  // column j+2 reuses this half: every lane is past j+1's barrier by then
                    cb[lane] = l;
                    __syncwarp();
#pragma unroll
                    for (int k = j + 1; k < 32; ++k) a[k] = fma(-l, cb[k], a[k]);
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


// out[lane] without dynamic register indexing
template <int ROWS>
__device__ __forceinline__ double pick(const double (&out)[ROWS], int lane) {
    double v = 0.0;
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr)
        if (lane == rr) v = out[rr];
    return v;
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
#if defined(GOGP_SIMT_HOST)
    return simt::ld_acquire(p);
#else
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
#if defined(GOGP_SIMT_HOST)
    simt::st_release(p, v);
#else
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
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
    GOGP_DYN_SMEM(double, Ws);  // [128][WP]
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
    GOGP_DYN_SMEM(double, Ws);  // [128][WP]
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

