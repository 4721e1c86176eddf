// Tile-level (128x128) and vector kernels around the DMMA GEMM: diagonal-block
// Cholesky + inverse, blocked triangular solves for alpha = K^-1 y
// (gp/gp.go:232-236), log-determinant and y.alpha (gp/gp.go:250-251), and small
// reductions used by Produce (gp/gp.go:335-357).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>

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

#include "leaf_kernels.cuh"

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

__global__ void __launch_bounds__(1024) logdet_dot_kernel(const double* __restrict__ L, int64_t ld,
                                                          const double* __restrict__ y,
                                                          const double* __restrict__ alpha, int64_t N,
                                                          double* __restrict__ out) {
    __shared__ double a[1024], b[1024], lo[1024], hi[1024];
    double s0 = 0.0, s1 = 0.0, mn = INFINITY, mx = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += 1024) {
        const double d = L[i * ld + i];
        s0 += log(d);
        mn = fmin(mn, d);
        mx = fmax(mx, d);
        if (y) s1 += y[i] * alpha[i];
    }
    a[threadIdx.x] = s0;
    b[threadIdx.x] = s1;
    lo[threadIdx.x] = mn;
    hi[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            a[threadIdx.x] += a[threadIdx.x + o];
            b[threadIdx.x] += b[threadIdx.x + o];
            lo[threadIdx.x] = fmin(lo[threadIdx.x], lo[threadIdx.x + o]);
            hi[threadIdx.x] = fmax(hi[threadIdx.x], hi[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = a[0];
        out[1] = b[0];
        out[2] = lo[0];  // extreme diagonal entries of L: (max / min)^2 is a lower bound of cond_2(K)
        out[3] = hi[0];
    }
}

// ---- pieces of the condition estimate (capi.cu cond_estimate; rare path, plain bandwidth kernels) ----
// w[j] = sum_{i >= j} L[i][j] v[i]  (L^T v): a thread per column, coalesced across the threads of a row
__global__ void __launch_bounds__(256) trmv_t_kernel(const double* __restrict__ L, int64_t ld,
                                                     const double* __restrict__ v, double* __restrict__ w, int64_t N) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= N) return;
    double s = 0.0;
    for (int64_t i = j; i < N; ++i) s = fma(L[i * ld + j], v[i], s);
    w[j] = s;
}
// u[i] = sum_{j <= i} L[i][j] w[j]  (L w): a warp per row
__global__ void __launch_bounds__(256) trmv_kernel(const double* __restrict__ L, int64_t ld,
                                                   const double* __restrict__ w, double* __restrict__ u, int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= N) return;
    double s = 0.0;
    for (int64_t j = lane; j <= i; j += 32) s = fma(L[i * ld + j], w[j], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) u[i] = s;
}
// W = T^-1 for the 128 x 128 lower-triangular diagonal tiles of L (gogp_set_state: the tile inverses the
// blocked solves multiply with are not part of the stored state): one CTA per tile, a thread per column of W,
// forward substitution through shared memory.
__global__ void __launch_bounds__(TILE) tile_inverse_kernel(const double* __restrict__ L, int64_t ld,
                                                            double* __restrict__ winv) {
    extern __shared__ double W[];  // column-major [128][129]: thread c owns column c
    const int t = blockIdx.x, c = threadIdx.x;
    const double* Lt = L + (int64_t)t * TILE * ld + (int64_t)t * TILE;  // rows are read by all threads at once: broadcast
    double* wc = W + c * (TILE + 1);
    for (int i = 0; i < TILE; ++i) {
        double s = i == c ? 1.0 : 0.0;
        for (int k = c; k < i; ++k) s = fma(-Lt[(int64_t)i * ld + k], wc[k], s);
        wc[i] = i >= c ? s / Lt[(int64_t)i * ld + i] : 0.0;
    }
    __syncthreads();
    double* out = winv + (int64_t)t * TILE * TILE;
    for (int r = 0; r < TILE; ++r) out[r * TILE + c] = W[c * (TILE + 1) + r];
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

// variant 0: blocked kernel (shipped), 1: first version (Crout, one barrier per column), 2: blocked kernel with
// the shared-memory column broadcast (timing candidate); -1: the process default (GOGP_LEAF, else 0).  The variant
// is a launch argument, not a global: handles on other host threads are not affected by a timing run.
void launch_potrf_leaf(double* A, int64_t ld, double* winv, int* info, int base, cudaStream_t s, int variant) {
    static int default_variant = 0;
    static std::once_flag knob;
    std::call_once(knob, [] {
        const char* e = getenv("GOGP_LEAF");
        default_variant = e ? atoi(e) : 0;
    });
    const int g_leaf_variant = variant >= 0 ? variant : default_variant;
    const size_t smem_crout = (size_t)TILE * LP * sizeof(double);
    const size_t smem = ((size_t)TILE * LP + 32 * MP + 64) * sizeof(double);
    static std::atomic<bool> configured[64];  // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63].load(std::memory_order_acquire)) {
        cudaFuncSetAttribute(potrf_leaf_crout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_crout);
        cudaFuncSetAttribute(potrf_leaf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(potrf_leaf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63].store(true, std::memory_order_release);
    }
    if (g_leaf_variant == 1)
        potrf_leaf_crout_kernel<<<1, 512, smem_crout, s>>>(A, ld, winv, info, base);
    else if (g_leaf_variant == 2)
        potrf_leaf_kernel<true><<<1, LEAF_THREADS, smem, s>>>(A, ld, winv, info, base);
    else
        potrf_leaf_kernel<false><<<1, LEAF_THREADS, smem, s>>>(A, ld, winv, info, base);
}

void launch_trtri_leaf(const double* winv, double* dst, int64_t ld, cudaStream_t s) {
    trtri_leaf_kernel<<<16, 256, 0, s>>>(winv, dst, ld);
}

void launch_trsv_lower(const double* L, int64_t ld, const double* winv, double* rhs, double* out, int64_t Npad,
                       bool transposed, cudaStream_t s, int64_t* launches, unsigned* sync) {
    // rhs may be consumed (the step kernels use it as the running right-hand side); out must not alias it.
    const int T = (int)(Npad / TILE);
    static int chain = 1;  // 1: one launch per direction (default), 0: one launch per block (GOGP_TRSV=0)
    static std::once_flag knob;
    std::call_once(knob, [] {
        const char* e = getenv("GOGP_TRSV");
        chain = e ? atoi(e) : 1;
    });
    if (chain && sync && T > 1) {
        const size_t smem = (size_t)TILE * WP * sizeof(double);
        static std::atomic<bool> configured[64];
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured[dev & 63].load(std::memory_order_acquire)) {
            cudaFuncSetAttribute(trsv_fwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(trsv_bwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            configured[dev & 63].store(true, std::memory_order_release);
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

void launch_trmv_lower(const double* L, int64_t ld, const double* v, double* out, int64_t N, bool transposed,
                       cudaStream_t s) {
    if (N <= 0) return;
    if (transposed)
        trmv_t_kernel<<<(int)((N + 255) / 256), 256, 0, s>>>(L, ld, v, out, N);
    else
        trmv_kernel<<<(int)((N + 7) / 8), 256, 0, s>>>(L, ld, v, out, N);
}

void launch_tile_inverse(const double* L, int64_t ld, double* winv, int tiles, cudaStream_t s) {
    if (tiles <= 0) return;
    const size_t smem = (size_t)TILE * (TILE + 1) * sizeof(double);
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63].load(std::memory_order_acquire)) {
        cudaFuncSetAttribute(tile_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63].store(true, std::memory_order_release);
    }
    tile_inverse_kernel<<<tiles, TILE, smem, s>>>(L, ld, winv);
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
