// Tile-level (128x128) and vector kernels around the DMMA GEMM: diagonal-block
// Cholesky + inverse, blocked triangular solves for alpha = K^-1 y
// (gp/gp.go:232-236), log-determinant and y.alpha (gp/gp.go:250-251), and small
// reductions used by Produce (gp/gp.go:335-357).
#include <cstdio>

#include "kernels.h"

namespace gogp {

namespace {

constexpr int LP = 129;  // smem pitch (doubles): odd -> column walks are conflict-free

// Right-looking Cholesky of one 128x128 tile held in shared memory, fused with a
// Gauss-Jordan build of the inverse: after column j of L is final, row j of
// W = L^-1 is final too, and the rank-1 step that updates the trailing block
// of A also updates rows i > j of W.  W[i][c] (c <= i) lives at S[c][i+1], the
// unused upper triangle of the same array.
__global__ void __launch_bounds__(512, 1) potrf_leaf_kernel(double* __restrict__ A, int64_t ld,
                                                            double* __restrict__ winv, int* __restrict__ info,
                                                            int base) {
    extern __shared__ double S[];  // [128][LP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < TILE * TILE; idx += 512) {
        const int r = idx >> 7, c = idx & 127;
        S[r * LP + c] = (c <= r) ? A[(int64_t)r * ld + c] : 0.0;
    }
    if (tid < TILE) S[tid * LP + TILE] = 0.0;
    __syncthreads();
    if (tid < TILE) S[tid * LP + tid + 1] = 1.0;  // W = I
    __syncthreads();

    for (int j = 0; j < TILE; ++j) {
        double ajj = S[j * LP + j];
        if (!(ajj > 0.0)) {  // not positive definite (or NaN): flag and keep going
            if (tid == 0 && atomicCAS(info, 0, base + j + 1) == 0) {
            }
            ajj = 1.0;
        }
        const double d = sqrt(ajj);
        __syncthreads();  // everyone has read the pivot
        if (tid < TILE) {
            if (tid > j)
                S[tid * LP + j] /= d;
            else if (tid == j)
                S[j * LP + j] = d;
        } else if (tid < 2 * TILE) {
            const int c = tid - TILE;
            if (c <= j) S[c * LP + j + 1] /= d;
        }
        __syncthreads();
        for (int i = j + 1 + warp; i < TILE; i += 16) {
            const double lij = S[i * LP + j];
            for (int k = j + 1 + lane; k <= i; k += 32) S[i * LP + k] -= lij * S[k * LP + j];
            for (int c = lane; c <= j; c += 32) S[c * LP + i + 1] -= lij * S[c * LP + j + 1];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < TILE * TILE; idx += 512) {
        const int r = idx >> 7, c = idx & 127;
        A[(int64_t)r * ld + c] = (c <= r) ? S[r * LP + c] : 0.0;
        winv[idx] = (c <= r) ? S[c * LP + r + 1] : 0.0;
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

// One step of the blocked forward solve  L z = y  (w is the running right-hand
// side): every CTA recomputes z_I = Winv_I w_I; CTA 0 stores it, CTA b >= 1
// updates w_{I+b} -= L[I+b, I] z_I.
__global__ void __launch_bounds__(256) trsv_fwd_step_kernel(const double* __restrict__ L, int64_t ld,
                                                            const double* __restrict__ winv, double* __restrict__ w,
                                                            double* __restrict__ z, int I) {
    __shared__ double ws[TILE], zs[TILE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < TILE) ws[tid] = w[(int64_t)I * TILE + tid];
    __syncthreads();
    const double* Wi = winv + (int64_t)I * TILE * TILE;
    for (int r = warp * 16; r < warp * 16 + 16; ++r) {
        double s = 0.0;
        for (int c = lane; c <= r; c += 32) s += Wi[r * TILE + c] * ws[c];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) zs[r] = s;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        if (tid < TILE) z[(int64_t)I * TILE + tid] = zs[tid];
        return;
    }
    const int64_t rb = ((int64_t)I + blockIdx.x) * TILE;
    const double* Lb = L + rb * ld + (int64_t)I * TILE;
    for (int r = warp * 16; r < warp * 16 + 16; ++r) {
        double s = 0.0;
        for (int c = lane; c < TILE; c += 32) s += Lb[(int64_t)r * ld + c] * zs[c];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) w[rb + r] -= s;
    }
}

// One step of the blocked backward solve  L^T x = z: x_I = Winv_I^T w_I; CTA b >= 1
// updates w_J -= L[I, J]^T x_I for J = b - 1 < I.
__global__ void __launch_bounds__(256) trsv_bwd_step_kernel(const double* __restrict__ L, int64_t ld,
                                                            const double* __restrict__ winv, double* __restrict__ w,
                                                            double* __restrict__ x, int I) {
    __shared__ double ws[TILE], xs[TILE], part[TILE];
    const int tid = threadIdx.x;
    if (tid < TILE) ws[tid] = w[(int64_t)I * TILE + tid];
    __syncthreads();
    const double* Wi = winv + (int64_t)I * TILE * TILE;
    {
        // x[c] = sum_{r >= c} Winv[r][c] w[r]; two threads per column split the rows
        const int c = tid & 127, half = tid >> 7;
        double s = 0.0;
        for (int r = c + half; r < TILE; r += 2) s += Wi[r * TILE + c] * ws[r];
        if (half) part[c] = s;
        __syncthreads();
        if (!half) xs[c] = s + part[c];
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        if (tid < TILE) x[(int64_t)I * TILE + tid] = xs[tid];
        return;
    }
    const int64_t J = blockIdx.x - 1;
    const double* Lb = L + (int64_t)I * TILE * ld + J * TILE;
    const int c = tid & 127, half = tid >> 7;
    double s = 0.0;
    for (int r = half * 64; r < half * 64 + 64; ++r) s += Lb[(int64_t)r * ld + c] * xs[r];
    if (half) part[c] = s;
    __syncthreads();
    if (!half) w[J * TILE + c] -= s + part[c];
}

__global__ void __launch_bounds__(1024) logdet_dot_kernel(const double* __restrict__ L, int64_t ld,
                                                          const double* __restrict__ y,
                                                          const double* __restrict__ alpha, int64_t N,
                                                          double* __restrict__ out) {
    __shared__ double a[1024], b[1024];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += 1024) {
        s0 += log(L[i * ld + i]);
        s1 += y[i] * alpha[i];
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

void launch_fill(double* v, int64_t n, double value, cudaStream_t s) {
    if (n > 0) fill_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(v, n, value, 0);
}
void launch_fill_pattern(double* v, int64_t n, cudaStream_t s) {
    if (n > 0) fill_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(v, n, 0.0, 1);
}

void launch_potrf_leaf(double* A, int64_t ld, double* winv, int* info, int base, cudaStream_t s) {
    const size_t smem = (size_t)TILE * LP * sizeof(double);
    static bool configured[64] = {false};  // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev & 63] = true;
    }
    potrf_leaf_kernel<<<1, 512, smem, s>>>(A, ld, winv, info, base);
}

void launch_trtri_leaf(const double* winv, double* dst, int64_t ld, cudaStream_t s) {
    trtri_leaf_kernel<<<16, 256, 0, s>>>(winv, dst, ld);
}

void launch_trsv_lower(const double* L, int64_t ld, const double* winv, double* rhs, double* out, int64_t Npad,
                       bool transposed, cudaStream_t s, int64_t* launches) {
    // rhs is consumed (it is the running right-hand side); out must not alias it:
    // CTA 0 of a step stores block I of the result while the others still read rhs_I.
    const int T = (int)(Npad / TILE);
    if (!transposed) {
        for (int I = 0; I < T; ++I) trsv_fwd_step_kernel<<<T - I, 256, 0, s>>>(L, ld, winv, rhs, out, I);
    } else {
        for (int I = T - 1; I >= 0; --I) trsv_bwd_step_kernel<<<I + 1, 256, 0, s>>>(L, ld, winv, rhs, out, I);
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
