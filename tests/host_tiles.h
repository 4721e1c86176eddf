// TEST INFRASTRUCTURE: a naive host backend for gogp_b200/csrc/blocked.hpp (tile GEMM with the kernels' tile-map
// modes, 128 x 128 Cholesky + inverse, transposed tile inverse).  It mirrors the tile semantics of the CUDA
// kernels (dgemm.cu, leaf.cu) but shares no code with them and is never linked into libgogp_b200.so.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../gogp_b200/csrc/blocked.hpp"

namespace gogp_host {
using namespace gogp;

struct HostBackend {
    int info = 0;
    long gemms = 0, leaves = 0;
    void gemm(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t m, int64_t n,
              int64_t k, double alpha, double beta, int mode, double* cdiag) {
        ++gemms;
        const int64_t tm = m / kTile, tn = n / kTile;
        std::vector<double> tile(kTile * kTile);
        for (int64_t ti = 0; ti < tm; ++ti)
            for (int64_t tj = 0; tj < tn; ++tj) {
                if ((mode & BL_LOWER) && tj > ti) continue;
                const int64_t k_lo = (mode & BL_KTRI) ? ti * kTile : 0;
                // whole tile first, then the store: C may alias A (in-place solve with a diagonal block)
                for (int64_t i = 0; i < kTile; ++i)
                    for (int64_t j = 0; j < kTile; ++j) {
                        double s = 0.0;
                        const double* a = A + (ti * kTile + i) * lda;
                        const double* b = B + (tj * kTile + j) * ldb;
                        for (int64_t kk = k_lo; kk < k; ++kk) s += a[kk] * b[kk];
                        tile[i * kTile + j] = s;
                    }
                double* Cb;
                int64_t ld;
                if ((mode & BL_DIAG_OUT) && ti == tj) {
                    Cb = cdiag + ti * kTile * kTile;
                    ld = kTile;
                } else {
                    Cb = C + ti * kTile * ldc + tj * kTile;
                    ld = ldc;
                }
                for (int64_t i = 0; i < kTile; ++i)
                    for (int64_t j = 0; j < kTile; ++j) {
                        double v = alpha * tile[i * kTile + j];
                        if (beta != 0.0) v += beta * Cb[i * ld + j];
                        Cb[i * ld + j] = v;
                    }
            }
    }
    void potrf_leaf(double* A, int64_t ld, double* winv, int base) {
        ++leaves;
        const int n = (int)kTile;
        for (int j = 0; j < n; ++j) {
            double d = A[j * ld + j];
            for (int k = 0; k < j; ++k) d -= A[j * ld + k] * A[j * ld + k];
            if (!(d > 0.0)) {
                if (!info) info = base + j + 1;
                d = 1.0;
            }
            d = std::sqrt(d);
            A[j * ld + j] = d;
            for (int i = j + 1; i < n; ++i) {
                double s = A[i * ld + j];
                for (int k = 0; k < j; ++k) s -= A[i * ld + k] * A[j * ld + k];
                A[i * ld + j] = s / d;
            }
        }
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) A[i * ld + j] = 0.0;
        // inverse by forward substitution, column by column
        std::memset(winv, 0, sizeof(double) * n * n);
        for (int c = 0; c < n; ++c)
            for (int i = c; i < n; ++i) {
                double s = (i == c) ? 1.0 : 0.0;
                for (int k = c; k < i; ++k) s -= A[i * ld + k] * winv[k * n + c];
                winv[i * n + c] = s / A[i * ld + i];
            }
    }
    void par_begin() {}
    void par_use(int) {}
    void par_end() {}
    void set_scratch_row(int64_t) {}
    void la_fork(int) {}
    void la_bulk_begin(int) {}
    void la_mark_below(int) {}
    void la_bulk_end(int) {}
    void la_wait_bulk(int) {}
    void la_wait_below(int) {}
    void trtri_leaf(const double* winv, double* dst, int64_t ld) {
        const int n = (int)kTile;
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) dst[r * ld + c] = (c >= r) ? winv[c * n + r] : 0.0;
    }
};

}  // namespace gogp_host
