// TEST INFRASTRUCTURE: a naive host backend for gogp_b200/csrc/blocked.hpp, so
// the recursion's index arithmetic (which tile, which k range, which operand) is
// checked on a CPU-only machine.  It mirrors the tile semantics of the CUDA
// kernels (dgemm.cu, leaf.cu) but shares no code with them and is never linked
// into libgogp_b200.so.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../gogp_b200/csrc/blocked.hpp"

using namespace gogp;

namespace {

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

}  // namespace

static int64_t g_rl_max = 0;

extern "C" {

// diagonal blocks up to this size take the short-chain variants (0: pure recursion)
void cpu_blocked_set_rl_max(int64_t v) { g_rl_max = v; }

// A: n x n row-major (lower used), factored in place; winv: (n/128) x 128 x 128.
int cpu_blocked_potrf(double* A, int64_t n, double* winv) {
    HostBackend be;
    Blocked<HostBackend> bl{be, A, n, winv, g_rl_max, g_rl_max};
    bl.potrf(0, n);
    return be.info;
}

// The look-ahead (flat right-looking) factorisation with nb-wide block columns.
int cpu_blocked_potrf_la(double* A, int64_t n, double* winv, int64_t nb0, int64_t nb1, int64_t nb2) {
    HostBackend be;
    Blocked<HostBackend> bl{be, A, n, winv, g_rl_max, g_rl_max};
    bl.la_nb[0] = nb0;
    bl.la_nb[1] = nb1;
    bl.la_nb[2] = nb2;
    bl.potrf_la(0, n, 0);
    return be.info;
}

// L (from cpu_blocked_potrf) -> K^-1: strictly-lower tiles of Bm, diagonal tiles in dg.
void cpu_blocked_kinv(double* L, int64_t n, double* winv, double* Bm, double* dg) {
    HostBackend be;
    Blocked<HostBackend> bl{be, L, n, winv, g_rl_max, g_rl_max};
    if (g_rl_max == 1024)
        bl.trtri_t_levels(Bm, n, 256);  // level-order traversal, small parallel blocks
    else
        bl.trtri_t(Bm, 0, n);
    bl.lauum(Bm, dg, n);
}

// Only the transposed inverse U = L^-T (upper triangle of Bm).
void cpu_blocked_trtri(double* L, int64_t n, double* winv, double* Bm) {
    HostBackend be;
    Blocked<HostBackend> bl{be, L, n, winv, g_rl_max, g_rl_max};
    bl.trtri_t(Bm, 0, n);
}

// B (m x n) <- B L^-T
void cpu_blocked_trsm(double* L, int64_t n, double* winv, double* B, int64_t m) {
    HostBackend be;
    Blocked<HostBackend> bl{be, L, n, winv, g_rl_max, g_rl_max};
    bl.trsm(B, n, m, 0, n);
}
}
