// TEST INFRASTRUCTURE: driver of the ThreadSanitizer run of the TMA GEMM kernel under the SIMT emulator
// (tests/test_simt_gemm.py): a k long enough for the 6-stage ring to wrap, plain and in-place modes.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>
extern "C" void simt_dgemm_tma(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb,
                               int64_t m, int64_t n, int64_t k, double alpha, double beta, int mode, double* cdiag);
int main() {
    const int m = 128, n = 128, k = 256;
    std::vector<double> A(m * k), B(n * k), C(m * n, 1.0);
    for (int i = 0; i < m * k; ++i) A[i] = std::sin(0.01 * i);
    for (int i = 0; i < n * k; ++i) B[i] = std::cos(0.02 * i);
    simt_dgemm_tma(C.data(), n, A.data(), k, B.data(), k, m, n, k, -1.0, 1.0, 0, nullptr);
    double ref = 1.0;
    for (int kk = 0; kk < k; ++kk) ref -= A[5 * k + kk] * B[7 * k + kk];
    std::printf("full C[5][7] %.12f ref %.12f\n", C[5 * n + 7], ref);
    // in place: X = P W^T with C aliasing A (mode 8), P is m x 128 inside a wider panel
    std::vector<double> P(m * 256, 0.5), W(128 * 128, 0.0);
    for (int i = 0; i < 128; ++i) for (int j = 0; j <= i; ++j) W[i * 128 + j] = 1.0 / (1.0 + i + j);
    simt_dgemm_tma(P.data() + 128, 256, P.data() + 128, 256, W.data(), 128, m, 128, 128, 1.0, 0.0, 8, nullptr);
    std::printf("inplace P[3][128] %.12f\n", P[3 * 256 + 128]);
    return std::fabs(C[5 * n + 7] - ref) < 1e-10 ? 0 : 1;
}
