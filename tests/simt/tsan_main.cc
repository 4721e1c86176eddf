// TEST INFRASTRUCTURE: driver of the ThreadSanitizer run (tests/test_simt_kernels.py): the tile kernel and both
// single-launch solves under the SIMT emulator; any "data race" report fails the test.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
extern "C" void simt_potrf_leaf(double* A, int64_t ld, double* winv, int* info, int base, int variant);
extern "C" void simt_trsv(const double* L, int64_t ld, const double* winv, const double* rhs, double* out, int T, int transposed, unsigned* sync);
int main() {
    const int n = 128;
    std::vector<double> A(n * n), W(n * n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A[i * n + j] = (i == j) ? 2.0 + 0.01 * i : 0.3 / (1.0 + std::abs(i - j));
    int info = 0;
    {   // the shared-memory broadcast variant first, on a copy
        std::vector<double> A2(A), W2(n * n);
        simt_potrf_leaf(A2.data(), n, W2.data(), &info, 0, 2);
        std::printf("variant 2 info %d L00 %.6f W127 %.6f\n", info, A2[0], W2[127 * n + 127]);
    }
    simt_potrf_leaf(A.data(), n, W.data(), &info, 0, 0);
    std::printf("leaf info %d L00 %.6f W127 %.6f\n", info, A[0], W[127 * n + 127]);
    // two-tile solve: L = [[A,0],[B,A]] with B small
    const int T = 2, N = T * n;
    std::vector<double> L(N * N, 0.0), winv(T * n * n), rhs(N, 1.0), out(N, 0.0);
    for (int t = 0; t < T; ++t) for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { L[(t*n+i)*N + t*n+j] = A[i*n+j]; winv[t*n*n + i*n + j] = W[i*n+j]; }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) L[(n+i)*N + j] = 0.01 * ((i * 7 + j * 3) % 11);
    std::vector<unsigned> sync(T + 1);
    simt_trsv(L.data(), N, winv.data(), rhs.data(), out.data(), T, 0, sync.data());
    std::printf("fwd out0 %.6f out255 %.6f\n", out[0], out[N-1]);
    simt_trsv(L.data(), N, winv.data(), rhs.data(), out.data(), T, 1, sync.data());
    std::printf("bwd out0 %.6f out255 %.6f\n", out[0], out[N-1]);
    return 0;
}
