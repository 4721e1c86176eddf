// TEST INFRASTRUCTURE: the product's tile kernel and single-launch triangular solves
// (gogp_b200/csrc/leaf_kernels.cuh, unmodified source) compiled for the host under the SIMT emulator.
#define GOGP_SIMT_HOST 1
#include "simt.h"

namespace gogp {
constexpr int TILE = 128;
constexpr int LP = 132;  // as in leaf.cu
#include "../../gogp_b200/csrc/leaf_kernels.cuh"
}  // namespace gogp

extern "C" {

// A: 128 x 128 row-major with leading dimension ld, factored in place; winv: 128 x 128; *info as the kernel leaves it.
void simt_potrf_leaf(double* A, int64_t ld, double* winv, int* info, int base, int variant) {
    using namespace gogp;
    const size_t smem = ((size_t)TILE * LP + 32 * MP + 64) * sizeof(double);
    if (variant == 2)
        simt::launch(1, LEAF_THREADS, smem, [&] { potrf_leaf_kernel<true>(A, ld, winv, info, base); });
    else
        simt::launch(1, LEAF_THREADS, smem, [&] { potrf_leaf_kernel<false>(A, ld, winv, info, base); });
}

// out = L^-1 rhs (transposed = 0) or L^-T rhs (1); L: Npad x Npad lower, winv: T tile inverses; sync: T+1 words.
void simt_trsv(const double* L, int64_t ld, const double* winv, const double* rhs, double* out, int T, int transposed,
               unsigned* sync) {
    using namespace gogp;
    const size_t smem = (size_t)TILE * WP * sizeof(double);
    for (int i = 0; i <= T; ++i) sync[i] = 0;
    if (!transposed)
        simt::launch((unsigned)T, 256, smem, [&] { trsv_fwd_chain_kernel(L, ld, winv, rhs, out, T, sync); });
    else
        simt::launch((unsigned)T, 256, smem, [&] { trsv_bwd_chain_kernel(L, ld, winv, rhs, out, T, sync); });
}
}
