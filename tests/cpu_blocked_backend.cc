// TEST INFRASTRUCTURE: a naive host backend for gogp_b200/csrc/blocked.hpp, so
// the recursion's index arithmetic (which tile, which k range, which operand) is
// checked on a CPU-only machine.  It mirrors the tile semantics of the CUDA
// kernels (dgemm.cu, leaf.cu) but shares no code with them and is never linked
// into libgogp_b200.so.
#include "host_tiles.h"

using namespace gogp;
using gogp_host::HostBackend;

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
