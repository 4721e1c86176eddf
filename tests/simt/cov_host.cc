// TEST INFRASTRUCTURE: the product's covariance-side kernels (gogp_b200/csrc/cov_kernels.cuh, unmodified
// source: build -- interpreted and specialised --, fused gradient trace incl. the block mode of the
// distributed path, input gradient) compiled for the host under the SIMT emulator, driven through the
// product's own descriptor lowering (csrc/program.cc).
#define GOGP_SIMT_HOST 1
#include "simt.h"

#include <string>
#include <vector>

#include "../../gogp_b200/csrc/kexpr.cuh"
#include "../../gogp_b200/csrc/program.cc"

namespace gogp {
constexpr int TILE = 128;
#include "../../gogp_b200/csrc/cov_kernels.cuh"
}  // namespace gogp

using namespace gogp;

namespace {
struct Problem {
    DevProgram prog;
    int D;
    int64_t N, Npad;
    std::vector<double> Xt;
};
bool setup(Problem& p, const gogp_op* ops, int nops, int ntheta, int ndim, const double* theta, const double* events,
           int nev, const double* X, int64_t N) {
    Program pr;
    std::string err;
    if (!pr.lower(ops, nops, ntheta, ndim, true, &err)) return false;
    if (nev > 0) pr.events.assign(events, events + 3 * nev);
    pr.bind(theta, &p.prog);
    p.D = ndim;
    p.N = N;
    p.Npad = (N + TILE - 1) / TILE * TILE;
    p.Xt.assign((size_t)p.Npad * ndim, 0.0);
    std::vector<double> Xc(X, X + N * ndim);
    double* Xt = p.Xt.data();
    const double* Xp = Xc.data();
    const int64_t Npad = p.Npad;
    simt::launch((unsigned)((Npad + 255) / 256), 256, 0, [&] { transpose_x_kernel(Xp, Xt, N, Npad, ndim); });
    return true;
}
size_t cov_smem(int D) { return (size_t)2 * D * TILE * sizeof(double) + 16; }
}  // namespace

extern "C" {

// out: Npad x Npad row-major (caller pre-fills it; only lower tiles are written).  fast: 0 interpreted kernel,
// 1 the specialised kernel (returns 2 when the program has no fast shape).
int simt_cov_build(const gogp_op* ops, int nops, int ntheta, int ndim, const double* theta, const double* events,
                   int nev, const double* X, int64_t N, double noise, int fast, double* out) {
    Problem p;
    if (!setup(p, ops, nops, ntheta, ndim, theta, events, nev, X, N)) return 1;
    const int T = (int)(p.Npad / TILE), ntiles = T * (T + 1) / 2;
    const double* Xt = p.Xt.data();
    const DevProgram& prog = p.prog;
    const int D = ndim;
    const int64_t Npad = p.Npad;
    if (!fast) {
        simt::launch((unsigned)ntiles, 256, cov_smem(D),
                     [&] { cov_tile_kernel<true>(prog, Xt, Npad, N, Xt, Npad, N, D, noise, out, Npad, T); });
        return 0;
    }
    if (prog.nterms != 1) return 2;
    const int nn = prog.nnorm[0], rest = prog.fbeg[1] - nn;
    if (rest < 0 || rest > kFastMaxRest) return 2;
#define CASE(NNV)                                                                                              \
    case NNV:                                                                                                  \
        simt::launch((unsigned)ntiles, 256, cov_smem(D), [&] {                                                 \
            cov_tile_fast_kernel<NNV, true>(prog, Xt, Npad, N, Xt, Npad, N, D, noise, out, Npad, T);           \
        });                                                                                                    \
        return 0;
    switch (nn) {
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(8)
    }
#undef CASE
    return 2;
}

// out[ntheta + 1] = the fused trace over the whole matrix (kinv: strictly-lower tiles, kdiag: diagonal tiles)
// or, with block = 1, over the rows x cols block at (row0, col0) of the same matrix given as one dense array
// (kdiag ignored), accumulated into out.
int simt_grad_trace(const gogp_op* ops, int nops, int ntheta, int ndim, const double* theta, const double* events,
                    int nev, const double* X, int64_t N, const double* alpha, const double* kinv, const double* kdiag,
                    int block, int64_t row0, int64_t rows, int64_t col0, int64_t cols, double* out) {
    Problem p;
    if (!setup(p, ops, nops, ntheta, ndim, theta, events, nev, X, N)) return 1;
    const DevProgram& prog = p.prog;
    const double* Xt = p.Xt.data();
    const int D = ndim;
    const int64_t Npad = p.Npad;
    const size_t smem = (size_t)2 * D * TILE * sizeof(double) + 16 + (size_t)2 * GT_E * 256 * sizeof(double) +
                        (size_t)(prog.ntheta + 1) * 256 * sizeof(double);
    int ntiles;
    std::vector<double> partial;
    if (!block) {
        const int T = (int)(Npad / TILE);
        ntiles = T * (T + 1) / 2;
        partial.assign((size_t)ntiles * (prog.ntheta + 1), 0.0);
        double* pp = partial.data();
        simt::launch((unsigned)ntiles, 256, smem,
                     [&] { grad_trace_kernel(prog, Xt, Npad, alpha, kinv, Npad, kdiag, N, D, pp, 0, 0, 0); });
    } else {
        const int rt = (int)(rows / TILE), ct = (int)(cols / TILE);
        ntiles = rt * ct;
        partial.assign((size_t)ntiles * (prog.ntheta + 1), 0.0);
        double* pp = partial.data();
        const double* blk = kinv + row0 * Npad + col0;
        simt::launch((unsigned)ntiles, 256, smem, [&] {
            grad_trace_kernel(prog, Xt, Npad, alpha, blk, Npad, nullptr, N, D, pp, ct, row0, col0);
        });
    }
    const double* pp = partial.data();
    const int nslots = prog.ntheta + 1;
    simt::launch((unsigned)nslots, 256, 0, [&] { grad_reduce_kernel(pp, ntiles, nslots, out, block); });
    return 0;
}

// gx[N * D]: the with_obs input gradient
int simt_grad_inputs(const gogp_op* ops, int nops, int ntheta, int ndim, const double* theta, const double* events,
                     int nev, const double* X, int64_t N, const double* alpha, const double* kinv, const double* kdiag,
                     double* gx) {
    Problem p;
    if (!setup(p, ops, nops, ntheta, ndim, theta, events, nev, X, N)) return 1;
    const DevProgram& prog = p.prog;
    const double* Xt = p.Xt.data();
    const int64_t Npad = p.Npad;
    simt::launch((unsigned)N, 256, 0,
                 [&] { grad_inputs_kernel(prog, Xt, Npad, alpha, kinv, Npad, kdiag, N, ndim, gx); });
    return 0;
}
}
