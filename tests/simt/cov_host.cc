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

// The specialised trace kernel's instantiation for a program (mirrors trace_fast_shape of csrc/cov.cu)
static bool trace_shape(const DevProgram& prog, int* nn_out, int* maxr_out) {
    static const int inst[] = {8, 4, 3, 2, 1, 0};
    for (int nn : inst) {
        if (nn > 0 && (prog.nterms != 1 || prog.nnorm[0] < nn)) continue;
        int worst = 0;
        for (int t = 0; t < prog.nterms; ++t) {
            int rest = 0;
            for (int fi = prog.fbeg[t] + nn; fi < prog.fbeg[t + 1]; ++fi)
                if (prog.f[fi].kind != F_PARAM) ++rest;
            if (rest > worst) worst = rest;
        }
        if (worst > kTraceMaxRest) continue;
        *nn_out = nn;
        *maxr_out = worst == 0 ? 0 : (worst == 1 ? 1 : kTraceMaxRest);
        return true;
    }
    return false;
}

// out[ntheta + 1] = the fused trace.  block = 0: over the whole matrix (kinv: strictly-lower tiles, kdiag: diagonal
// tiles); block = 1: over the rows x cols block at (row0, col0) of the same matrix given as one dense array (kdiag
// ignored), accumulated into out; block = 2: kinv is ONE RANK's local matrix (rows x cols, ld = cols) of a
// pr x pc block-cyclic distribution with tb tiles per block whose first block row / column is global block
// row0 / col0 (bc = {tb, pr, pc}), accumulated into out.  fast: 0 the interpreted kernel, 1 the specialised one
// (returns 2 when the program has no fast shape).
int simt_grad_trace(const gogp_op* ops, int nops, int ntheta, int ndim, const double* theta, const double* events,
                    int nev, const double* X, int64_t N, const double* alpha, const double* kinv, const double* kdiag,
                    int block, int64_t row0, int64_t rows, int64_t col0, int64_t cols, const int* bc, int fast,
                    double* out) {
    Problem p;
    if (!setup(p, ops, nops, ntheta, ndim, theta, events, nev, X, N)) return 1;
    const DevProgram& prog = p.prog;
    const double* Xt = p.Xt.data();
    const int D = ndim;
    const int64_t Npad = p.Npad;
    const size_t smem = (size_t)2 * D * TILE * sizeof(double) + 16 + (size_t)2 * GT_E * 256 * sizeof(double) +
                        (size_t)(prog.ntheta + 1) * 256 * sizeof(double);
    TraceMap map{};
    int ntiles;
    const double* base = kinv;
    int64_t ld = Npad;
    if (block == 0) {
        const int T = (int)(Npad / TILE);
        ntiles = T * (T + 1) / 2;
        map.mode = 0;
    } else if (block == 1) {
        const int rt = (int)(rows / TILE), ct = (int)(cols / TILE);
        ntiles = rt * ct;
        map.mode = 1;
        map.ctiles = ct;
        map.grow0 = row0;
        map.gcol0 = col0;
        base = kinv + row0 * Npad + col0;
        kdiag = nullptr;
    } else {
        const int rt = (int)(rows / TILE), ct = (int)(cols / TILE);
        ntiles = rt * ct;
        map.mode = 2;
        map.ctiles = ct;
        map.tb = bc[0];
        map.r0 = (int)row0;
        map.pr = bc[1];
        map.c0 = (int)col0;
        map.pc = bc[2];
        ld = cols;
        kdiag = nullptr;
    }
    std::vector<double> partial((size_t)ntiles * (prog.ntheta + 1), 0.0);
    double* pp = partial.data();
    if (!fast) {
        simt::launch((unsigned)ntiles, 256, smem,
                     [&] { grad_trace_kernel(prog, Xt, Npad, alpha, base, ld, kdiag, N, D, pp, map); });
    } else {
        int nn = -1, maxr = 0;
        if (!trace_shape(prog, &nn, &maxr)) return 2;
        bool done = false;
#define CASE(NNV, MR)                                                                                       \
    if (!done && nn == NNV && maxr == MR) {                                                                 \
        simt::launch((unsigned)ntiles, 256, smem, [&] {                                                     \
            grad_trace_fast_kernel<NNV, MR, 1>(prog, Xt, Npad, alpha, base, ld, kdiag, N, D, pp, map);      \
        });                                                                                                 \
        done = true;                                                                                        \
    }
        CASE(0, 0) CASE(0, 1) CASE(0, 4) CASE(1, 0) CASE(1, 1) CASE(1, 4) CASE(2, 0) CASE(2, 1) CASE(2, 4)
        CASE(3, 0) CASE(3, 1) CASE(3, 4) CASE(4, 0) CASE(4, 1) CASE(4, 4) CASE(8, 0) CASE(8, 1) CASE(8, 4)
#undef CASE
        if (!done) return 2;
    }
    const int nslots = prog.ntheta + 1;
    simt::launch((unsigned)nslots, 256, 0, [&] { grad_reduce_kernel(pp, ntiles, nslots, out, block != 0); });
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
