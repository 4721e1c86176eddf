// TEST INFRASTRUCTURE: runs the product's block-cyclic orchestration (gogp_b200/csrc/grid.hpp, unmodified) on the
// HOST: the ranks of a Pr x Pc grid are threads of this process, the collectives are copies through a shared
// mailbox between two barriers, the tile algebra is the naive backend of tests/host_tiles.h driven by the
// product's blocked recursion (csrc/blocked.hpp), and the covariance / trace element functions are the product's
// own (csrc/kexpr.cuh + csrc/program.cc compiled as plain C++).  Checks the index arithmetic of the distribution,
// the panel organisation, the tile masks and the ORDER of the collectives (a mismatch deadlocks or corrupts)
// without a GPU.  Never linked into libgogp_b200.so.
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../gogp_b200/csrc/grid.hpp"
#include "../gogp_b200/csrc/kexpr.cuh"
#include "../gogp_b200/csrc/program.cc"
#include "host_tiles.h"

using namespace gogp;

namespace {

struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    int n, waiting = 0, gen = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const int g = gen;
        if (++waiting == n) {
            waiting = 0;
            ++gen;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};

struct ThreadComm {
    int world;
    Barrier bar;
    std::vector<double*> slot;
    std::vector<long> calls;  // per rank: number of collectives entered (all ranks must agree)
    explicit ThreadComm(int w) : world(w), bar(w), slot(w, nullptr), calls(w, 0) {}
};

struct HostGridBackend {
    ThreadComm& comm;
    int rank;
    const DevProgram* prog;
    double noise_var;
    const double* Xt;  // [D][Npad]
    int64_t Npad, N;
    int D;
    int info = 0;
    long gemm_tiles = 0, gemm_tiles_skipped = 0;
    int64_t comm_bytes = 0;

    double* alloc(int64_t n) {  // poisoned: nothing may depend on memory the orchestration did not write
        double* p = new double[n > 0 ? n : 1];
        for (int64_t i = 0; i < n; ++i) p[i] = std::nan("");
        return p;
    }
    void free(double* p) { delete[] p; }
    void zero(double* p, int64_t n, int) { std::memset(p, 0, sizeof(double) * n); }
    void zero2d(double* p, int64_t ld, int64_t rows, int64_t cols, int) {
        for (int64_t r = 0; r < rows; ++r) std::memset(p + r * ld, 0, sizeof(double) * cols);
    }
    void copy2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, int) {
        for (int64_t r = 0; r < rows; ++r) std::memcpy(dst + r * ldd, src + r * lds, sizeof(double) * cols);
    }
    // queues and events: everything is synchronous here
    int record(int) { return 0; }
    void wait(int, int) {}
    void sync(int) {}
    int tic(int) { return 0; }
    void tic_reset() {}
    double toc(int, int) { return 0.0; }
    void d2h(double* host, const double* dev, int64_t n, int) { std::memcpy(host, dev, sizeof(double) * n); }

    void cov_block(int64_t row0, int64_t rows, int64_t col0, int64_t cols, bool diagonal, double* out, int64_t ld, int) {
        for (int64_t r = 0; r < rows; ++r)
            for (int64_t c = 0; c < cols; ++c) {
                const int64_t gi = row0 + r, gj = col0 + c;
                if (diagonal && (c / 128) > (r / 128)) continue;  // lower tiles only, as the device kernel
                double v;
                if (gi >= N || gj >= N) {
                    v = (diagonal && gi == gj) ? 1.0 : 0.0;
                } else {
                    auto xa = [&](int d) { return Xt[d * Npad + gj]; };
                    auto xb = [&](int d) { return Xt[d * Npad + gi]; };
                    v = 0.0;
                    for (int t = 0; t < prog->nterms; ++t) v += term_value(*prog, t, xa, xb);
                    if (gi == gj) v += noise_var;
                }
                out[r * ld + c] = v;
            }
    }
    void potrf(double* A, int64_t ld, int64_t n, double* winv, int base, int) {
        // the leaf reports base + local index: shift the matrix origin instead of the index (as capi.cu does)
        struct Shifted : gogp_host::HostBackend {
            int shift = 0;
            void potrf_leaf(double* At, int64_t l, double* w, int b) { HostBackend::potrf_leaf(At, l, w, b + shift); }
        } hb;
        hb.shift = base;
        Blocked<Shifted> bl{hb, A, ld, winv, 0, 0};
        bl.potrf(0, n);
        if (hb.info && !info) info = hb.info;
    }
    void trsm(double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl, int64_t n, const double* winv, int) {
        gogp_host::HostBackend hb;
        Blocked<gogp_host::HostBackend> bl{hb, const_cast<double*>(L), ldl, const_cast<double*>(winv), 0, 0};
        bl.trsm(B, ldb, m, 0, n);
    }
    void trtri_t(const double* L, int64_t ld, int64_t n, const double* winv, double* out, int) {
        gogp_host::HostBackend hb;
        Blocked<gogp_host::HostBackend> bl{hb, const_cast<double*>(L), ld, const_cast<double*>(winv), 0, 0};
        bl.trtri_t(out, 0, n);
    }
    void gemm(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t m, int64_t n,
              int64_t k, double alpha, double beta, const BcMask* mk, int flags, int) {
        std::vector<double> tile(128 * 128);
        for (int64_t ti = 0; ti < m / 128; ++ti)
            for (int64_t tj = 0; tj < n / 128; ++tj) {
                if (mk && mk->tb > 0) {
                    const int64_t I = mk->r0 + mk->pr * (ti / mk->tb), J = mk->c0 + mk->pc * (tj / mk->tb);
                    if (J > I || (J == I && tj % mk->tb > ti % mk->tb)) {
                        ++gemm_tiles_skipped;
                        continue;
                    }
                }
                ++gemm_tiles;
                for (int i = 0; i < 128; ++i)
                    for (int j = 0; j < 128; ++j) {
                        const double* a = A + (ti * 128 + i) * lda;
                        const double* b = B + (tj * 128 + j) * ldb;
                        double s = 0.0;
                        for (int64_t kk = (flags & GF_KTRI) ? ti * 128 : 0; kk < k; ++kk) s += a[kk] * b[kk];
                        tile[i * 128 + j] = s;
                    }
                for (int i = 0; i < 128; ++i)
                    for (int j = 0; j < 128; ++j) {
                        double* c = C + (ti * 128 + i) * ldc + tj * 128 + j;
                        *c = alpha * tile[i * 128 + j] + (beta != 0.0 ? beta * *c : 0.0);
                    }
            }
    }
    void sumlogdiag_add(const double* L, int64_t ld, int64_t nvalid, double* accp, int) {
        double s = 0.0;
        for (int64_t i = 0; i < nvalid; ++i) s += std::log(L[i * ld + i]);
        *accp += s;
    }
    void gemv_acc(const double* B, int64_t ld, int64_t rows, int64_t cols, const double* v, double* accp, double a, int) {
        for (int64_t r = 0; r < rows; ++r) {
            double s = 0.0;
            for (int64_t c = 0; c < cols; ++c) s += B[r * ld + c] * v[c];
            accp[r] += a * s;
        }
    }
    void axpy(double* yv, const double* x, double a, int64_t n, int) {
        for (int64_t i = 0; i < n; ++i) yv[i] += a * x[i];
    }
    void dot_add(const double* x, const double* yv, int64_t n, double* accp, int) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += x[i] * yv[i];
        *accp += s;
    }
    void trsv(const double* L, int64_t ld, const double*, double* rhs, double* zv, int64_t n, int) {
        for (int64_t i = 0; i < n; ++i) {
            double s = rhs[i];
            for (int64_t k = 0; k < i; ++k) s -= L[i * ld + k] * zv[k];
            zv[i] = s / L[i * ld + i];
        }
    }
    void trace_block(const double* alpha, const double* kinv, int64_t ld, int64_t row0, int64_t rows, int64_t col0,
                     int64_t cols, double* accp, int) {
        const int nts = prog->ntheta;
        for (int64_t r = 0; r < rows; ++r)
            for (int64_t c = 0; c < cols; ++c) {
                const int64_t gi = row0 + r, gj = col0 + c;
                if (gi >= N || gj >= N || gi < gj) continue;
                const double W = alpha[gi] * alpha[gj] - kinv[r * ld + c];
                const double w = gi == gj ? 0.5 * W : W;
                if (gi == gj) accp[nts] += W;
                auto xa = [&](int d) { return Xt[d * Npad + gj]; };
                auto xb = [&](int d) { return Xt[d * Npad + gi]; };
                for (int t = 0; t < prog->nterms; ++t) {
                    const double P = w * term_value(*prog, t, xa, xb);
                    for (int fi = prog->fbeg[t]; fi < prog->fbeg[t + 1]; ++fi) {
                        const DevFactor& f = prog->f[fi];
                        if (f.p0 < 0) continue;
                        double g0, g1;
                        factor_dlog_theta(f, xa(f.dim), xb(f.dim), g0, g1);
                        accp[f.p0] += P * g0;
                        if (f.p1 >= 0) accp[f.p1] += P * g1;
                    }
                }
            }
    }
    void trace_local(const double* alpha, const double* kinv, int64_t ld, int64_t rows, int64_t cols, const BcMask& mk,
                     double* accp, int q) {
        // tile by tile with the kernel's mapping (cov_kernels.cuh trace_tile, mode 2)
        for (int64_t ti = 0; ti < rows / 128; ++ti)
            for (int64_t tj = 0; tj < cols / 128; ++tj) {
                const int64_t row0 = ((mk.r0 + mk.pr * (ti / mk.tb)) * mk.tb + ti % mk.tb) * 128;
                const int64_t col0 = ((mk.c0 + mk.pc * (tj / mk.tb)) * mk.tb + tj % mk.tb) * 128;
                if (col0 > row0 + 127) continue;
                trace_block(alpha, kinv + ti * 128 * ld + tj * 128, ld, row0, 128, col0, 128, accp, q);
            }
    }
    void range_push(const char*) {}
    void range_pop() {}
    void info_reset(int) { info = 0; }
    void info_to(double* dst, int) { *dst = (double)info; }
    int info_host() { return info; }

    // ---- collectives: a mailbox between barriers -------------------------------------------------
    void bcast(double* p, int64_t n, int root, int) {
        ++comm.calls[rank];
        comm.slot[rank] = p;
        comm.bar.wait();
        if (rank != root) {
            std::memcpy(p, comm.slot[root], sizeof(double) * n);
            comm_bytes += 8 * n;
        }
        comm.bar.wait();
    }
    template <class OP>
    void allreduce(double* p, int64_t n, OP op) {
        ++comm.calls[rank];
        comm.slot[rank] = p;
        comm.bar.wait();
        std::vector<double> tmp(comm.slot[0], comm.slot[0] + n);
        for (int r = 1; r < comm.world; ++r)
            for (int64_t i = 0; i < n; ++i) tmp[i] = op(tmp[i], comm.slot[r][i]);
        comm.bar.wait();
        std::memcpy(p, tmp.data(), sizeof(double) * n);
        comm_bytes += 8 * n;
        comm.bar.wait();
    }
    void allreduce_sum(double* p, int64_t n, int) {
        allreduce(p, n, [](double a, double b) { return a + b; });
    }
    void allreduce_max(double* p, int64_t n, int) {
        allreduce(p, n, [](double a, double b) { return a > b ? a : b; });
    }
};

}  // namespace

// One LML + gradient evaluation of the block-cyclic path on a Pr x Pc grid of host threads.
// Returns 0, 1 (bad descriptor), 2 (not positive definite; *lml holds the pivot), 3 (ranks disagree).
// grad: ntheta_s + ntheta_n; alpha_out: N; stats: [gemm tiles computed, gemm tiles skipped by the mask,
// collectives entered by rank 0, bytes received by rank 0].
extern "C" int cpu_grid_eval(const gogp_op* sops, int n_sops, int nts, const gogp_op* nops, int n_nops, int ntn, int ndim,
                             const double* theta_s, const double* theta_n, const double* X, const double* Y, int64_t N,
                             int64_t NB, int Pr, int Pc, double* lml, double* grad, double* alpha_out, double* stats) {
    Program simil, noise;
    std::string err;
    if (!simil.lower(sops, n_sops, nts, ndim, true, &err)) return 1;
    if (!noise.lower(nops, n_nops, ntn, ndim, false, &err)) return 1;
    DevProgram prog;
    simil.bind(theta_s, &prog);
    std::vector<double> ndlog(ntn > 0 ? ntn : 1, 0.0);
    const double nvar = noise.eval_scalar(theta_n, ndlog.data());
    const int world = Pr * Pc;
    const int64_t nb = (N + NB - 1) / NB, Npad = nb * NB;
    std::vector<double> Xt((size_t)ndim * Npad, 0.0), ypad(Npad, 0.0);
    for (int64_t i = 0; i < N; ++i) {
        for (int d = 0; d < ndim; ++d) Xt[d * Npad + i] = X[i * ndim + d];
        ypad[i] = Y[i];
    }
    ThreadComm comm(world);
    std::vector<double> lmls(world, 0.0), grads((size_t)world * (nts + 1), 0.0);
    std::vector<int> rc(world, 0), piv(world, 0);
    std::vector<std::vector<double>> alphas(world);
    std::vector<long> tiles(world, 0), skipped(world, 0);
    std::vector<int64_t> bytes(world, 0);
    std::vector<std::thread> th;
    for (int rank = 0; rank < world; ++rank)
        th.emplace_back([&, rank] {
            HostGridBackend be{comm, rank, &prog, nvar, Xt.data(), Npad, N, ndim};
            BlockCyclic<HostGridBackend> bc(be);
            bc.init(N, NB, rank, world, Pr, Pc, nts);
            std::memcpy(bc.y, ypad.data(), sizeof(double) * Npad);
            if (!bc.observe(&lmls[rank], &piv[rank])) {
                rc[rank] = 2;
            } else {
                bc.gradient(&grads[(size_t)rank * (nts + 1)]);
                alphas[rank].assign(bc.alpha, bc.alpha + N);
            }
            tiles[rank] = be.gemm_tiles;
            skipped[rank] = be.gemm_tiles_skipped;
            bytes[rank] = be.comm_bytes;
            bc.release();
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < world; ++r)
        if (rc[r] != rc[0] || comm.calls[r] != comm.calls[0]) return 3;
    if (rc[0] == 2) {
        *lml = (double)piv[0];
        return 2;
    }
    for (int r = 1; r < world; ++r) {  // replicated results must be identical on every rank
        if (lmls[r] != lmls[0]) return 3;
        for (int q = 0; q <= nts; ++q)
            if (grads[(size_t)r * (nts + 1) + q] != grads[q]) return 3;
        for (int64_t i = 0; i < N; ++i)
            if (alphas[r][i] != alphas[0][i]) return 3;
    }
    *lml = lmls[0];
    for (int q = 0; q < nts; ++q) grad[q] = grads[q];
    for (int q = 0; q < ntn; ++q) grad[nts + q] = 0.5 * grads[nts] * ndlog[q];  // all shipped noises are input-independent
    for (int64_t i = 0; i < N; ++i) alpha_out[i] = alphas[0][i];
    long tsum = 0, ssum = 0;
    for (int r = 0; r < world; ++r) {
        tsum += tiles[r];
        ssum += skipped[r];
    }
    stats[0] = (double)tsum;
    stats[1] = (double)ssum;
    stats[2] = (double)comm.calls[0];
    stats[3] = (double)bytes[0];
    return 0;
}
