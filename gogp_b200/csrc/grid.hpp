// 2D block-cyclic log marginal likelihood + gradient across the GPUs of one box (SURVEY.md section 8e,
// BASELINE configs[4]: N = 131072), SPMD: one BlockCyclic object per rank (a process under torchrun, or a host
// thread per device inside one Go/C++ process), every rank makes the same calls in the same order.
//
// Reference path: gp.GP.Observe -> absorb -> LML (gp/gp.go:374-413, 89-239, 244-253) and gp.GP.Gradient
// (gp/gp.go:418-499), for a K that no longer fits one GPU.
//
// K is cut into NB x NB blocks; block (I, J) lives on process (I mod Pr, J mod Pc) of a Pr x Pc grid
// (rank = r * Pc + c) in ONE local row-major matrix whose block rows / columns are the owned ones in
// increasing order.  The lower blocks (J <= I) hold K -> L -> K^-1, the strictly-upper blocks (J > I) hold
// V = L^-T, the diagonal blocks of V sit in a side array: N^2 * 8 / world bytes per rank for the whole
// LML + gradient evaluation.
//
//   build    every rank evaluates its own blocks from the replicated inputs (no communication)
//   factor   right-looking Cholesky over block columns: the owner factors (k, k) and broadcasts it; process
//            column k mod Pc solves the panel and broadcasts it (every rank ends up with the whole panel,
//            organised by process row); every rank updates its trailing blocks with ONE masked DMMA GEMM
//            launch; one step of look-ahead (block column k+1 first, on the priority queue)
//   solve    z = L^-1 y by block forward substitution (an NB-vector all-reduce + broadcast per step); LML
//   sweep    V = L^-T and K^-1 = V V^T in ONE right-looking pass.  With V L^T = I,
//              V_ci = -(sum_{k=c}^{i-1} V_ck L_ik^T) L_ii^-T   (c < i),   V_cc = L_cc^-T,
//              K^-1_ij = sum_{k >= i} V_ik V_jk^T              (j <= i).
//            Step k broadcasts "panel k" = block column k of V (rows <= k, final by then) together with
//            block column k of L (rows > k); every rank then adds  -V_ck L_ik^T  to its blocks (c, i),
//            c <= k < i, and  V_ik V_jk^T  to its blocks (i, j), j <= i <= k (overwriting L, which is dead
//            there).  No reduction along process rows and no dependent chain beyond one block column: the
//            look-ahead finishes column k+1 of V (update + solve with L_{k+1,k+1}) on the priority queue
//            while the bulk of step k runs on the side queue.
//   alpha    alpha = V z: block-row GEMVs over the owned blocks + one all-reduce of N doubles
//   trace    the fused gradient trace over the owned lower blocks + one all-reduce of ntheta + 1 doubles
//
// Every O(N^3) flop is the one "NT" GEMM  C = beta C + alpha A B^T.  The class is written against a small
// backend (queues, events, tile algebra, collectives) so that the same orchestration drives the CUDA + NCCL
// backend of grid.cu (the product) and, in tests/cpu_grid_backend.cc only, a host backend whose ranks are
// threads -- the index arithmetic and the collective order are checked on a machine without a GPU.
#pragma once
#include <stdint.h>

#include <cmath>
#include <vector>

namespace gogp {

enum GridQueue : int { GQ_MAIN = 0, GQ_SIDE = 1 };

// Block-cyclic tile mask of a GEMM whose C is a window of the local matrix (kernels.h GemmMask).
struct BcMask {
    int tb, r0, pr, c0, pc;
};

enum GridGemmFlags : int { GF_KTRI = 1 };  // A is upper triangular w.r.t. its own origin: k starts at the row tile

enum GridPhase : int {
    GP_BUILD = 0,
    GP_FACTOR = 1,
    GP_SOLVE = 2,
    GP_SWEEP = 3,  // V = L^-T and K^-1 = V V^T
    GP_ALPHA = 4,
    GP_TRACE = 5,
    GP_NPHASE = 6
};

template <class BE>
struct BlockCyclic {
    static constexpr int64_t kT = 128;
    BE& be;
    int64_t N = 0, NB = 0, Npad = 0, ld = 0, bsz = 0;  // bsz = NB * NB
    int nb = 0, Pr = 1, Pc = 1, rank = 0, world = 1, myr = 0, myc = 0, nr = 0, nc = 0, tb = 0, nts = 0;
    double* M = nullptr;      // local matrix, (nr NB) x (nc NB)
    double* vdiag = nullptr;  // [ndiag][NB][NB]: V_cc of the owned diagonal blocks
    double* wdiag = nullptr;  // [ndiag][NB/128][128][128]: tile inverses of the owned diagonal blocks of L
    double* panel[2] = {nullptr, nullptr};  // a whole panel (nb blocks), by process row; two generations
    double* pcb[2] = {nullptr, nullptr};    // the panel blocks that face my block columns, contiguous
    double* dk = nullptr;                   // staging of a diagonal block for its broadcast: L_kk then its tile inverses
    double *y = nullptr, *z = nullptr, *zc = nullptr, *alpha = nullptr, *acc = nullptr, *scal = nullptr;
    double* gacc = nullptr;  // ntheta + 1 trace sums
    std::vector<int64_t> prow_off;  // first panel slot of process row rr
    std::vector<int> dslot;         // global diagonal block -> index into vdiag / wdiag, -1 when not mine
    int ndiag = 0;
    bool factored = false, have_kinv = false;
    double phase_ms[GP_NPHASE] = {0};
    double comm_ms[GP_NPHASE] = {0};  // device time of the collectives on the priority queue (includes waiting for peers)

    enum { S_SUMLOG = 0, S_QUAD = 1, S_INFO = 2, S_TMP = 4, NSCAL = 8 };

    explicit BlockCyclic(BE& b) : be(b) {}

    // ---- geometry ------------------------------------------------------------------------------
    static int count_below(int I, int first, int stride) {  // #{n >= 0 : first + n * stride < I}
        return I <= first ? 0 : (I - first + stride - 1) / stride;
    }
    int rows_of(int rr) const { return count_below(nb, rr, Pr); }
    int owner(int I, int J) const { return (I % Pr) * Pc + (J % Pc); }
    int grow(int i) const { return myr + i * Pr; }
    int gcol(int j) const { return myc + j * Pc; }
    int lrow(int I) const { return I / Pr; }
    int lcol(int J) const { return J / Pc; }
    double* blk(int i, int j) const { return M + (int64_t)i * NB * ld + (int64_t)j * NB; }
    double* pblk(int g, int I) const { return panel[g] + (prow_off[I % Pr] + I / Pr) * bsz; }
    double* myrows(int g, int i) const { return panel[g] + (prow_off[myr] + i) * bsz; }
    double* cblk(int g, int j) const { return pcb[g] + (int64_t)j * bsz; }
    double* vd(int I) const { return vdiag + (int64_t)dslot[I] * bsz; }
    double* wd(int I) const { return wdiag + (int64_t)dslot[I] * NB * kT; }
    double* lkk() const { return dk; }
    double* wkk() const { return dk + bsz; }
    int64_t dk_len() const { return bsz + NB * kT; }
    BcMask mask_at(int i0, int j0) const { return BcMask{tb, grow(i0), Pr, gcol(j0), Pc}; }

    // Bytes of device memory one rank needs (the caller may check it against the free memory first).
    static int64_t bytes_needed(int64_t N, int64_t NB, int Pr, int Pc) {
        const int64_t nb = (N + NB - 1) / NB;
        const int64_t nr = (nb + Pr - 1) / Pr, nc = (nb + Pc - 1) / Pc;
        const int64_t nd = (nb + (Pr > Pc ? Pr : Pc) - 1) / (Pr > Pc ? Pr : Pc) + 1;
        int64_t d = nr * nc * NB * NB + nd * (NB * NB + NB * 128) + 2 * nb * NB * NB + NB * NB + NB * 128;
        if (Pr * Pc > 1) d += 2 * nc * NB * NB;
        return 8 * (d + 6 * nb * NB);
    }

    bool init(int64_t N_, int64_t NB_, int rank_, int world_, int Pr_, int Pc_, int nts_) {
        N = N_;
        NB = NB_;
        rank = rank_;
        world = world_;
        Pr = Pr_;
        Pc = Pc_;
        nts = nts_;
        myr = rank / Pc;
        myc = rank % Pc;
        nb = (int)((N + NB - 1) / NB);
        Npad = (int64_t)nb * NB;
        bsz = NB * NB;
        tb = (int)(NB / kT);
        nr = count_below(nb, myr, Pr);
        nc = count_below(nb, myc, Pc);
        ld = (int64_t)(nc > 0 ? nc : 1) * NB;
        prow_off.assign(Pr + 1, 0);
        for (int rr = 0; rr < Pr; ++rr) prow_off[rr + 1] = prow_off[rr] + rows_of(rr);
        dslot.assign(nb, -1);
        ndiag = 0;
        for (int I = 0; I < nb; ++I)
            if (I % Pr == myr && I % Pc == myc) dslot[I] = ndiag++;
        M = be.alloc((int64_t)(nr > 0 ? nr : 1) * NB * ld);
        vdiag = be.alloc((int64_t)(ndiag > 0 ? ndiag : 1) * bsz);
        wdiag = be.alloc((int64_t)(ndiag > 0 ? ndiag : 1) * NB * kT);
        for (int g = 0; g < 2; ++g) {
            panel[g] = be.alloc((int64_t)nb * bsz);
            // one rank: panel slot I is block I, the column operand IS the panel
            pcb[g] = world == 1 ? panel[g] : be.alloc((int64_t)(nc > 0 ? nc : 1) * bsz);
        }
        dk = be.alloc(dk_len());
        y = be.alloc(Npad);
        z = be.alloc(Npad);
        zc = be.alloc(ld);
        alpha = be.alloc(Npad);
        acc = be.alloc(NB);
        scal = be.alloc(NSCAL);
        gacc = be.alloc(nts + 1);
        return M && vdiag && wdiag && panel[0] && panel[1] && pcb[0] && pcb[1] && dk && y && z && zc && alpha && acc &&
               scal && gacc;
    }

    void release() {
        be.free(M);
        be.free(vdiag);
        be.free(wdiag);
        for (int g = 0; g < 2; ++g) {
            if (world > 1) be.free(pcb[g]);
            be.free(panel[g]);
        }
        be.free(dk);
        be.free(y);
        be.free(z);
        be.free(zc);
        be.free(alpha);
        be.free(acc);
        be.free(scal);
        be.free(gacc);
        M = nullptr;
    }

    // collectives on the priority queue, timed per phase
    struct CommSpan {
        int e0, e1, phase;
    };
    std::vector<CommSpan> spans;
    int cur_phase = GP_BUILD;
    void comm_begin() { spans.push_back(CommSpan{be.tic(GQ_MAIN), -1, cur_phase}); }
    void comm_end() { spans.back().e1 = be.tic(GQ_MAIN); }

    // ---- build (gp/gp.go:109-156, 220-225): every rank evaluates its own lower blocks -------------
    void build() {
        for (int i = 0; i < nr; ++i)
            for (int j = 0; j < nc && gcol(j) <= grow(i); ++j)
                be.cov_block((int64_t)grow(i) * NB, NB, (int64_t)gcol(j) * NB, NB, grow(i) == gcol(j), blk(i, j), ld,
                             GQ_MAIN);
        factored = false;
        have_kinv = false;
    }

    // ---- factor (gp/gp.go:228) ---------------------------------------------------------------------
    void factor_diag(int k) {
        if (rank != owner(k, k)) return;
        double* d = blk(lrow(k), lcol(k));
        be.potrf(d, ld, NB, wd(k), (int)(k * NB), GQ_MAIN);
        const int64_t nvalid = N - (int64_t)k * NB < NB ? N - (int64_t)k * NB : NB;
        be.sumlogdiag_add(d, ld, nvalid, scal + S_SUMLOG, GQ_MAIN);
    }
    // owner of (k, k): L_kk and its tile inverses into the broadcast staging
    void stage_diag(int k) {
        if (rank != owner(k, k)) return;
        be.copy2d(lkk(), NB, blk(lrow(k), lcol(k)), ld, NB, NB, GQ_MAIN);
        be.copy2d(wkk(), NB * kT, wd(k), NB * kT, 1, NB * kT, GQ_MAIN);
    }
    // the panel blocks that face my block columns [j0, j1), in increasing J
    void gather(int g, int j0, int j1) {
        if (world == 1) return;
        for (int j = j0; j < j1; ++j) be.copy2d(cblk(g, j), NB, pblk(g, gcol(j)), NB, NB, NB, GQ_MAIN);
    }

    // Queues.  The priority queue carries the chain of one block column (diagonal factor, panel solve, collectives,
    // look-ahead update of block column k+1); the side queue carries the bulk of every trailing update, split so
    // that the NEXT look-ahead column is updated first (event e1) and the rest after it (event done[g]).  The
    // chain of step k+1 therefore overlaps the bulk of step k and waits only for e1; the two panel generations
    // are recycled behind done[g].
    void factor() {
        cur_phase = GP_FACTOR;
        be.zero(scal, NSCAL, GQ_MAIN);
        be.info_reset(GQ_MAIN);
        int done[2] = {-1, -1}, e1_prev = -1;
        factor_diag(0);
        for (int k = 0; k + 1 < nb; ++k) {
            const int g = k & 1, kc = k % Pc;
            stage_diag(k);
            comm_begin();
            be.bcast(dk, dk_len(), owner(k, k), GQ_MAIN);
            comm_end();
            // panel solve A_Ik <- A_Ik L_kk^-T on process column kc: every source solves its own rows BEFORE any
            // broadcast is queued, so the Pr solves run concurrently
            const int ir0 = count_below(k + 1, myr, Pr), jc0 = count_below(k + 1, myc, Pc);
            const int mrows = nr - ir0, ncols = nc - jc0;
            if (kc == myc && mrows > 0) be.trsm(blk(ir0, lcol(k)), ld, (int64_t)mrows * NB, lkk(), NB, NB, wkk(), GQ_MAIN);
            if (done[g] >= 0) be.wait(GQ_MAIN, done[g]);  // the bulk of step k-2 read this panel generation
            if (kc == myc && mrows > 0)
                be.copy2d(myrows(g, ir0), NB, blk(ir0, lcol(k)), ld, (int64_t)mrows * NB, NB, GQ_MAIN);
            comm_begin();
            for (int rr = 0; rr < Pr; ++rr) {
                const int first = count_below(k + 1, rr, Pr), cnt = rows_of(rr) - first;
                if (cnt > 0) be.bcast(panel[g] + (prow_off[rr] + first) * bsz, (int64_t)cnt * bsz, rr * Pc + kc, GQ_MAIN);
            }
            comm_end();
            gather(g, jc0, nc);
            const int fork = be.record(GQ_MAIN);
            // step k-1's bulk updated block column k+1 first: that is all the look-ahead depends on
            if (e1_prev >= 0) be.wait(GQ_MAIN, e1_prev);
            e1_prev = -1;
            if (mrows > 0 && ncols > 0) {
                const double* A = myrows(g, ir0);
                int first = 0;
                if (gcol(jc0) == k + 1) {
                    // look-ahead: block column k+1 on the priority queue -- the next step's diagonal
                    // factorisation, panel solve and broadcasts depend on nothing else
                    const BcMask m = mask_at(ir0, jc0);
                    be.gemm(blk(ir0, jc0), ld, A, NB, cblk(g, jc0), NB, (int64_t)mrows * NB, NB, NB, -1.0, 1.0, &m,
                            0, GQ_MAIN);
                    first = 1;
                }
                if (ncols > first) {
                    be.wait(GQ_SIDE, fork);
                    const int j1 = jc0 + first, w1 = gcol(j1) == k + 2 ? 1 : 0;
                    if (w1) {
                        const BcMask m = mask_at(ir0, j1);
                        be.gemm(blk(ir0, j1), ld, A, NB, cblk(g, j1), NB, (int64_t)mrows * NB, NB, NB, -1.0, 1.0, &m,
                                0, GQ_SIDE);
                    }
                    e1_prev = be.record(GQ_SIDE);
                    if (ncols - first - w1 > 0) {
                        const BcMask m = mask_at(ir0, j1 + w1);
                        be.gemm(blk(ir0, j1 + w1), ld, A, NB, cblk(g, j1 + w1), NB, (int64_t)mrows * NB,
                                (int64_t)(ncols - first - w1) * NB, NB, -1.0, 1.0, &m, 0, GQ_SIDE);
                    }
                    done[g] = be.record(GQ_SIDE);
                }
            }
            factor_diag(k + 1);
        }
        for (int g = 0; g < 2; ++g)
            if (done[g] >= 0) be.wait(GQ_MAIN, done[g]);
        factored = true;
        have_kinv = false;
    }

    // ---- z = L^-1 y, log det, y^T K^-1 y (gp/gp.go:232-236, 244-253) -------------------------------
    // Returns false when the covariance was not positive definite (bad_pivot holds the 1-based index or -1
    // when another rank saw it).
    bool solve_lml(double* lml, int* bad_pivot) {
        cur_phase = GP_SOLVE;
        for (int k = 0; k < nb; ++k) {
            be.zero(acc, NB, GQ_MAIN);
            if (k % Pr == myr) {
                const int ncl = count_below(k, myc, Pc);  // my block columns J < k
                if (ncl > 0) be.gemv_acc(blk(lrow(k), 0), ld, NB, (int64_t)ncl * NB, zc, acc, -1.0, GQ_MAIN);
            }
            comm_begin();
            be.allreduce_sum(acc, NB, GQ_MAIN);  // ranks outside the process row add zeros
            comm_end();
            const int ok = owner(k, k);
            double* zk = z + (int64_t)k * NB;
            if (rank == ok) {
                be.axpy(acc, y + (int64_t)k * NB, 1.0, NB, GQ_MAIN);  // rhs = y_k - sum_J L_kJ z_J
                be.trsv(blk(lrow(k), lcol(k)), ld, wd(k), acc, zk, NB, GQ_MAIN);
            }
            comm_begin();
            be.bcast(zk, NB, ok, GQ_MAIN);
            comm_end();
            be.dot_add(zk, zk, NB, scal + S_QUAD, GQ_MAIN);
            if (k % Pc == myc) be.copy2d(zc + (int64_t)lcol(k) * NB, NB, zk, NB, 1, NB, GQ_MAIN);
        }
        be.info_to(scal + S_INFO, GQ_MAIN);
        comm_begin();
        be.allreduce_sum(scal + S_SUMLOG, 1, GQ_MAIN);
        be.allreduce_max(scal + S_INFO, 1, GQ_MAIN);
        comm_end();
        double hs[NSCAL];
        be.d2h(hs, scal, NSCAL, GQ_MAIN);  // synchronises the queue
        if (hs[S_INFO] != 0.0) {
            const int mine = be.info_host();
            *bad_pivot = mine ? mine : (int)hs[S_INFO];
            return false;
        }
        // gp/gp.go:244-253: -N/2 log(2 pi) - 1/2 log det K - 1/2 y^T alpha, log det = 2 sum log L_ii, y^T alpha = z^T z
        *lml = -0.5 * (double)N * std::log(2 * M_PI) - 0.5 * (2.0 * hs[S_SUMLOG]) - 0.5 * hs[S_QUAD];
        return true;
    }

    // ---- V = L^-T and K^-1 = V V^T in one right-looking pass (replaces gp/gp.go:454,480) ---------------
    void sweep() {
        cur_phase = GP_SWEEP;
        // the strictly-upper blocks are accumulated into (beta = 1); V_cc lives in vdiag (only its upper
        // tiles are written by the tile algebra)
        for (int i = 0; i < nr; ++i) {
            const int j0 = count_below(grow(i) + 1, myc, Pc);
            if (nc > j0) be.zero2d(blk(i, j0), ld, NB, (int64_t)(nc - j0) * NB, GQ_MAIN);
        }
        be.zero(vdiag, (int64_t)(ndiag > 0 ? ndiag : 1) * bsz, GQ_MAIN);
        stage_diag(0);
        if (rank == owner(0, 0)) be.trtri_t(lkk(), NB, NB, wkk(), vd(0), GQ_MAIN);
        int done[2] = {-1, -1}, e1_prev = -1;  // as in factor()
        for (int k = 0; k < nb; ++k) {
            const int g = k & 1, kc = k % Pc;
            if (k + 1 < nb) {
                // L_{k+1,k+1}: the look-ahead's solve needs it on process column (k+1) mod Pc; its owner also
                // forms V_{k+1,k+1} now, off the next step's critical path
                stage_diag(k + 1);
                comm_begin();
                be.bcast(dk, dk_len(), owner(k + 1, k + 1), GQ_MAIN);
                comm_end();
                if (rank == owner(k + 1, k + 1)) be.trtri_t(lkk(), NB, NB, wkk(), vd(k + 1), GQ_MAIN);
            }
            if (done[g] >= 0) be.wait(GQ_MAIN, done[g]);  // the bulk of step k-2 read this panel generation
            // panel k = [V_ck (c <= k) ; L_ik (i > k)], from process column kc
            if (kc == myc && nr > 0) {
                be.copy2d(myrows(g, 0), NB, blk(0, lcol(k)), ld, (int64_t)nr * NB, NB, GQ_MAIN);
                if (k % Pr == myr) be.copy2d(myrows(g, lrow(k)), NB, vd(k), NB, NB, NB, GQ_MAIN);
            }
            comm_begin();
            for (int rr = 0; rr < Pr; ++rr)
                if (rows_of(rr) > 0)
                    be.bcast(panel[g] + prow_off[rr] * bsz, (int64_t)rows_of(rr) * bsz, rr * Pc + kc, GQ_MAIN);
            comm_end();
            gather(g, 0, nc);
            const int fork = be.record(GQ_MAIN);
            if (e1_prev >= 0) be.wait(GQ_MAIN, e1_prev);  // step k-1's bulk updated block column k+1 first
            e1_prev = -1;
            const int nrk = count_below(k + 1, myr, Pr);  // my block rows <= k
            const int nck = count_below(k + 1, myc, Pc);  // my block columns <= k
            const double* A = myrows(g, 0);
            int first = 0;
            if (nck < nc && gcol(nck) == k + 1) {
                // look-ahead: finish block column k+1 of V (last update, then the solve with L_{k+1,k+1})
                if (nrk > 0) {
                    be.gemm(blk(0, nck), ld, A, NB, cblk(g, nck), NB, (int64_t)nrk * NB, NB, NB, -1.0, 1.0, nullptr,
                            0, GQ_MAIN);
                    be.trsm(blk(0, nck), ld, (int64_t)nrk * NB, lkk(), NB, NB, wkk(), GQ_MAIN);
                }
                first = 1;
            }
            if (nrk == 0) continue;
            be.wait(GQ_SIDE, fork);
            // V_ci -= V_ck L_ik^T for my blocks c <= k < i beyond the look-ahead column; the next look-ahead
            // column (k+2) first
            const int j1 = nck + first, w1 = (j1 < nc && gcol(j1) == k + 2) ? 1 : 0;
            if (w1)
                be.gemm(blk(0, j1), ld, A, NB, cblk(g, j1), NB, (int64_t)nrk * NB, NB, NB, -1.0, 1.0, nullptr, 0, GQ_SIDE);
            e1_prev = be.record(GQ_SIDE);
            const bool own_k = k % Pr == myr;  // my last block row <= k is row k itself: its A operand is V_kk,
                                               // upper triangular -> triangular k range, half the flops
            if (nc - j1 - w1 > 0) {
                const int full = own_k ? nrk - 1 : nrk;
                if (full > 0)
                    be.gemm(blk(0, j1 + w1), ld, A, NB, cblk(g, j1 + w1), NB, (int64_t)full * NB,
                            (int64_t)(nc - j1 - w1) * NB, NB, -1.0, 1.0, nullptr, 0, GQ_SIDE);
                if (own_k)
                    be.gemm(blk(nrk - 1, j1 + w1), ld, myrows(g, nrk - 1), NB, cblk(g, j1 + w1), NB, NB,
                            (int64_t)(nc - j1 - w1) * NB, NB, -1.0, 1.0, nullptr, GF_KTRI, GQ_SIDE);
            }
            // K^-1_ij += V_ik V_jk^T for my blocks j <= i <= k; block row k is touched for the first time
            if (nck > 0) {
                int nrows = nrk;
                if (own_k) {
                    const BcMask m = mask_at(nrk - 1, 0);
                    be.gemm(blk(nrk - 1, 0), ld, myrows(g, nrk - 1), NB, cblk(g, 0), NB, NB, (int64_t)nck * NB, NB, 1.0,
                            0.0, &m, GF_KTRI, GQ_SIDE);
                    nrows = nrk - 1;
                }
                if (nrows > 0) {
                    const BcMask m = mask_at(0, 0);
                    be.gemm(blk(0, 0), ld, A, NB, cblk(g, 0), NB, (int64_t)nrows * NB, (int64_t)nck * NB, NB, 1.0, 1.0,
                            &m, 0, GQ_SIDE);
                }
            }
            done[g] = be.record(GQ_SIDE);
        }
        for (int g = 0; g < 2; ++g)
            if (done[g] >= 0) be.wait(GQ_MAIN, done[g]);
        factored = false;  // L has been overwritten
        have_kinv = true;
    }

    // ---- alpha = K^-1 y = V z (gp/gp.go:232-233) -----------------------------------------------------
    void solve_alpha() {
        cur_phase = GP_ALPHA;
        be.zero(alpha, Npad, GQ_MAIN);
        for (int i = 0; i < nr; ++i) {
            const int I = grow(i), j0 = count_below(I + 1, myc, Pc);
            double* ai = alpha + (int64_t)I * NB;
            if (nc > j0) be.gemv_acc(blk(i, j0), ld, NB, (int64_t)(nc - j0) * NB, zc + (int64_t)j0 * NB, ai, 1.0, GQ_MAIN);
            if (dslot[I] >= 0) be.gemv_acc(vd(I), NB, NB, NB, z + (int64_t)I * NB, ai, 1.0, GQ_MAIN);
        }
        comm_begin();
        be.allreduce_sum(alpha, Npad, GQ_MAIN);
        comm_end();
    }

    // ---- gradient trace (gp/gp.go:434-486): out = [0.5 tr(W dK/dlog theta_s) ..., tr(W)] -----------------
    void trace(double* out) {
        cur_phase = GP_TRACE;
        be.zero(gacc, nts + 1, GQ_MAIN);
        // one launch over the local matrix: the kernel maps a local tile to its global position and skips the blocks
        // above the global diagonal (they hold V)
        if (nr > 0 && nc > 0) be.trace_local(alpha, M, ld, (int64_t)nr * NB, (int64_t)nc * NB, mask_at(0, 0), gacc, GQ_MAIN);
        comm_begin();
        be.allreduce_sum(gacc, nts + 1, GQ_MAIN);
        comm_end();
        be.d2h(out, gacc, nts + 1, GQ_MAIN);
    }

    // ---- one evaluation, with per-phase device times ---------------------------------------------------
    bool observe(double* lml, int* bad_pivot) {
        spans.clear();
        be.tic_reset();
        for (double& v : phase_ms) v = 0.0;
        for (double& v : comm_ms) v = 0.0;
        const int t0 = be.tic(GQ_MAIN);
        be.range_push("gogp_grid:build");
        build();
        be.range_pop();
        const int t1 = be.tic(GQ_MAIN);
        be.range_push("gogp_grid:factor");
        factor();
        be.range_pop();
        const int t2 = be.tic(GQ_MAIN);
        be.range_push("gogp_grid:solve");
        const bool ok = solve_lml(lml, bad_pivot);
        be.range_pop();
        const int t3 = be.tic(GQ_MAIN);
        be.sync(GQ_MAIN);
        phase_ms[GP_BUILD] = be.toc(t0, t1);
        phase_ms[GP_FACTOR] = be.toc(t1, t2);
        phase_ms[GP_SOLVE] = be.toc(t2, t3);
        collect_comm();
        if (!ok) factored = false;
        return ok;
    }
    void gradient(double* out /* nts + 1 */) {
        spans.clear();
        be.tic_reset();
        const int t0 = be.tic(GQ_MAIN);
        be.range_push("gogp_grid:sweep");
        if (!have_kinv) sweep();
        be.range_pop();
        const int t1 = be.tic(GQ_MAIN);
        be.range_push("gogp_grid:alpha+trace");
        solve_alpha();
        const int t2 = be.tic(GQ_MAIN);
        trace(out);
        be.range_pop();
        const int t3 = be.tic(GQ_MAIN);
        be.sync(GQ_MAIN);
        phase_ms[GP_SWEEP] = be.toc(t0, t1);
        phase_ms[GP_ALPHA] = be.toc(t1, t2);
        phase_ms[GP_TRACE] = be.toc(t2, t3);
        comm_ms[GP_SWEEP] = comm_ms[GP_ALPHA] = comm_ms[GP_TRACE] = 0.0;
        collect_comm();
    }
    void collect_comm() {
        for (const CommSpan& s : spans)
            if (s.e1 >= 0) comm_ms[s.phase] += be.toc(s.e0, s.e1);
    }
};

}  // namespace gogp
