// Recursive blocked dense algebra over 128x128 tiles, written against a small
// backend of tile primitives so the same recursion drives the CUDA kernels (the
// product) and, in tests/cpu_blocked_test.cc only, a naive host backend that
// checks the index arithmetic without a GPU.
//
// Storage: row-major, lower triangle, leading dimension ld; everything is a
// multiple of 128.  All O(n^3) work is expressed as  C = beta C + alpha A B^T
// ("NT" GEMM, k contiguous in both operands):
//   potrf(A)       A = L L^T             n^3/3   (reference: gp.L.Factorize, gp/gp.go:228)
//   trsm(B, L)     X L^T = B             m n^2   (Produce's L.SolveTo, gp/gp.go:337-340)
//   trtri_t(L)     U = L^-T (upper)      n^3/3   \  together K^-1, replacing the per-parameter
//   lauum(U)       K^-1 = U U^T (lower)  n^3/3   /  L.SolveTo of gp/gp.go:454,480
// The 128x128 diagonal blocks are factored by a leaf kernel that also returns
// their inverses, so every triangular solve with a diagonal block is a GEMM.
#pragma once
#include <stdint.h>

#include <vector>

namespace gogp {

constexpr int64_t kTile = 128;

enum : int { BL_FULL = 0, BL_LOWER = 1, BL_KTRI = 2, BL_DIAG_OUT = 4, BL_INPLACE = 8 };  // == GemmMode

template <class BE>
struct Blocked {
    BE& be;
    double* A;     // Npad x Npad: K -> L
    int64_t ld;
    double* winv;  // [T][128][128] inverses of L's diagonal tiles

    static int64_t split(int64_t n) { return (n / kTile / 2) * kTile; }
    double* at(double* M, int64_t r, int64_t c) const { return M + r * ld + c; }

    int64_t rl_max = 0;    // diagonal blocks up to this size are factored right-looking (0: never)
    int64_t cols_max = 0;  // diagonal blocks up to this size are inverted column-wise (0: never)

    // Right-looking factorisation of a small diagonal block with 128-wide panels:
    // 3 launches per panel and every GEMM has K = 128, so the chain of dependent
    // kernels is 3n/128 short ones instead of the recursion's ~5n/128 longer ones.
    // Used below rl_max, where launches are latency-bound and flops are negligible.
    void potrf_rl(int64_t o, int64_t n) {
        for (int64_t p = 0; p < n; p += kTile) {
            double* wp = winv + ((o + p) / kTile) * kTile * kTile;
            be.potrf_leaf(at(A, o + p, o + p), ld, wp, (int)(o + p));
            const int64_t m = n - p - kTile;
            if (m <= 0) break;
            double* panel = at(A, o + p + kTile, o + p);
            be.gemm(panel, ld, panel, ld, wp, kTile, m, kTile, kTile, 1.0, 0.0, BL_INPLACE, nullptr);
            be.gemm(at(A, o + p + kTile, o + p + kTile), ld, panel, ld, panel, ld, m, m, kTile, -1.0, 1.0, BL_LOWER,
                    nullptr);
        }
    }

    // A[o:o+n, o:o+n] = L L^T
    void potrf(int64_t o, int64_t n) {
        if (n == kTile) {
            be.potrf_leaf(at(A, o, o), ld, winv + (o / kTile) * kTile * kTile, (int)o);
            return;
        }
        if (n <= rl_max) {
            potrf_rl(o, n);
            return;
        }
        const int64_t n1 = split(n), n2 = n - n1;
        potrf(o, n1);
        trsm(at(A, o + n1, o), ld, n2, o, n1);
        be.gemm(at(A, o + n1, o + n1), ld, at(A, o + n1, o), ld, at(A, o + n1, o), ld, n2, n2, n1, -1.0, 1.0, BL_LOWER,
                nullptr);
        potrf(o + n1, n2);
    }

    // Right-looking factorisation over block columns with one step of look-ahead, nested:
    // level 0 cuts the matrix into la_nb[0]-wide block columns (2048), level 1 cuts each of
    // their diagonal blocks into la_nb[1]-wide ones, ... down to 128 = one leaf each; a level
    // whose width is 0 or too wide for the block at hand (n < 3 nb) is skipped.  The recursion above
    // leaves the GPU nearly idle while a diagonal block is factored (a chain of one-CTA leaf
    // kernels and K = 128 updates).  Here, per level, with P = panel solve, U1d / U1b = update of
    // the next diagonal block / of the rows below it, U2 = rest of the trailing update, D = the
    // next diagonal block's factorisation (one level down):
    //   current stream:  P_k  [wait U2_k-1]  U1d_k  D_k+1  [wait U1b_k]  P_k+1 ...
    //   bulk stream   :       [after P_k]    U1b_k  U2_k                 U1b_k+1  U2_k+1 ...
    // U1d/U1b write block column k+1, U2 the columns right of it: disjoint, all read P_k only.
    // Same flops as the recursion; the la_* hooks are no-ops for a backend without streams
    // (then this is a plain right-looking sweep).  Level l's bulk stream must outrank level
    // l-1's, and the current stream both.
    static constexpr int kLaLevels = 3;
    int64_t la_nb[kLaLevels] = {0, 0, 0};

    void potrf_la(int64_t o0, int64_t n, int lvl) {
        const int64_t nb = lvl < kLaLevels ? la_nb[lvl] : 0;
        if (nb < kTile || nb % kTile || n < 3 * nb) {
            if (lvl + 1 < kLaLevels)
                potrf_la(o0, n, lvl + 1);
            else
                potrf(o0, n);
            return;
        }
        const int64_t end = o0 + n;
        potrf_la(o0, nb, lvl + 1);
        for (int64_t o = o0; o + nb < end; o += nb) {
            const int64_t r0 = o + nb, m = end - r0, w1 = m < nb ? m : nb, m2 = m - w1;
            double* P = at(A, r0, o);
            trsm(P, ld, m, o, nb);
            be.la_fork(lvl);       // the bulk stream may start once the panel is solved
            be.la_wait_bulk(lvl);  // step k-1's bulk update wrote the block U1d touches
            be.gemm(at(A, r0, r0), ld, P, ld, P, ld, w1, w1, nb, -1.0, 1.0, BL_LOWER, nullptr);
            if (m2 > 0) {
                double* P2 = at(A, r0 + w1, o);
                be.la_bulk_begin(lvl);
                be.gemm(at(A, r0 + w1, r0), ld, P2, ld, P, ld, m2, w1, nb, -1.0, 1.0, BL_FULL, nullptr);
                be.la_mark_below(lvl);
                be.gemm(at(A, r0 + w1, r0 + w1), ld, P2, ld, P2, ld, m2, m2, nb, -1.0, 1.0, BL_LOWER, nullptr);
                be.la_bulk_end(lvl);
            }
            potrf_la(r0, w1, lvl + 1);
            be.la_wait_below(lvl);  // the next panel solve reads the rows U1b wrote
        }
        be.la_wait_bulk(lvl);
    }

    // B (m x n, ldb) <- B L^-T with L = A[o:o+n, o:o+n]
    void trsm(double* B, int64_t ldb, int64_t m, int64_t o, int64_t n) {
        if (n == kTile) {
            // X = B Winv^T, in place: one CTA owns a full 128-row block of B (n == BN)
            be.gemm(B, ldb, B, ldb, winv + (o / kTile) * kTile * kTile, kTile, m, kTile, kTile, 1.0, 0.0, BL_INPLACE,
                    nullptr);
            return;
        }
        const int64_t n1 = split(n), n2 = n - n1;
        trsm(B, ldb, m, o, n1);
        be.gemm(B + n1, ldb, B, ldb, at(A, o + n1, o), ld, m, n2, n1, -1.0, 1.0, BL_FULL, nullptr);
        trsm(B + n1, ldb, m, o + n1, n2);
    }

    // U[o:o+n, o:o+n] (upper, in Bm with the same ld) = L[o:o+n, o:o+n]^-T
    // Small-block variant of trtri_t, one block column of U at a time:
    //   U[J,J] = Winv_J^T,   U[I<J, J] = -(sum_{K=I}^{J-1} U[I,K] L[J,K]^T) Winv_J^T
    // (3 launches per column instead of the recursion's nested solve chains).
    void trtri_cols(double* Bm, int64_t o, int64_t n) {
        for (int64_t c = 0; c < n; c += kTile) {
            double* wc = winv + ((o + c) / kTile) * kTile * kTile;
            be.trtri_leaf(wc, at(Bm, o + c, o + c), ld);
            if (c == 0) continue;
            double* col = at(Bm, o, o + c);
            be.gemm(col, ld, at(Bm, o, o), ld, at(A, o + c, o), ld, c, kTile, c, -1.0, 0.0, BL_KTRI, nullptr);
            be.gemm(col, ld, col, ld, wc, kTile, c, kTile, kTile, 1.0, 0.0, BL_INPLACE, nullptr);
        }
    }

    void trtri_t(double* Bm, int64_t o, int64_t n) {
        if (n == kTile) {
            be.trtri_leaf(winv + (o / kTile) * kTile * kTile, at(Bm, o, o), ld);
            return;
        }
        if (n <= cols_max) {
            trtri_cols(Bm, o, n);
            return;
        }
        const int64_t n1 = split(n), n2 = n - n1;
        trtri_t(Bm, o, n1);
        trtri_t(Bm, o + n1, n2);
        // U12 = -U11 L21^T (U11 upper triangular: k >= row), then U12 <- U12 L22^-T
        be.gemm(at(Bm, o, o + n1), ld, at(Bm, o, o), ld, at(A, o + n1, o), ld, n1, n2, n1, -1.0, 0.0, BL_KTRI, nullptr);
        trsm(at(Bm, o, o + n1), ld, n1, o + n1, n2);
    }

    // The same inverse, traversed level by level instead of depth first: the diagonal blocks
    // of size <= par_block are mutually independent (each a long chain of small, latency-bound
    // launches), and so are the nodes of one recursion level, so the backend may spread them
    // over concurrent streams (par_begin / par_use / par_end).  Same arithmetic as trtri_t.
    struct Node {
        int64_t o, n;
    };
    void collect(int64_t o, int64_t n, int depth, int64_t par_block, std::vector<Node>& leaves,
                 std::vector<std::vector<Node>>& levels) {
        if (n <= par_block || n == kTile) {
            leaves.push_back({o, n});
            return;
        }
        const int64_t n1 = split(n);
        collect(o, n1, depth + 1, par_block, leaves, levels);
        collect(o + n1, n - n1, depth + 1, par_block, leaves, levels);
        if ((int)levels.size() <= depth) levels.resize(depth + 1);
        levels[depth].push_back({o, n});
    }
    void trtri_t_levels(double* Bm, int64_t n, int64_t par_block) {
        std::vector<Node> leaves;
        std::vector<std::vector<Node>> levels;
        collect(0, n, 0, par_block, leaves, levels);
        be.par_begin();
        for (size_t i = 0; i < leaves.size(); ++i) {
            be.par_use((int)i);
            trtri_t(Bm, leaves[i].o, leaves[i].n);
        }
        be.par_end();
        for (int d = (int)levels.size() - 1; d >= 0; --d) {
            const bool par = levels[d].size() > 1;
            if (par) be.par_begin();
            for (size_t i = 0; i < levels[d].size(); ++i) {
                const int64_t o = levels[d][i].o, nn = levels[d][i].n, n1 = split(nn), n2 = nn - n1;
                if (par) be.par_use((int)i);
                be.set_scratch_row(o);
                be.gemm(at(Bm, o, o + n1), ld, at(Bm, o, o), ld, at(A, o + n1, o), ld, n1, n2, n1, -1.0, 0.0, BL_KTRI,
                        nullptr);
                trsm(at(Bm, o, o + n1), ld, n1, o + n1, n2);
            }
            if (par) be.par_end();
            be.set_scratch_row(0);
        }
    }

    // K^-1 = U U^T: strictly-lower tiles into Bm's lower triangle, diagonal tiles into dg
    void lauum(double* Bm, double* dg, int64_t n) {
        be.gemm(Bm, ld, Bm, ld, Bm, ld, n, n, n, 1.0, 0.0, BL_LOWER | BL_KTRI | BL_DIAG_OUT, dg);
    }
};

}  // namespace gogp
