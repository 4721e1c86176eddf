"""2D block-cyclic Cholesky / log marginal likelihood across the GPUs of one box
(BASELINE configs[4]: N = 131072 does not fit one GPU).

One process per GPU.  K is cut into NB x NB blocks; block (I, J), J <= I, lives on
process (I mod Pr, J mod Pc) of a Pr x Pc grid (rank = r * Pc + c), stored in a local
row-major matrix whose block rows / columns are the owned ones in increasing order.
Every rank builds its own blocks from the replicated inputs (no communication for
the build).  Right-looking factorisation, one block column k at a time:

  1. the owner of (k, k) factors it (single-GPU blocked Cholesky) and broadcasts
     L_kk and its tile inverses;
  2. the ranks of process column k mod Pc solve their part of the panel
     A_Ik <- A_Ik L_kk^-T and broadcast it (NCCL over NVLink), so that every rank
     holds the whole panel, organised by process row;
  3. every rank updates its own trailing blocks A_IJ -= L_Ik L_Jk^T with the
     DMMA GEMM, one launch per owned block row (M = NB, N = owned columns <= I, K = NB).

This module is orchestration only: memory, views and collectives come from torch /
torch.distributed (plumbing); every flop is a kernel of libgogp_b200.so reached
through the device-level C-ABI (gogp_dev_*).  The compute backend is injected so
that the same orchestration runs in a 2-rank gloo test on CPU with a NumPy backend
that lives under tests/.
"""
import math

import numpy as np
import torch


class CudaBlocks:
    """Compute backend: the gogp_dev_* entry points on torch CUDA tensors."""

    def __init__(self, simil, noise, ndim, device):
        import ctypes as C
        from . import _lib
        self.C, self._lib, self.L = C, _lib, _lib.lib()
        self.device = torch.device("cuda", device)
        sd = simil.Descriptor()
        nd = noise.Descriptor() if noise is not None else None
        self.h = C.c_void_p()
        st = self.L.gogp_create(ndim, sd, len(sd), simil.NTheta(), nd, len(nd) if nd is not None else 0,
                                noise.NTheta() if noise is not None else 0, device, C.byref(self.h))
        self._ck(st)
        self._nts = simil.NTheta()
        self.info = torch.zeros(1, dtype=torch.int32, device=self.device)

    def _ck(self, st):
        if st != self._lib.OK:
            raise RuntimeError("gogp: " + self.L.gogp_last_error(self.h).decode())

    def _stream(self):
        return self.C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        assert t.stride(-1) == 1
        return t.data_ptr()

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float64, device=self.device)

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    def set_inputs(self, X, block=128):
        X = np.ascontiguousarray(X, dtype=np.float64)
        self._ck(self.L.gogp_dev_set_inputs(self.h, self._lib.dptr(X.reshape(-1)), X.shape[0], block))

    def cov_block(self, theta_s, theta_n, row0, rows, col0, cols, diagonal, out):
        ts = np.ascontiguousarray(theta_s, dtype=np.float64)
        tn = np.ascontiguousarray(theta_n, dtype=np.float64)
        self._ck(self.L.gogp_dev_cov_block(self.h, self._lib.dptr(ts), self._lib.dptr(tn), row0, rows, col0, cols,
                                           1 if diagonal else 0, self._p(out), out.stride(0), self._stream()))

    def potrf(self, A, winv, base):
        self._ck(self.L.gogp_dev_potrf(self.h, self._p(A), A.stride(0), A.shape[0], self._p(winv),
                                       self.info.data_ptr(), base, self._stream()))

    def trsm(self, B, L, winv):
        self._ck(self.L.gogp_dev_trsm(self.h, self._p(B), B.stride(0), B.shape[0], self._p(L), L.stride(0),
                                      L.shape[0], self._p(winv), self._stream()))

    def gemm(self, Cm, A, B, alpha, beta, lower=False):
        self._ck(self.L.gogp_dev_gemm(self.h, self._p(Cm), Cm.stride(0), self._p(A), A.stride(0), self._p(B),
                                      B.stride(0), Cm.shape[0], Cm.shape[1], A.shape[1], alpha, beta,
                                      1 if lower else 0, self._stream()))

    def sumlogdiag(self, Lb, nvalid, out2):
        self._ck(self.L.gogp_dev_sumlogdiag(self.h, self._p(Lb), Lb.stride(0), nvalid, out2.data_ptr(),
                                            self._stream()))

    def gemv_sub(self, B, v, acc, scratch):
        self._ck(self.L.gogp_dev_gemv_sub(self.h, self._p(B), B.stride(0), B.shape[0], B.shape[1], v.data_ptr(),
                                          acc.data_ptr(), scratch.data_ptr(), self._stream()))

    def trsv(self, Lb, winv, rhs, z):
        self._ck(self.L.gogp_dev_trsv(self.h, self._p(Lb), Lb.stride(0), self._p(winv), rhs.data_ptr(),
                                      z.data_ptr(), Lb.shape[0], self._stream()))

    def trtri_t(self, Lb, winv, out):
        assert Lb.stride(0) == out.stride(0)
        self._ck(self.L.gogp_dev_trtri_t(self.h, self._p(Lb), Lb.stride(0), Lb.shape[0], self._p(winv), self._p(out),
                                         self._stream()))

    def trace_block(self, theta_s, alpha, blk, row0, col0, acc, scratch):
        ts = np.ascontiguousarray(theta_s, dtype=np.float64)
        self._ck(self.L.gogp_dev_trace_block(self.h, self._lib.dptr(ts), alpha.data_ptr(), self._p(blk), blk.stride(0),
                                             row0, blk.shape[0], col0, blk.shape[1], acc.data_ptr(),
                                             scratch.data_ptr(), self._stream()))

    def noise_eval(self, theta_n):
        tn = np.ascontiguousarray(theta_n, dtype=np.float64)
        var = self.C.c_double(0.0)
        dlog = np.zeros(max(1, len(tn)))
        self._ck(self.L.gogp_noise_eval(self.h, self._lib.dptr(tn), self.C.byref(var), self._lib.dptr(dlog)))
        return var.value, dlog[:len(tn)]

    def nsimil(self):
        return self._nts

    # look-ahead plumbing: run a batch of launches on a side stream, ordered after the
    # current point of the main (current) stream; wait_side() orders the main stream after it
    def side(self, fn):
        if not hasattr(self, "_side"):
            self._side = torch.cuda.Stream(self.device)
            self._side_done = None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            fn()
            self._side_done = torch.cuda.Event()
            self._side_done.record(self._side)
        return self._side_done

    def wait_event(self, ev):
        """order the main (current) stream after a batch returned by side()"""
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def wait_side(self):
        if getattr(self, "_side_done", None) is not None:
            torch.cuda.current_stream(self.device).wait_event(self._side_done)
            self._side_done = None

    def bad_pivot(self):
        return int(self.info.item())

    def launches(self):
        return int(self.L.gogp_launch_count(self.h))

    def close(self):
        if self.h:
            self.L.gogp_destroy(self.h)
            self.h = None


def default_grid(world):
    """Pr x Pc with Pr >= Pc: the panel solve is shared by the Pr ranks of one process
    column, and with whole-panel broadcasts the shape does not change the traffic."""
    pc = 1
    while (pc * 2) * (pc * 2) <= world and world % (pc * 2) == 0:
        pc *= 2
    return world // pc, pc


class BlockCyclicCholesky:
    def __init__(self, backend, N, NB, rank=0, world=1, grid=None, dist=None):
        assert NB % 128 == 0
        self.be, self.N, self.NB, self.rank, self.world, self.dist = backend, N, NB, rank, world, dist
        self.Pr, self.Pc = grid if grid is not None else default_grid(world)
        assert self.Pr * self.Pc == world
        self.r, self.c = rank // self.Pc, rank % self.Pc
        self.nb = (N + NB - 1) // NB
        self.my_rows = [I for I in range(self.nb) if I % self.Pr == self.r]
        self.my_cols = [J for J in range(self.nb) if J % self.Pc == self.c]
        self.ri = {I: i for i, I in enumerate(self.my_rows)}
        self.ci = {J: j for j, J in enumerate(self.my_cols)}
        self.local = backend.empty(max(1, len(self.my_rows)) * NB, max(1, len(self.my_cols)) * NB)
        # the whole panel of a step, organised by process row; two generations, because the
        # bulk of step k's trailing update (side stream) overlaps step k+1's panel work
        self.rows_of = [[I for I in range(self.nb) if I % self.Pr == rr] for rr in range(self.Pr)]
        self.panels = [[backend.empty(max(1, len(self.rows_of[rr])), NB, NB) for rr in range(self.Pr)]
                       for _ in range(2)]
        self.pc_bufs = [backend.empty(max(1, len(self.my_cols)), NB, NB) for _ in range(2)]
        self.lkk = backend.empty(NB, NB)
        self.wkk = backend.empty(NB // 128, 128, 128)
        self.winv_diag = {}   # tile inverses of the diagonal blocks this rank owns (for the solve)
        self.logdet2 = backend.zeros(2)
        self.sumlog = backend.zeros(1)
        self.factored = False
        self.V = None        # L^-T, upper block triangular, same local layout (invert)
        self.z = None        # L^-1 y, replicated (solve_lml)
        self.alpha = None    # K^-1 y, replicated (solve_alpha)
        self.have_kinv = False
        import os
        self.diag_after_bulk = os.environ.get("GOGP_DIAG_AFTER_BULK", "1") == "1"

    def owner(self, I, J):
        return (I % self.Pr) * self.Pc + (J % self.Pc)

    def block(self, I, J):
        i, j, NB = self.ri[I], self.ci[J], self.NB
        return self.local[i * NB:(i + 1) * NB, j * NB:(j + 1) * NB]

    def _bcast(self, t, src, group=None):
        if self.world > 1:
            self.dist.broadcast(t, src=src, group=group)

    def _groups(self):
        """process-row and process-column groups (every rank creates all of them, in the same order)"""
        if self.world == 1 or hasattr(self, "_row_groups"):
            return
        Pr, Pc = self.Pr, self.Pc
        self._row_groups = [self.dist.new_group([r * Pc + c for c in range(Pc)]) for r in range(Pr)]
        self._col_groups = [self.dist.new_group([r * Pc + c for r in range(Pr)]) for c in range(Pc)]

    def vblock(self, I, J):
        i, j, NB = self.ri[I], self.ci[J], self.NB
        return self.V[i * NB:(i + 1) * NB, j * NB:(j + 1) * NB]

    # ---- build: every rank evaluates its own blocks from the replicated inputs ----
    def build(self, theta_simil, theta_noise):
        NB = self.NB
        for I in self.my_rows:
            for J in self.my_cols:
                if J <= I:
                    self.be.cov_block(theta_simil, theta_noise, I * NB, NB, J * NB, NB, I == J, self.block(I, J))
        self.factored = False

    # ---- factor ---------------------------------------------------------------------
    def _factor_diag(self, k):
        """Owner of (k, k): factor the block, keep its tile inverses, stage it for the broadcast."""
        if self.rank != self.owner(k, k):
            return
        NB, be = self.NB, self.be
        d = self.block(k, k)
        w = be.empty(NB // 128, 128, 128)
        be.potrf(d, w, k * NB)
        be.sumlogdiag(d, min(NB, self.N - k * NB), self.logdet2)
        self.sumlog += self.logdet2[:1]
        self.winv_diag[k] = w
        self.lkk.copy_(d)
        self.wkk.copy_(w)

    def factor(self, lookahead=True):
        NB, be = self.NB, self.be
        self.sumlog.zero_()
        self._factor_diag(0)
        for k in range(self.nb - 1):
            ok = self.owner(k, k)
            panel, pc_buf = self.panels[k % 2], self.pc_bufs[k % 2]
            self._bcast(self.lkk, ok)
            self._bcast(self.wkk, ok)
            # panel solve on the ranks of process column k mod Pc -- every source rank solves its own
            # rows FIRST, so the Pr solves run concurrently -- then the whole panel is broadcast
            kc = k % self.Pc
            srcs = []
            for rr in range(self.Pr):
                rows = [I for I in self.rows_of[rr] if I > k]
                if not rows:
                    continue
                src = rr * self.Pc + kc
                buf = panel[rr][:len(rows)]
                srcs.append((src, buf))
                if self.rank == src:
                    i0, j = self.ri[rows[0]], self.ci[k]
                    sub = self.local[i0 * NB:(i0 + len(rows)) * NB, j * NB:(j + 1) * NB]
                    be.trsm(sub, self.lkk, self.wkk)
                    buf.copy_(sub.reshape(len(rows), NB, NB))
            for src, buf in srcs:
                self._bcast(buf, src)
            # rows of the panel that face my block columns, in increasing J
            cols = [J for J in self.my_cols if J > k]
            for n, J in enumerate(cols):
                rr = J % self.Pr
                pc_buf[n].copy_(panel[rr][self.rows_of[rr].index(J) - self._first_after(rr, k)])
            pcm = pc_buf.reshape(-1, NB)
            mine = [I for I in self.my_rows if I > k]
            # the previous step's bulk update (side stream) touches the same blocks: order after it
            be.wait_side()
            first = 0
            if mine and cols and cols[0] == k + 1:
                # look-ahead: block column k+1 first, on the main stream -- it is all the next
                # step's factor / panel solve / broadcasts depend on
                mypanel = panel[self.r][:len(mine)].reshape(-1, NB)   # my block rows > k, contiguous
                i0, j = self.ri[mine[0]], self.ci[k + 1]
                be.gemm(self.local[i0 * NB:(i0 + len(mine)) * NB, j * NB:(j + 1) * NB], mypanel, pcm[:NB], -1.0, 1.0)
                first = 1
            # block (k+1, k+1) is final now.  Its owner factors it on the priority main stream while
            # the bulk update runs on the side stream (diag_after_bulk, default): the TMA GEMM keeps one
            # 197 KB CTA per SM, so every CTA that retires frees a whole SM for the 135 KB leaf kernel
            # (measured on 1 GPU, N = 32768: 426 -> 399 ms).  With the 2-CTA/SM cp.async GEMM the leaf
            # starved behind 90 KB CTAs and factoring BEFORE queueing the bulk was faster (GOGP_DIAG_AFTER_BULK=0).
            if not self.diag_after_bulk:
                self._factor_diag(k + 1)
            if not mine or not cols:
                if self.diag_after_bulk:
                    self._factor_diag(k + 1)
                continue
            base = self._first_after(self.r, k)

            def bulk(first=first, mine=mine, cols=cols, panel=panel, pcm=pcm, base=base):
                # the rest of my trailing blocks: one GEMM per owned block row
                for I in mine:
                    ncols = sum(1 for J in cols if J <= I)
                    if ncols <= first:
                        continue
                    i, j0 = self.ri[I], self.ci[cols[first]]
                    Cm = self.local[i * NB:(i + 1) * NB, j0 * NB:(j0 + ncols - first) * NB]
                    A = panel[self.r][self.rows_of[self.r].index(I) - base]
                    be.gemm(Cm, A, pcm[first * NB:ncols * NB], -1.0, 1.0)

            if lookahead:
                be.side(bulk)
            else:
                bulk()
            if self.diag_after_bulk:
                self._factor_diag(k + 1)
        be.wait_side()
        self.factored = True

    def _first_after(self, rr, k):
        """index in rows_of[rr] of the first block row > k (the panel buffers start there)"""
        return sum(1 for I in self.rows_of[rr] if I <= k)

    def logdet(self):
        """log det K = 2 sum log L_ii (one scalar all-reduce)."""
        t = self.sumlog.clone()
        if self.world > 1:
            self.dist.all_reduce(t)
        return 2.0 * float(t.item())

    # ---- z = L^-1 y and the log marginal likelihood (gp/gp.go:244-253) -----------------
    def solve_lml(self, y):
        NB, be, nb = self.NB, self.be, self.nb
        ypad = be.zeros(nb * NB)
        ypad[:self.N] = torch.as_tensor(np.asarray(y, dtype=np.float64)).to(ypad.device)
        zcols = be.zeros(max(1, len(self.my_cols)) * NB)   # z_J for my block columns, in order
        acc = be.zeros(NB)
        scratch = be.zeros(NB)
        zk = be.zeros(NB)
        ss = be.zeros(1)
        zfull = be.zeros(nb * NB)
        for k in range(nb):
            acc.zero_()
            if k % self.Pr == self.r:
                ncols = sum(1 for J in self.my_cols if J < k)
                if ncols:
                    i = self.ri[k]
                    be.gemv_sub(self.local[i * NB:(i + 1) * NB, :ncols * NB], zcols[:ncols * NB], acc, scratch)
            if self.world > 1:
                self.dist.all_reduce(acc)   # NB doubles; ranks outside the process row add zeros
            ok = self.owner(k, k)
            if self.rank == ok:
                rhs = ypad[k * NB:(k + 1) * NB] + acc
                be.trsv(self.block(k, k), self.winv_diag[k], rhs, zk)
            self._bcast(zk, ok)
            ss += (zk * zk).sum()
            if k in self.ci:
                j = self.ci[k]
                zcols[j * NB:(j + 1) * NB].copy_(zk)
            zfull[k * NB:(k + 1) * NB].copy_(zk)
        self.z = zfull
        quad = float(ss.item())
        return -0.5 * self.N * math.log(2 * math.pi) - 0.5 * self.logdet() - 0.5 * quad

    # ---- K^-1 and the gradient across the grid (gp/gp.go:418-499) --------
    # Everything stays in the one GEMM form C = beta C + alpha A B^T, as on one GPU:
    #   V = L^-T (upper):  V_cc = L_cc^-T,  V_ic = -(sum_{k=i}^{c-1} V_ik L_ck^T) L_cc^-T   (i < c)
    #   alpha = V z,       z = L^-1 y from solve_lml
    #   K^-1 = V V^T:      K^-1_ij = sum_{c >= i} V_ic V_jc^T                               (i >= j)
    #   grad_q = sum over owned blocks of the fused trace kernel, then one all-reduce of P+1 doubles.
    def invert(self):
        """V = L^-T, one block column at a time.  Step c: block row c of L goes down the process
        columns; the owner of V_ik forms V_ik L_ck^T (one GEMM per owned block row, K = its owned
        k in [i, c)); the partial sums are reduced along process rows to process column c mod Pc,
        which solves with L_cc."""
        assert self.factored
        NB, be, nb = self.NB, self.be, self.nb
        self._groups()
        if self.V is None:
            self.V = be.zeros(*self.local.shape)
        else:
            self.V.zero_()
        maxc, maxr = max(1, len(self.my_cols)), max(1, len(self.my_rows))
        lrow = be.empty(NB * maxc * NB)
        sbuf = be.empty(maxr * NB * NB)
        colg = self._col_groups[self.c] if self.world > 1 else None
        rowg = self._row_groups[self.r] if self.world > 1 else None
        for c in range(nb):
            oc, pcc = self.owner(c, c), c % self.Pc
            if self.c == pcc:
                if self.rank == oc:
                    be.trtri_t(self.block(c, c), self.winv_diag[c], self.vblock(c, c))
                    self.lkk.copy_(self.block(c, c))
                    self.wkk.copy_(self.winv_diag[c])
                if c > 0:
                    self._bcast(self.lkk, oc, colg)
                    self._bcast(self.wkk, oc, colg)
            if c == 0:
                continue
            ncl = sum(1 for J in self.my_cols if J < c)   # my block columns k < c
            mine = [I for I in self.my_rows if I < c]     # my block rows i < c
            buf = lrow[:NB * ncl * NB].view(NB, ncl * NB)
            if ncl:
                src = (c % self.Pr) * self.Pc + self.c
                if self.rank == src:
                    i = self.ri[c]
                    buf.copy_(self.local[i * NB:(i + 1) * NB, :ncl * NB])
                self._bcast(buf, src, colg)
            if not mine:
                continue
            S = sbuf[:len(mine) * NB * NB].view(len(mine) * NB, NB)
            for n, I in enumerate(mine):
                j0 = sum(1 for J in self.my_cols if J < I)
                Sn = S[n * NB:(n + 1) * NB]
                if ncl > j0:
                    i = self.ri[I]
                    be.gemm(Sn, self.V[i * NB:(i + 1) * NB, j0 * NB:ncl * NB], buf[:, j0 * NB:ncl * NB], -1.0, 0.0)
                else:
                    Sn.zero_()
            dst = self.r * self.Pc + pcc
            if self.world > 1 and self.Pc > 1:
                self.dist.reduce(S, dst=dst, group=rowg)
            if self.rank == dst:
                be.trsm(S, self.lkk, self.wkk)
                j = self.ci[c]
                self.V[:len(mine) * NB, j * NB:(j + 1) * NB].copy_(S)

    def solve_alpha(self):
        """alpha = K^-1 y = V z (replicated): block-row GEMVs with the owned blocks, one all-reduce."""
        assert self.V is not None and self.z is not None
        NB, be, nb = self.NB, self.be, self.nb
        zc = be.zeros(max(1, len(self.my_cols)) * NB)
        for j, J in enumerate(self.my_cols):
            zc[j * NB:(j + 1) * NB].copy_(self.z[J * NB:(J + 1) * NB])
        acc = be.zeros(nb * NB)
        scratch = be.zeros(NB)
        for I in self.my_rows:
            j0 = sum(1 for J in self.my_cols if J < I)
            if j0 < len(self.my_cols):
                i = self.ri[I]
                be.gemv_sub(self.V[i * NB:(i + 1) * NB, j0 * NB:len(self.my_cols) * NB], zc[j0 * NB:],
                            acc[I * NB:(I + 1) * NB], scratch)
        if self.world > 1:
            self.dist.all_reduce(acc)
        acc.neg_()
        self.alpha = acc
        return acc

    def kinv(self):
        """K^-1 = V V^T into the local blocks (the factor is overwritten).  Step c: block column c of V
        goes to every rank, organised by process row as in factor(); every rank adds
        V_ic V_jc^T to its blocks (i, j), j <= i <= c -- one GEMM per owned block row, on the side
        stream, while the next block column is being broadcast."""
        assert self.V is not None
        NB, be, nb = self.NB, self.be, self.nb
        events = {}
        for c in range(nb):
            panel, pc_buf = self.panels[c % 2], self.pc_bufs[c % 2]
            be.wait_event(events.pop(c - 2, None))   # step c-2 read the buffers about to be overwritten
            kc = c % self.Pc
            for rr in range(self.Pr):
                nrows = sum(1 for I in self.rows_of[rr] if I <= c)
                if not nrows:
                    continue
                src = rr * self.Pc + kc
                buf = panel[rr][:nrows]
                if self.rank == src:
                    j = self.ci[c]
                    buf.copy_(self.V[:nrows * NB, j * NB:(j + 1) * NB].reshape(nrows, NB, NB))
                self._bcast(buf, src)
            mine = [I for I in self.my_rows if I <= c]
            cols = [J for J in self.my_cols if J <= c]
            if not mine or not cols:
                continue
            for n, J in enumerate(cols):
                rr = J % self.Pr
                pc_buf[n].copy_(panel[rr][self.rows_of[rr].index(J)])
            pcm = pc_buf.reshape(-1, NB)

            def bulk(c=c, mine=mine, cols=cols, panel=panel, pcm=pcm):
                for I in mine:
                    ncols = sum(1 for J in cols if J <= I)
                    if not ncols:
                        continue
                    i = self.ri[I]
                    A = panel[self.r][self.rows_of[self.r].index(I)]
                    be.gemm(self.local[i * NB:(i + 1) * NB, :ncols * NB], A, pcm[:ncols * NB], 1.0,
                            0.0 if c == I else 1.0)

            events[c] = be.side(bulk)
        be.wait_side()
        self.factored = False
        self.have_kinv = True

    def gradient(self, theta_simil, theta_noise):
        """d LML / d log theta (gp/gp.go:434-486): the fused trace over every owned block of
        W = alpha alpha^T - K^-1, one all-reduce of ntheta_simil + 1 doubles; the noise parameters
        follow from tr(W) on the host (all shipped noises are input-independent)."""
        assert self.have_kinv and self.alpha is not None
        NB, be = self.NB, self.be
        nts = be.nsimil()
        acc = be.zeros(nts + 1)
        scratch = be.zeros((NB // 128) ** 2 * (nts + 1))
        for I in self.my_rows:
            for J in self.my_cols:
                if J <= I:
                    be.trace_block(theta_simil, self.alpha, self.block(I, J), I * NB, J * NB, acc, scratch)
        if self.world > 1:
            self.dist.all_reduce(acc)
        g = acc.cpu().numpy() if hasattr(acc, "cpu") else np.asarray(acc)
        _, dlog = be.noise_eval(theta_noise)
        return np.concatenate([g[:nts], 0.5 * g[nts] * np.asarray(dlog, dtype=np.float64)])

    def lml_and_gradient(self, theta_simil, theta_noise, y):
        """build + factor + solve + K^-1 + trace: one LML + gradient evaluation across the grid."""
        self.build(theta_simil, theta_noise)
        self.factor()
        lml = self.solve_lml(y)
        self.invert()
        self.solve_alpha()
        self.kinv()
        return lml, self.gradient(theta_simil, theta_noise)
