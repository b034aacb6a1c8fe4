"""Second, independent restatement of the BP+OSD path in pure Python.  TEST INFRASTRUCTURE ONLY.

Written from the published algorithm (arXiv:2005.07016) and the ldpc-v2 conventions listed in
SURVEY.md section 8(a), deliberately with different data structures from oracle/bposd_oracle.c:
per-edge dictionaries for BP, and for OSD an explicit sparse LU factorisation with row swaps,
minimum-row-weight pivot choice, forward and back substitution (the `RowReduce::rref` /
`lu_solve` shape of upstream) instead of a dense bit-packed Gauss-Jordan.  The two must agree
bit for bit (tests/test_oracle.py); that they do is evidence that the OSD result depends only
on the column order, not on the pivot-row heuristic.  Pure-Python loops: small cases only.

Reference call sites this stands behind: /root/reference/README.md:176-197,
src/bposd/css_decode_sim.py:444-463,174-202.
"""
from __future__ import annotations

import math
import sys

import numpy as np
import scipy.sparse as sp

DBL_MAX = sys.float_info.max


class SlowDecoder:
    def __init__(self, h, probs, max_iter, bp_method, ms_scaling_factor, osd_method, osd_order, tanh=None, log=None,
                 schedule="parallel", serial_schedule_order=None):
        # tanh / log: the product-sum functions (default: the host libm, as ldpc calls them; the goldens pass the portable
        # functions of include/bposd_math.h through oracle.oracle.lib() so that both restatements agree bit for bit)
        self._tanh = tanh or math.tanh
        self._log = log or (lambda x: float(np.log(np.float64(x))))
        h = sp.csr_matrix(h).astype(np.uint8)
        h.data %= 2
        h.eliminate_zeros()
        h.sort_indices()
        self.m, self.n = h.shape
        self.rows = [list(map(int, h.indices[h.indptr[i]:h.indptr[i + 1]])) for i in range(self.m)]
        hc = h.tocsc()
        hc.sort_indices()
        self.cols = [list(map(int, hc.indices[hc.indptr[j]:hc.indptr[j + 1]])) for j in range(self.n)]
        self.probs = [float(p) for p in probs]
        self.max_iter = max_iter if max_iter > 0 else self.n
        self.bp_method = bp_method  # "ms" | "ps"
        self.alpha0 = float(ms_scaling_factor)
        self.osd_method = osd_method  # "osd0" | "osd_e" | "osd_cs"
        self.osd_order = 0 if osd_method == "osd0" else int(osd_order)
        self.schedule = schedule
        self.serial_order = list(range(self.n)) if serial_schedule_order is None else [int(j) for j in serial_schedule_order]

    # ---- BP, serial schedule (ldpc option `schedule="serial"`) ----
    def bp_serial(self, synd):
        """Bit after bit: every edge of the bit gets a fresh check-to-bit message from the check's other edges, the bit's
        outgoing messages are the sum of the prior and the OTHER incoming messages (accumulated as a prefix in ascending
        check order plus a suffix from the last edge backwards, the reference's order of additions)."""
        n, m = self.n, self.m
        with np.errstate(divide="ignore"):
            prior = [float(np.log(np.float64(1.0 - p) / np.float64(p))) for p in self.probs]
        out = {(i, j): prior[j] for i in range(m) for j in self.rows[i]}      # bit-to-check
        self.converge, self.iter = False, 0
        llr, dec = list(prior), [0] * n
        for it in range(1, self.max_iter + 1):
            alpha = (1.0 - 2.0 ** (-it)) if self.alpha0 == 0.0 else self.alpha0
            for j in self.serial_order:
                inc = []
                total = prior[j]
                for i in self.cols[j]:
                    others = [out[(i, jj)] for jj in self.rows[i] if jj != j]
                    if self.bp_method == "ps":
                        x = 1.0
                        for v in others:
                            x *= self._tanh(v / 2)
                        with np.errstate(divide="ignore", invalid="ignore"):
                            msg = float((-1.0 if synd[i] else 1.0) * self._log(float(np.float64(1 + x) / np.float64(1 - x))))
                    else:
                        mag = min([abs(v) for v in others], default=DBL_MAX)
                        neg = int(synd[i]) + sum(1 for v in others if v <= 0)
                        msg = (1.0 if neg % 2 == 0 else -1.0) * alpha * mag
                    out[(i, j)] = total
                    total = total + msg
                    inc.append(msg)
                llr[j] = total
                dec[j] = 1 if total <= 0 else 0
                tail = 0.0
                for i, msg in zip(reversed(self.cols[j]), reversed(inc)):
                    out[(i, j)] = out[(i, j)] + tail
                    tail = tail + msg
            self.iter = it
            if all(sum(dec[j] for j in self.rows[i]) % 2 == (int(synd[i]) & 1) for i in range(m)):
                self.converge = True
                break
        self.llr, self.bp_decoding = llr, dec
        return dec

    # ---- BP, flooding schedule -----------------------------------------
    def bp(self, synd):
        if self.schedule == "serial":
            return self.bp_serial(synd)
        n, m = self.n, self.m
        with np.errstate(divide="ignore"):
            prior = [float(np.log(np.float64(1.0 - p) / np.float64(p))) for p in self.probs]
        b2c = {(i, j): prior[j] for i in range(m) for j in self.rows[i]}
        c2b = {}
        self.converge, self.iter = False, 0
        llr = list(prior)
        dec = [0] * n
        for it in range(1, self.max_iter + 1):
            if self.bp_method == "ps":
                for i in range(m):
                    t = 1.0
                    for j in self.rows[i]:
                        c2b[(i, j)] = t
                        t *= self._tanh(b2c[(i, j)] / 2)
                    t = 1.0
                    for j in reversed(self.rows[i]):
                        v = c2b[(i, j)] * t
                        sgn = -1.0 if synd[i] else 1.0
                        with np.errstate(divide="ignore", invalid="ignore"):
                            c2b[(i, j)] = float(sgn * self._log(float(np.float64(1 + v) / np.float64(1 - v))))
                        t *= self._tanh(b2c[(i, j)] / 2)
            else:
                alpha = (1.0 - 2.0 ** (-it)) if self.alpha0 == 0.0 else self.alpha0
                for i in range(m):
                    tot = int(synd[i])
                    t = DBL_MAX
                    for j in self.rows[i]:
                        if b2c[(i, j)] <= 0:
                            tot += 1
                        c2b[(i, j)] = t
                        a = abs(b2c[(i, j)])
                        if a < t:
                            t = a
                    t = DBL_MAX
                    for j in reversed(self.rows[i]):
                        sgn = tot + (1 if b2c[(i, j)] <= 0 else 0)
                        v = c2b[(i, j)]
                        if t < v:
                            v = t
                        ms = 1.0 if sgn % 2 == 0 else -1.0
                        c2b[(i, j)] = v * (ms * alpha)
                        a = abs(b2c[(i, j)])
                        if a < t:
                            t = a
            cand = [0] * m
            for j in range(n):
                t = prior[j]
                for i in self.cols[j]:
                    b2c[(i, j)] = t
                    t = t + c2b[(i, j)]
                llr[j] = t
                dec[j] = 1 if t <= 0 else 0
                if dec[j]:
                    for i in self.cols[j]:
                        cand[i] ^= 1
            self.iter = it
            if all(cand[i] == (int(synd[i]) & 1) for i in range(m)):
                self.converge = True
                break
            for j in range(n):
                t = 0.0
                for i in reversed(self.cols[j]):
                    b2c[(i, j)] = b2c[(i, j)] + t
                    t = t + c2b[(i, j)]
        self.llr, self.bp_decoding = llr, dec
        return dec

    # ---- OSD through an explicit LU factorisation -------------------------
    def _lu(self, order):
        """rref(lower_triangular=True) in the given column order with min-row-weight pivots."""
        m = self.m
        U = [set(r) for r in self.rows]          # row -> set of columns
        L = [set() for _ in range(m)]            # row -> set of pivot ranks
        rowmap = list(range(m))                  # position -> original row
        piv_cols, nonpiv = [], []
        rank, max_rank = 0, min(self.m, self.n)
        for pos, c in enumerate(order):
            if rank == max_rank:
                nonpiv.extend(order[pos:])
                break
            best, best_w = -1, None
            for r in range(rank, m):
                if c in U[r] and (best_w is None or len(U[r]) < best_w):
                    best, best_w = r, len(U[r])
            if best < 0:
                nonpiv.append(c)
                continue
            if best != rank:
                U[best], U[rank] = U[rank], U[best]
                L[best], L[rank] = L[rank], L[best]
                rowmap[best], rowmap[rank] = rowmap[rank], rowmap[best]
            L[rank].add(rank)
            for r in range(rank + 1, m):
                if c in U[r]:
                    U[r] ^= U[rank]
                    L[r].add(rank)
            piv_cols.append(c)
            rank += 1
        return U, L, rowmap, piv_cols, nonpiv, rank

    @staticmethod
    def _lu_solve(U, L, rowmap, piv_cols, rank, y, n):
        m = len(U)
        b = [int(y[rowmap[r]]) & 1 for r in range(m)]
        # forward substitution with unit lower-triangular L (entries L[r] are ranks < r or == r)
        for r in range(m):
            acc = b[r]
            for q in L[r]:
                if q < r:
                    acc ^= b[q]
            b[r] = acc
        x = [0] * n
        for r in range(rank - 1, -1, -1):
            acc = b[r]
            for c in U[r]:
                if c != piv_cols[r]:
                    acc ^= x[c]
            x[piv_cols[r]] = acc
        return x

    def _weight(self, x):
        w = 0.0
        for j in range(self.n):
            if x[j]:
                with np.errstate(divide="ignore"):
                    w += float(np.log(np.float64(1) / np.float64(self.probs[j])))
        return w

    def osd(self, synd):
        n = self.n
        order = sorted(range(n), key=lambda j: self.llr[j])  # Python's sort is stable
        U, L, rowmap, piv_cols, nonpiv, rank = self._lu(order)
        x0 = self._lu_solve(U, L, rowmap, piv_cols, rank, synd, n)
        self.osd0_decoding = list(x0)
        best_x, best_w = list(x0), self._weight(x0)
        k, w = n - rank, self.osd_order
        cands = []
        if self.osd_method == "osd_e" and w > 0:
            for v in range(1, 1 << w):
                cands.append([b for b in range(w) if (v >> b) & 1])
        elif self.osd_method == "osd_cs" and w > 0:
            cands = [[i] for i in range(k)]
            cands += [[i, j] for i in range(w) for j in range(w) if j > i]
        for sel in cands:
            t = [int(s) & 1 for s in synd]
            for q in sel:
                for i in self.cols[nonpiv[q]]:
                    t[i] ^= 1
            x = self._lu_solve(U, L, rowmap, piv_cols, rank, t, n)
            for q in sel:
                x[nonpiv[q]] = 1
            wx = self._weight(x)
            if wx < best_w:
                best_w, best_x = wx, x
        self.osdw_decoding = best_x
        return best_x

    def decode(self, synd):
        self.bp(synd)
        if self.converge:
            self.osd0_decoding = list(self.bp_decoding)
            self.osdw_decoding = list(self.bp_decoding)
        else:
            self.osd(synd)
        return self.osdw_decoding
