"""CSS code container: hx, hz, logical operators, validity test.

Host-side mirror of the reference's ``bposd.css.css_code`` surface
(/root/reference/src/bposd/css.py:7-191): attributes ``hx hz lx lz N K D L Q``,
``compute_dimension``, ``compute_logicals``, ``test``, ``code_params``.  Built
on :mod:`bp_osd_b200.mod2` because the ``ldpc`` package the reference imports is
not available offline.  Not on the decode hot path.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import mod2

__all__ = ["css_code"]


def _as_csr_u8(mat) -> sp.csr_matrix:
    if sp.issparse(mat):
        out = sp.csr_matrix(mat)
    else:
        arr = np.asarray(mat)
        if arr.ndim == 1:
            arr = arr.reshape(1, -1)
        out = sp.csr_matrix(arr)
    out = out.astype(np.uint8)
    out.data %= 2
    out.eliminate_zeros()
    return out


def _logicals_of(h_commute: sp.csr_matrix, h_same: sp.csr_matrix) -> sp.csr_matrix:
    """Vectors in ker(h_commute) that are independent of the rows of h_same.

    Same recipe as css.py:76-88 of the reference: stack h_same on a kernel
    basis and keep the pivot rows that come after rank(h_same).
    """
    ker = mod2.nullspace(h_commute)
    stack = sp.vstack([h_same, ker], format="csr").astype(np.uint8)
    keep = mod2.pivot_rows(stack)[mod2.rank(h_same):]
    return sp.csr_matrix(stack[keep], dtype=np.uint8)


class css_code:
    def __init__(self, hx=None, hz=None, code_distance=np.nan,
                 name="<Unnamed CSS code>", compute_logicals=True):
        empty = hx is None and hz is None
        self.hx = _as_csr_u8(np.zeros((0, 0)) if hx is None else hx)
        self.hz = _as_csr_u8(np.zeros((0, 0)) if hz is None else hz)
        self.lx = sp.csr_matrix((0, 0), dtype=np.uint8)
        self.lz = sp.csr_matrix((0, 0), dtype=np.uint8)
        self.N = np.nan
        self.K = np.nan
        self.D = code_distance
        self.L = np.nan
        self.Q = np.nan
        self.name = name
        if not empty:
            if self.hx.shape[1] != self.hz.shape[1]:
                raise Exception("Error: hx and hz matrices must have equal numbers of columns!")
            if self.hx.shape[1] != 0:
                self.compute_dimension()
                if compute_logicals:
                    self.compute_logicals()

    # -- parameters -------------------------------------------------------
    def compute_dimension(self):
        self.N = int(self.hx.shape[1])
        assert self.N == self.hz.shape[1], "Code block length (N) inconsistent!"
        self.K = self.N - mod2.rank(self.hx) - mod2.rank(self.hz)
        self._weights()
        return self.K

    def _weights(self):
        try:
            col_w = max(int(np.max(self.hx.sum(axis=0))), int(np.max(self.hz.sum(axis=0))))
            row_w = max(int(np.max(self.hx.sum(axis=1))), int(np.max(self.hz.sum(axis=1))))
            self.L, self.Q = col_w, row_w
        except ValueError:
            pass

    def compute_logicals(self):
        if isinstance(self.K, float) and np.isnan(self.K):
            self.compute_dimension()
        self.lx = _logicals_of(self.hz, self.hx)
        self.lz = _logicals_of(self.hx, self.hz)
        return self.lx, self.lz

    def canonical_logicals(self):
        """Rescale lx so that lx @ lz.T = I (mod 2).  The reference's code generator calls this
        (/root/reference/examples/codes/hgp_codes/generate_codes.py:11) and the lx files it ships are in
        this form; lz is left as computed."""
        pair = (self.lx @ self.lz.T).toarray() % 2
        self.lx = sp.csr_matrix((mod2.inverse(pair).astype(np.int64) @ self.lx.toarray().astype(np.int64)) % 2,
                                dtype=np.uint8)
        return self.lx

    @property
    def h(self):
        zx = sp.csr_matrix(self.hz.shape, dtype=np.uint8)
        zz = sp.csr_matrix(self.hx.shape, dtype=np.uint8)
        return sp.hstack([sp.vstack([zx, self.hx]), sp.vstack([self.hz, zz])], format="csr")

    @property
    def l(self):
        zx = sp.csr_matrix(self.lz.shape, dtype=np.uint8)
        zz = sp.csr_matrix(self.lx.shape, dtype=np.uint8)
        return sp.hstack([sp.vstack([zx, self.lx]), sp.vstack([self.lz, zz])], format="csr")

    @property
    def code_params(self):
        return f"({self.L},{self.Q})-[[{self.N},{self.K},{self.D}]]"

    # -- validity ---------------------------------------------------------
    def test(self, show_tests=True):
        """Same five checks, same printed lines, as css.py:122-191 of the reference."""
        ok = True
        say = print if show_tests else (lambda *a, **k: None)
        say(f"{self.name}, {self.code_params}")

        def odd(prod):
            prod = sp.csr_matrix(prod)
            return bool(np.any(prod.data % 2))

        dims = (self.N == self.hz.shape[1] == self.lz.shape[1] == self.lx.shape[1]
                and self.K == self.lz.shape[0] == self.lx.shape[0])
        if dims:
            say(" -Block dimensions: Pass")
        else:
            ok = False
            print(" -Block dimensions incorrect")

        for a, b, label in ((self.hz, self.hx, "hz@hx.T"), (self.hx, self.hz, "hx@hz.T")):
            if odd(a @ b.T):
                ok = False
                print(f" -PCMs commute {label}==0: Fail")
            else:
                say(f" -PCMs commute {label}==0: Pass")

        ker_ok = True
        if self.lx.shape[1] == self.hz.shape[1] and self.lz.shape[1] == self.hx.shape[1]:
            if odd(self.hz @ self.lx.T) or odd(self.hx @ self.lz.T):
                ker_ok = False
        if ker_ok:
            say(r" -lx \in ker{hz} AND lz \in ker{hx}: Pass")
        else:
            ok = False
            print(r" -lx \in ker{hz} AND lz \in ker{hx}: Fail")

        anti = False
        if self.lx.shape[1] == self.lz.shape[1] and self.lx.shape[0] and self.lz.shape[0]:
            prod = sp.csr_matrix(self.lx @ self.lz.T)
            prod.data = prod.data % 2
            anti = mod2.rank(prod) == self.K
        elif self.K == 0:
            anti = True
        if anti:
            say(" -lx and lz anticommute: Pass")
        else:
            ok = False
            print(" -lx and lz anitcommute: Fail")

        if ok:
            say(f" -{self.name} is a valid CSS code w/ params {self.code_params}")
        return ok
