"""Hypergraph-product CSS codes.

Host-side mirror of ``bposd.hgp.hgp`` (/root/reference/src/bposd/hgp.py:8-94):
hx = [h1 (x) I_n2 | I_m1 (x) h2^T], hz = [I_n1 (x) h2 | h1^T (x) I_m2]
(hgp.py:48-54), N = n1 n2 + m1 m2, K = k1 k2 + k1t k2t (hgp.py:41-44).
Not on the decode hot path.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import mod2
from .css import css_code, _as_csr_u8

__all__ = ["hgp", "hgp_single", "compute_exact_code_distance"]


def compute_exact_code_distance(h) -> int:
    """Minimum weight of a non-zero codeword of ker(h), by enumerating the kernel span."""
    ker = mod2.nullspace(h).toarray()
    k = ker.shape[0]
    if k == 0:
        return np.inf
    if k > 24:
        raise ValueError("compute_exact_code_distance: kernel dimension too large to enumerate")
    words = np.zeros((1, ker.shape[1]), dtype=np.uint8)
    best = ker.shape[1]
    # Gray-code free doubling keeps memory at 2^k * n bytes for the small seeds this is meant for
    for i in range(k):
        words = np.concatenate([words, words ^ ker[i]], axis=0)
    wts = words[1:].sum(axis=1)
    best = int(wts.min())
    return best


class hgp(css_code):
    def __init__(self, h1, h2=None, compute_distance=False, compute_logicals=True):
        super().__init__()
        h1 = _as_csr_u8(h1)
        h2 = h1.copy() if h2 is None else _as_csr_u8(h2)
        self.h1, self.h2 = h1, h2
        self.m1, self.n1 = h1.shape
        self.m2, self.n2 = h2.shape

        self.r1 = mod2.rank(h1)
        self.r2 = mod2.rank(h2)
        self.k1, self.k1t = self.n1 - self.r1, self.m1 - self.r1
        self.k2, self.k2t = self.n2 - self.r2, self.m2 - self.r2

        eye = lambda k: sp.identity(k, format="csr", dtype=np.uint8)
        self.hx1 = sp.kron(h1, eye(self.n2), format="csr")
        self.hx2 = sp.kron(eye(self.m1), h2.T, format="csr")
        self.hz1 = sp.kron(eye(self.n1), h2, format="csr")
        self.hz2 = sp.kron(h1.T, eye(self.m2), format="csr")
        self.hx = _as_csr_u8(sp.hstack([self.hx1, self.hx2], format="csr"))
        self.hz = _as_csr_u8(sp.hstack([self.hz1, self.hz2], format="csr"))

        self.N = self.n1 * self.n2 + self.m1 * self.m2
        self.K = self.k1 * self.k2 + self.k1t * self.k2t
        self.D = None
        self._weights()

        if compute_logicals:
            self.compute_logicals()

        if compute_distance:
            def dist(h):
                return compute_exact_code_distance(h) if h.shape[1] != mod2.rank(h) else np.inf
            self.d1, self.d2 = dist(h1), dist(h2)
            self.d1t, self.d2t = dist(h1.T), dist(h2.T)
            self.D = int(np.min([self.d1, self.d1t, self.d2, self.d2t]))


class hgp_single(hgp):
    def __init__(self, h1, compute_distance=False):
        super().__init__(h1, compute_distance=compute_distance)
