"""Classical seed codes and the parity-check matrices of the benchmark configs.

Replaces the ``ldpc.codes`` generators the reference's docs and tests use
(/root/reference/README.md:57,120,148; tests/test_css.py:9; tests/test_hgp.py:10)
and builds the five BASELINE.json configurations (size table).
Host only; runs once per code.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = [
    "rep_code",
    "ring_code",
    "hamming_code",
    "circulant",
    "shift_matrix",
    "regular_ldpc",
    "lifted_product_b1",
    "config_code",
    "load_alist_txt",
]


def rep_code(distance: int) -> sp.csr_matrix:
    """(d-1) x d repetition-code checks: row i has ones at i, i+1."""
    d = int(distance)
    rows = np.repeat(np.arange(d - 1), 2)
    cols = np.stack([np.arange(d - 1), np.arange(1, d)], axis=1).ravel()
    return sp.csr_matrix((np.ones(rows.size, np.uint8), (rows, cols)), shape=(d - 1, d))


def ring_code(distance: int) -> sp.csr_matrix:
    """d x d cyclic repetition code."""
    d = int(distance)
    rows = np.repeat(np.arange(d), 2)
    cols = np.stack([np.arange(d), (np.arange(d) + 1) % d], axis=1).ravel()
    return sp.csr_matrix((np.ones(rows.size, np.uint8), (rows, cols)), shape=(d, d))


def hamming_code(rank: int) -> sp.csr_matrix:
    """r x (2^r - 1) Hamming checks; column j (1-based) is j in binary, MSB in row 0.

    Matches the matrix printed at /root/reference/README.md:66-68 for rank 3.
    """
    r = int(rank)
    n = (1 << r) - 1
    j = np.arange(1, n + 1)
    h = np.stack([(j >> (r - 1 - i)) & 1 for i in range(r)]).astype(np.uint8)
    return sp.csr_matrix(h)


def shift_matrix(size: int, power: int) -> np.ndarray:
    """Permutation matrix of x^power in F2[x]/(x^size - 1): row i has a one at (i+power) % size."""
    p = np.zeros((size, size), dtype=np.uint8)
    p[np.arange(size), (np.arange(size) + power) % size] = 1
    return p


def circulant(size: int, exponents) -> np.ndarray:
    """Circulant matrix of the polynomial sum_k x^k, k in exponents."""
    out = np.zeros((size, size), dtype=np.uint8)
    for k in exponents:
        out ^= shift_matrix(size, int(k))
    return out


def regular_ldpc(m: int, n: int, col_w: int, row_w: int, seed: int = 7) -> np.ndarray:
    """A (col_w,row_w)-regular m x n parity-check matrix without repeated edges.

    Socket-matching construction with a fixed seed; repeated edges are repaired
    by swapping with random sockets until the graph is simple.
    """
    assert m * row_w == n * col_w
    rng = np.random.default_rng(seed)
    vs = np.repeat(np.arange(n), col_w)
    for _attempt in range(200):
        cs = np.repeat(np.arange(m), row_w)
        rng.shuffle(cs)
        for _ in range(10000):
            key = cs.astype(np.int64) * n + vs
            _, first = np.unique(key, return_index=True)
            dup = np.setdiff1d(np.arange(key.size), first)
            if dup.size == 0:
                break
            other = rng.integers(0, key.size, size=dup.size)
            cs[dup], cs[other] = cs[other].copy(), cs[dup].copy()
        else:
            continue
        h = np.zeros((m, n), dtype=np.uint8)
        h[cs, vs] = 1
        if (h.sum(0) == col_w).all() and (h.sum(1) == row_w).all():
            return h
    raise RuntimeError("regular_ldpc: failed to build a simple regular graph")


def lifted_product_b1():
    """[[882,24]] lifted-product code (Panteleev-Kalachev B1, lift 63); returns (hx, hz).

    A is 7x7 over circulants of size 63 with A[j][j]=x^27, A[j+1][j]=x^54,
    A[j+2][j]=1 (indices mod 7); B=(1+x+x^6) I_7; hx=[A|B], hz=[B^T|A^T]
    (provenance paragraph).
    """
    l, w = 63, 7
    a = np.zeros((w * l, w * l), dtype=np.uint8)
    for j in range(w):
        for di, pw in ((0, 27), (1, 54), (2, 0)):
            i = (j + di) % w
            a[i * l : (i + 1) * l, j * l : (j + 1) * l] ^= shift_matrix(l, pw)
    b1 = circulant(l, (0, 1, 6))
    b = np.kron(np.eye(w, dtype=np.uint8), b1)
    hx = np.concatenate([a, b], axis=1)
    hz = np.concatenate([b.T, a.T], axis=1)
    return sp.csr_matrix(hx), sp.csr_matrix(hz)


def load_alist_txt(path) -> np.ndarray:
    """Load the reference's dense text matrices (np.savetxt float or int text).

    Format of /root/reference/examples/codes/**/*.txt (generate_codes.py:17-20).
    """
    return (np.loadtxt(path).astype(np.int64) & 1).astype(np.uint8)


# the 12x16 (3,4)-regular seed shipped as examples/codes/classical_seed_codes/mkmn_16_4_6.txt
# and spelled out in /root/reference/tests/test_hgp.py:22-37; stored as column supports.
_MKMN_16_4_6_ROWS = (
    (0, 1, 4, 5), (2, 6, 7, 12), (3, 4, 6, 13), (5, 8, 9, 15),
    (1, 7, 8, 11), (8, 12, 13, 14), (0, 7, 13, 15), (3, 5, 10, 12),
    (2, 3, 11, 15), (4, 9, 10, 11), (1, 6, 10, 14), (0, 2, 9, 14),
)


def mkmn_16_4_6() -> np.ndarray:
    h = np.zeros((12, 16), dtype=np.uint8)
    for i, cols in enumerate(_MKMN_16_4_6_ROWS):
        h[i, list(cols)] = 1
    return h


def config_code(cfg: int, logicals: bool = True):
    """Return the CSS code object of BASELINE.json ``configs[cfg-1]`` (cfg in 1..5)."""
    from .hgp import hgp
    from .css import css_code

    if cfg == 1:
        return hgp(rep_code(5))
    if cfg == 2:
        return hgp(mkmn_16_4_6())
    if cfg == 3:
        return hgp(circulant(31, (0, 2, 5)))
    if cfg == 4:
        hx, hz = lifted_product_b1()
        return css_code(hx, hz, compute_logicals=logicals)
    if cfg == 5:
        return hgp(regular_ldpc(120, 160, 3, 4, seed=7), compute_logicals=False)
    raise ValueError(f"unknown config {cfg}")
