"""Host-side GF(2) linear algebra (bit-packed numpy).

Stands in for the ``ldpc.mod2`` helpers the reference imports for code
construction (call sites: /root/reference/src/bposd/css.py:50,80,84,86,177;
hgp.py:29,36; stab.py:43,51,53,56,69,72,149).  Nothing here is on the decode
hot path: it runs once per code on the host, as BASELINE.json's north_star
prescribes ("css_code/hgp construction stays on the host").

All routines accept a dense ``numpy`` array or a ``scipy.sparse`` matrix with
0/1 entries.  Internally rows are packed 64 columns per ``uint64`` word and the
elimination is a vectorised Gauss-Jordan sweep over columns in ascending order.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = [
    "rank",
    "nullspace",
    "kernel",
    "pivot_rows",
    "pivot_cols",
    "row_echelon",
    "reduced_row_echelon",
    "row_span",
    "row_basis",
    "inverse",
    "pack_rows",
    "unpack_rows",
]

_ONE = np.uint64(1)


def _dense(mat) -> np.ndarray:
    if sp.issparse(mat):
        mat = mat.toarray()
    mat = np.asarray(mat)
    if mat.ndim == 1:
        mat = mat.reshape(1, -1)
    return (mat.astype(np.int64) & 1).astype(np.uint8)


def pack_rows(mat: np.ndarray) -> np.ndarray:
    """Pack a 0/1 matrix [m, n] into uint64 words [m, ceil(n/64)], bit c%64 of word c//64."""
    mat = np.ascontiguousarray(mat, dtype=np.uint8)
    m, n = mat.shape
    nw = max(1, (n + 63) // 64)
    padded = np.zeros((m, nw * 64), dtype=np.uint8)
    padded[:, :n] = mat
    by = np.packbits(padded, axis=1, bitorder="little")
    return np.ascontiguousarray(by).view(np.uint64).reshape(m, nw)


def unpack_rows(packed: np.ndarray, n: int) -> np.ndarray:
    m = packed.shape[0]
    by = np.ascontiguousarray(packed).view(np.uint8).reshape(m, -1)
    return np.unpackbits(by, axis=1, bitorder="little")[:, :n]


def _eliminate(mat: np.ndarray, full: bool = True, n_aug: int = 0):
    """Gauss(-Jordan) elimination scanning columns 0..n-1-n_aug in order.

    Returns (packed_reduced, pivot_cols, row_perm, rank).  Row ``r`` of the
    reduced matrix holds the pivot of ``pivot_cols[r]`` for r < rank; row_perm
    maps reduced row -> original row that was swapped into that position.
    The last ``n_aug`` columns ride along (augmented block) but are never pivots.
    """
    m, n = mat.shape
    a = pack_rows(mat)
    perm = np.arange(m)
    piv_cols = []
    r = 0
    for c in range(n - n_aug):
        if r == m:
            break
        w, b = divmod(c, 64)
        colbits = (a[:, w] >> np.uint64(b)) & _ONE
        cand = np.flatnonzero(colbits[r:])
        if cand.size == 0:
            continue
        p = r + int(cand[0])
        if p != r:
            a[[r, p]] = a[[p, r]]
            perm[[r, p]] = perm[[p, r]]
            colbits[[r, p]] = colbits[[p, r]]
        hit = colbits.astype(bool)
        hit[r] = False
        if not full:
            hit[:r] = False
        if hit.any():
            a[hit] ^= a[r]
        piv_cols.append(c)
        r += 1
    return a, np.asarray(piv_cols, dtype=np.int64), perm, r


def rank(mat) -> int:
    """GF(2) rank."""
    d = _dense(mat)
    if d.size == 0:
        return 0
    # eliminate the thinner orientation
    if d.shape[0] > d.shape[1]:
        d = d.T
    return int(_eliminate(d, full=False)[3])


def pivot_cols(mat) -> np.ndarray:
    """Indices of the first linearly independent columns, scanning left to right."""
    d = _dense(mat)
    if d.size == 0:
        return np.zeros(0, dtype=np.int64)
    return _eliminate(d, full=False)[1]


def pivot_rows(mat) -> np.ndarray:
    """Indices of the rows that are independent of all rows above them (ascending).

    Used by the reference as ``pivot_rows(vstack([h, ker]))[rank_h:]`` to pick
    kernel vectors outside the row space of ``h`` (css.py:82-88).
    """
    d = _dense(mat)
    if d.size == 0:
        return np.zeros(0, dtype=np.int64)
    return _eliminate(np.ascontiguousarray(d.T), full=False)[1]


def reduced_row_echelon(mat):
    """Return [rref, rank, transform, pivot_cols] with transform @ mat = rref (mod 2)."""
    d = _dense(mat)
    m, n = d.shape
    aug = np.concatenate([d, np.eye(m, dtype=np.uint8)], axis=1)
    a, pc, _perm, r = _eliminate(aug, full=True, n_aug=m)
    full = unpack_rows(a, n + m)
    return [full[:, :n].copy(), r, full[:, n:].copy(), pc]


def row_echelon(mat, full: bool = False):
    """Return [row_echelon_form, rank, transform, pivot_cols] (ldpc.mod2.row_echelon shape)."""
    d = _dense(mat)
    m, n = d.shape
    aug = np.concatenate([d, np.eye(m, dtype=np.uint8)], axis=1)
    a, pc, _perm, r = _eliminate(aug, full=full, n_aug=m)
    out = unpack_rows(a, n + m)
    return [out[:, :n].copy(), r, out[:, n:].copy(), pc]


def nullspace(mat) -> sp.csr_matrix:
    """Basis of {x : mat @ x = 0 mod 2} as the rows of a CSR uint8 matrix."""
    d = _dense(mat)
    m, n = d.shape
    if n == 0:
        return sp.csr_matrix((0, 0), dtype=np.uint8)
    a, pc, _perm, r = _eliminate(d, full=True)
    red = unpack_rows(a, n)[:r]
    is_piv = np.zeros(n, dtype=bool)
    is_piv[pc] = True
    free = np.flatnonzero(~is_piv)
    basis = np.zeros((free.size, n), dtype=np.uint8)
    basis[np.arange(free.size), free] = 1
    if r:
        # x_pivot(r) = sum_f red[r, f] x_f
        basis[:, pc] = red[:, free].T
    return sp.csr_matrix(basis, dtype=np.uint8)


kernel = nullspace


def row_basis(mat) -> sp.csr_matrix:
    d = _dense(mat)
    return sp.csr_matrix(d[pivot_rows(d)], dtype=np.uint8)


def row_span(mat) -> sp.csr_matrix:
    """All 2^r linear combinations of the rows (row 0 is the zero vector)."""
    d = _dense(mat)
    m, n = d.shape
    if m > 24:
        raise ValueError("row_span: refusing to enumerate more than 2^24 combinations")
    out = np.zeros((1 << m, n), dtype=np.uint8)
    for i in range(m):
        out[1 << i : 1 << (i + 1)] = out[: 1 << i] ^ d[i]
    return sp.csr_matrix(out, dtype=np.uint8)


def inverse(mat) -> np.ndarray:
    d = _dense(mat)
    m, n = d.shape
    if m != n:
        raise ValueError("inverse: matrix must be square")
    red, r, tr, _ = reduced_row_echelon(d)
    if r != n:
        raise ValueError("inverse: matrix is singular over GF(2)")
    return tr
