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
    "save_dense_txt",
    "save_code",
    "load_code",
    "example_seed",
    "regenerate_example_codes",
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


# the other two (3,4)-regular seeds of examples/codes/classical_seed_codes (mkmn_20_5_8.txt, mkmn_24_6_10.txt): the inputs
# from which generate_codes.py:5-20 builds the [[625,25,8]] and [[900,36,10]] codes whose hx/hz files the reference
# lists in .MISSING_LARGE_BLOBS:1-4.  Row supports, as above (data, not code).
_MKMN_20_5_8_ROWS = (
    (3, 4, 17, 19), (1, 6, 14, 19), (5, 9, 15, 16), (2, 6, 13, 16), (8, 9, 11, 18), (4, 10, 14, 15), (7, 9, 12, 17),
    (1, 2, 3, 12), (5, 8, 12, 14), (7, 10, 11, 16), (3, 5, 7, 18), (0, 11, 13, 19), (0, 4, 6, 18), (0, 2, 8, 15),
    (1, 10, 13, 17),
)
_MKMN_24_6_10_ROWS = (
    (1, 14, 15, 18), (6, 7, 15, 22), (3, 8, 19, 23), (2, 11, 13, 23), (9, 11, 16, 19), (5, 8, 17, 22), (0, 1, 7, 16),
    (5, 12, 13, 14), (0, 10, 18, 23), (4, 16, 21, 22), (1, 4, 6, 11), (3, 7, 14, 17), (8, 10, 20, 21), (0, 4, 19, 20),
    (2, 5, 6, 10), (9, 13, 18, 21), (2, 12, 15, 17), (3, 9, 12, 20),
)
_SEEDS = {"mkmn_16_4_6": (_MKMN_16_4_6_ROWS, 16), "mkmn_20_5_8": (_MKMN_20_5_8_ROWS, 20), "mkmn_24_6_10": (_MKMN_24_6_10_ROWS, 24)}


def example_seed(name: str) -> np.ndarray:
    """One of the reference's classical seed codes by file stem ("mkmn_16_4_6", "mkmn_20_5_8", "mkmn_24_6_10")."""
    rows, n = _SEEDS[name]
    h = np.zeros((len(rows), n), dtype=np.uint8)
    for i, cols in enumerate(rows):
        h[i, list(cols)] = 1
    return h


def mkmn_16_4_6() -> np.ndarray:
    return example_seed("mkmn_16_4_6")


# ---- on-disk code formats ---------------------------------------------------------------
def save_dense_txt(path, mat) -> None:
    """Write a matrix the way the reference stores its example codes: ``np.savetxt`` of the dense matrix with the default
    ``%.18e`` format (generate_codes.py:17-20; 1.9 MB for a 192 x 400 matrix).  Byte-identical to the shipped files."""
    a = mat.toarray() if sp.issparse(mat) else np.asarray(mat)
    np.savetxt(path, a.astype(np.float64))


_CODE_KEYS = ("hx", "hz", "lx", "lz")


def save_code(path, hx=None, hz=None, lx=None, lz=None, **meta) -> None:
    """Compact form of a CSS code: one ``.npz`` holding every given matrix as CSR (``<name>_indptr``, ``<name>_indices``,
    ``<name>_shape``; all entries are 1 over GF(2)) plus scalar metadata (``name``, ``N``, ``K``, ``D`` ...).
    The [[400,16,6]] code takes 7 KB instead of the 4.2 MB of its four dense text files."""
    out = {}
    for key, mat in zip(_CODE_KEYS, (hx, hz, lx, lz)):
        if mat is None:
            continue
        c = sp.csr_matrix(mat).astype(np.uint8)
        c.data %= 2
        c.eliminate_zeros()
        c.sort_indices()
        out[key + "_indptr"] = c.indptr.astype(np.int32)
        out[key + "_indices"] = c.indices.astype(np.int32)
        out[key + "_shape"] = np.asarray(c.shape, dtype=np.int64)
    for k, v in meta.items():
        out["meta_" + k] = np.asarray(v)
    np.savez_compressed(path, **out)


def load_code(path) -> dict:
    """Read a code written by ``save_code`` (``.npz``) -- or, for a ``.txt`` path, one dense text matrix of the reference
    (returned under the key ``"h"``).  Matrices come back as ``scipy.sparse.csr_matrix`` of uint8; metadata under its name."""
    if str(path).endswith(".txt"):
        return {"h": sp.csr_matrix(load_alist_txt(path))}
    z = np.load(path)
    out = {}
    for key in _CODE_KEYS:
        if key + "_indptr" in z.files:
            ip, ix = z[key + "_indptr"], z[key + "_indices"]
            out[key] = sp.csr_matrix((np.ones(ix.size, dtype=np.uint8), ix, ip), shape=tuple(int(x) for x in z[key + "_shape"]))
    for f in z.files:
        if f.startswith("meta_"):
            v = z[f]
            out[f[5:]] = v.item() if v.ndim == 0 else v
    return out


def regenerate_example_codes(out_dir, seeds=("mkmn_16_4_6", "mkmn_20_5_8", "mkmn_24_6_10"), dense_txt=True):
    """What examples/codes/hgp_codes/generate_codes.py does, from the seeds kept in this module: ``hgp(seed,
    compute_distance=True)`` + ``canonical_logicals()`` for each seed, written as ``hgp_<code_params>_{hx,hz,lx,lz}.txt``
    (the reference's dense text, optional) and as one compact ``hgp_<code_params>.npz``.  Restores the four hx/hz files
    the reference lists as missing (.MISSING_LARGE_BLOBS:1-4).  Returns the list of code_params strings."""
    import os
    from .hgp import hgp
    os.makedirs(out_dir, exist_ok=True)
    done = []
    for name in seeds:
        q = hgp(example_seed(name), compute_distance=True)
        q.canonical_logicals()
        params = q.code_params  # "(4,7)-[[400,16,6]]", the string the reference puts in its file names
        base = os.path.join(out_dir, f"hgp_{params}")
        if dense_txt:
            for key, mat in (("hx", q.hx), ("hz", q.hz), ("lx", q.lx), ("lz", q.lz)):
                save_dense_txt(f"{base}_{key}.txt", mat)
        save_code(base + ".npz", q.hx, q.hz, q.lx, q.lz, name=f"hgp_{params}", N=q.N, K=q.K, D=q.D, seed=name)
        done.append(params)
    return done


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
