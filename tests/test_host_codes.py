"""Host GF(2) toolkit and code builders against the reference's doc/test known answers."""
import os
import numpy as np
import pytest
import scipy.sparse as sp

from bp_osd_b200 import codes, mod2
from bp_osd_b200.css import css_code
from bp_osd_b200.hgp import hgp, compute_exact_code_distance


def test_hamming_matrix_matches_readme():
    # /root/reference/README.md:66-68
    want = np.array([[0, 0, 0, 1, 1, 1, 1], [0, 1, 1, 0, 0, 1, 1], [1, 0, 1, 0, 1, 0, 1]])
    assert (codes.hamming_code(3).toarray() == want).all()


@pytest.mark.parametrize("dense", [False, True])
def test_steane_code(dense):
    # /root/reference/tests/test_css.py:8-27 and README.md:85-88
    h = codes.hamming_code(3)
    if dense:
        h = h.toarray()
    q = css_code(hx=h, hz=h, code_distance=3, name="Steane code")
    assert (q.N, q.K, q.D) == (7, 1, 3)
    assert q.test(show_tests=False)
    assert (q.lx.toarray() == [[1, 1, 1, 0, 0, 0, 0]]).all()
    assert (q.lz.toarray() == [[1, 1, 1, 0, 0, 0, 0]]).all()


def test_invalid_css_code():
    # /root/reference/README.md:120-136: rep-code hx = hz gives K = -5 and test() False
    q = css_code(codes.rep_code(7), codes.rep_code(7))
    assert q.K == -5
    assert not q.test(show_tests=False)


def test_hgp_surface_code():
    # /root/reference/tests/test_hgp.py:9-18, README.md:153
    q = hgp(codes.rep_code(3), codes.rep_code(3), compute_distance=True)
    assert q.test(show_tests=False)
    assert (q.N, q.K, q.D) == (13, 1, 3)
    assert q.code_params == "(2,4)-[[13,1,3]]"


def test_hgp_seed_code_builds():
    # /root/reference/tests/test_hgp.py:21-39 (the 12x16 seed is mkmn_16_4_6)
    q = hgp(codes.mkmn_16_4_6(), codes.mkmn_16_4_6(), compute_distance=True)
    assert (q.N, q.K, q.D) == (400, 16, 6)
    assert q.test(show_tests=False)


def test_mismatched_columns_raise():
    with pytest.raises(Exception):
        css_code(np.ones((2, 3)), np.ones((2, 4)))


@pytest.mark.parametrize("seed", range(5))
def test_rank_nullspace_pivots_random(seed):
    rng = np.random.default_rng(seed)
    m, n = rng.integers(3, 40, size=2)
    a = (rng.random((m, n)) < 0.3).astype(np.uint8)
    r = mod2.rank(a)
    assert r == mod2.rank(a.T)
    ker = mod2.nullspace(a).toarray()
    assert ker.shape == (n - r, n)
    assert not ((a.astype(int) @ ker.T.astype(int)) % 2).any()
    assert mod2.rank(ker) == n - r
    pr = mod2.pivot_rows(a)
    assert len(pr) == r and mod2.rank(a[pr]) == r
    assert (np.diff(pr) > 0).all()
    red, rk, tr, pc = mod2.reduced_row_echelon(a)
    assert rk == r
    assert ((tr.astype(int) @ a.astype(int)) % 2 == red).all()
    assert (red[np.arange(r), pc] == 1).all() and (red[:, pc].sum(0) == 1).all()


def test_row_span_and_distance():
    h = codes.hamming_code(3)
    span = mod2.row_span(h).toarray()
    assert span.shape == (8, 7) and len({tuple(r) for r in span}) == 8
    assert compute_exact_code_distance(h) == 3
    assert compute_exact_code_distance(codes.rep_code(5)) == 5


@pytest.mark.parametrize("cfg,N,K,m,E,rank", [(1, 41, 1, 20, 72, 20), (2, 400, 16, 192, 1344, 192),
                                              (3, 1922, 50, 961, 5766, 936), (4, 882, 24, 441, 2646, 429)])
def test_config_codes(cfg_codes, cfg, N, K, m, E, rank):
    # sizes of SURVEY.md section 8 table
    q = cfg_codes(cfg)
    assert (q.N, q.K) == (N, K)
    assert q.hz.shape == (m, N) and q.hz.nnz == E
    assert mod2.rank(q.hz) == rank
    assert q.lz.shape == (K, N) and q.lx.shape == (K, N)
    assert q.test(show_tests=False)


def test_config5_shape():
    q = codes.config_code(5)
    assert q.hz.shape == (19200, 40000) and q.hz.nnz == 134400
    assert not ((q.hx @ q.hz.T).data % 2).any()


@pytest.mark.parametrize("tag", ["400_16_6", "625_25_8", "900_36_10"])
def test_reference_shipped_logicals(tag):
    """Golden data FROM THE REFERENCE (not from the oracle): the lx/lz files it ships under
    examples/codes/hgp_codes were written by the real ldpc.mod2 pipeline (generate_codes.py).  Our hgp() +
    GF(2) toolkit must give the same lz bit for bit; lx is shipped in canonical form (lx lz^T = I) from an
    older release, so it is pinned as a coset: canonical, commuting, and equal to ours modulo X stabilisers."""
    import os
    from bp_osd_b200.hgp import hgp
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_hgp_logicals.npz"))
    N, K, _ = (int(x) for x in tag.split("_"))
    shape = tuple(g[f"seed_shape_{tag}"])
    seed = np.unpackbits(g[f"seed_{tag}"], axis=1)[:, :shape[1]]
    lx_ref = np.unpackbits(g[f"lx_{tag}"], axis=1)[:, :N].astype(np.int64)
    lz_ref = np.unpackbits(g[f"lz_{tag}"], axis=1)[:, :N].astype(np.int64)
    q = hgp(seed)
    assert (q.N, q.K) == (N, K)
    assert (q.lz.toarray() == lz_ref).all()
    q.canonical_logicals()
    lx = q.lx.toarray().astype(np.int64)
    assert ((lx @ lz_ref.T) % 2 == np.eye(K, dtype=np.int64)).all()
    assert ((lx_ref @ lz_ref.T) % 2 == np.eye(K, dtype=np.int64)).all()
    hx = q.hx.toarray()
    assert not ((q.hz.toarray().astype(np.int64) @ lx_ref.T) % 2).any()
    assert mod2.rank(np.vstack([hx, (lx ^ lx_ref).astype(np.uint8)])) == mod2.rank(hx)


def test_regenerated_code_files_match_the_shipped_ones(tmp_path):
    """Row f3: codes.regenerate_example_codes() redoes examples/codes/hgp_codes/generate_codes.py from the seeds kept in
    bp_osd_b200.codes.  Every hx / hz / lz file the reference ships ([[400,16,6]] hx, hz, lz; [[625,25,8]] and
    [[900,36,10]] lz) must come out byte for byte (sha256 fixture made from the reference's files by
    tests/golden/make_code_golden.py); the four hx / hz files it lists in .MISSING_LARGE_BLOBS:1-4 are regenerated and
    checked against the logicals it does ship; the compact .npz form round-trips."""
    import hashlib
    import json
    import os
    here = os.path.dirname(__file__)
    with open(os.path.join(here, "golden", "ref_code_file_sha256.json")) as fh:
        shipped = json.load(fh)
    g = np.load(os.path.join(here, "golden", "ref_hgp_logicals.npz"))
    params = codes.regenerate_example_codes(tmp_path)
    assert params == ["(4,7)-[[400,16,6]]", "(4,7)-[[625,25,8]]", "(4,7)-[[900,36,10]]"]
    same = 0
    for name, digest in shipped.items():
        if name.endswith("_lx.txt"):
            continue  # shipped in an older release's canonical form: pinned as a coset by test_reference_shipped_logicals
        with open(tmp_path / name, "rb") as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == digest, name
        same += 1
    assert same == 5
    for p_, tag in zip(params, ("400_16_6", "625_25_8", "900_36_10")):
        N, K, D = (int(x) for x in tag.split("_"))
        c = codes.load_code(tmp_path / f"hgp_{p_}.npz")
        assert (c["N"], c["K"], c["D"]) == (N, K, D)
        for key in ("hx", "hz"):  # the dense text and the compact form hold the same matrix
            dense = codes.load_alist_txt(tmp_path / f"hgp_{p_}_{key}.txt")
            assert (c[key].toarray() == dense).all()
            assert codes.load_code(str(tmp_path / f"hgp_{p_}_{key}.txt"))["h"].shape == dense.shape
        lx_ref = np.unpackbits(g[f"lx_{tag}"], axis=1)[:, :N].astype(np.int64)
        lz_ref = np.unpackbits(g[f"lz_{tag}"], axis=1)[:, :N].astype(np.int64)
        hx, hz = c["hx"].toarray().astype(np.int64), c["hz"].toarray().astype(np.int64)
        assert not ((hx @ hz.T) % 2).any()
        assert not ((hx @ lz_ref.T) % 2).any() and not ((hz @ lx_ref.T) % 2).any()  # the shipped logicals commute with them
        assert (c["lz"].toarray() == lz_ref).all()
        assert mod2.rank(hx) + mod2.rank(hz) == N - K


def test_reference_package_name_shim():
    """A script written against the reference imports `bposd`, `bposd.hgp`, `bposd.css`, `bposd.css_decode_sim`
    (/root/reference/src/bposd/__init__.py:1, README.md:150-216): the shim package resolves every name to bp_osd_b200."""
    import bposd
    from bposd import bposd_decoder, BpOsdDecoder          # noqa: F401
    from bposd.css import css_code
    from bposd.hgp import hgp, hgp_single                  # noqa: F401
    from bposd.css_decode_sim import css_decode_sim
    import bp_osd_b200
    assert bposd_decoder is bp_osd_b200.bposd_decoder and css_code is bp_osd_b200.css_code
    assert css_decode_sim.__module__ == "bp_osd_b200.css_decode_sim"
    q = hgp(codes.rep_code(3))                              # README.md:176-180
    assert (q.N, q.K) == (13, 1)
    assert os.path.exists(os.path.join(bposd.get_include(), "bposd_b200.h"))
