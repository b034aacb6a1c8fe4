"""Pins the CPU oracle: the reference's documented known answer (G1), known-answer vectors of the
building blocks, a second independent restatement, brute force on small codes, and invariants."""
import itertools

import numpy as np
import pytest

from bp_osd_b200 import codes
from bp_osd_b200.hgp import hgp
from tests._util import golden_names, load_golden, random_syndromes


def test_g1_readme_known_answer(oracle_mod):
    # /root/reference/README.md:145-216: d=3 surface code, error on qubits {5,12}
    sc = hgp(codes.rep_code(3))
    d = oracle_mod.OracleDecoder(sc.hz, error_rate=0.05, channel_probs=[None], max_iter=sc.N, bp_method="ms",
                                 ms_scaling_factor=0, osd_method="osd_cs", osd_order=7)
    err = np.zeros(sc.N, dtype=int)
    err[[5, 12]] = 1
    syn = sc.hz @ err % 2
    assert (syn == [0, 0, 0, 0, 0, 1]).all()
    out = d.decode(syn)
    assert (out == [0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0]).all()
    residual = (out + err) % 2
    assert not (sc.lz @ residual % 2).any()  # "Logical Error: No"
    assert d.converge and d.iter == 2
    want = [6.2569, 7.3611, 6.2569, 6.2569, 6.2569, 4.0486, 4.0486, 2.9444, -0.3681, 7.3611, 6.2569, 6.2569, 2.9444]
    assert np.allclose(d.log_prob_ratios, want, atol=5e-5)


def test_philox_known_answers(oracle_mod):
    # Random123 kat_vectors for philox4x32-10
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert list(oracle_mod.philox4x32_10(ctr, key)) == want


def test_sampler_statistics_and_split(oracle_mod):
    n, B = 64, 4000
    pz, px, py = np.full(n, 0.02), np.full(n, 0.03), np.full(n, 0.01)
    ex, ez = oracle_mod.sample_errors(5, 0, B, pz, px, py)
    # Z-only 0.02, X-only 0.03, Y (both) 0.01
    assert abs(((ez == 1) & (ex == 0)).mean() - 0.02) < 0.002
    assert abs(((ex == 1) & (ez == 0)).mean() - 0.03) < 0.003
    assert abs(((ex == 1) & (ez == 1)).mean() - 0.01) < 0.002
    # subsequence = global shot index: a shifted window reproduces the overlap
    ex2, _ = oracle_mod.sample_errors(5, 1000, 100, pz, px, py)
    assert (ex2 == ex[1000:1100]).all()


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(oracle_mod, cfg_codes, name):
    g = load_golden(name)
    H = cfg_codes(g["cfg"]).hz
    out = oracle_mod.OracleDecoder(H, error_rate=g["p"], **g["kw"]).decode_batch(g["syndromes"])
    for k in ("osdw", "osd0", "bp", "converge", "iter"):
        assert (out[k] == g[k]).all(), k
    # NaN where the reference arithmetic itself gives inf - inf (product-sum without clipping): same places
    assert np.array_equal(out["llr"], g["llr"], equal_nan=True)


CASES = [
    dict(cfg=1, p=0.08, kw=dict(max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=7), shots=60),
    dict(cfg=1, p=0.12, kw=dict(max_iter=4, bp_method="ms", ms_scaling_factor=0.75, osd_method="osd_e", osd_order=7), shots=60),
    dict(cfg=1, p=0.10, kw=dict(max_iter=9, bp_method="ps", ms_scaling_factor=0, osd_method="osd_cs", osd_order=5), shots=60),
    dict(cfg=2, p=0.06, kw=dict(max_iter=15, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=4), shots=8),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"cfg{c['cfg']}-{c['kw']['bp_method']}-{c['kw']['osd_method']}")
def test_oracle_vs_second_restatement(oracle_mod, cfg_codes, case):
    from oracle.slow_ref import SlowDecoder
    H = cfg_codes(case["cfg"]).hz
    n = H.shape[1]
    kw = case["kw"]
    # the second restatement calls the host libm for tanh/log, so the C oracle is put in the same mode for product-sum
    o = oracle_mod.OracleDecoder(H, error_rate=case["p"], math="libm" if kw["bp_method"] == "ps" else "shared", **kw)
    s = SlowDecoder(H, [case["p"]] * n, kw["max_iter"], kw["bp_method"], kw["ms_scaling_factor"], kw["osd_method"], kw["osd_order"])
    _, syn = random_syndromes(H, case["p"], case["shots"], seed=3)
    for b in range(case["shots"]):
        a = o.decode(syn[b])
        x = s.decode(syn[b])
        assert (a == np.array(x)).all()
        assert (o.osd0_decoding == np.array(s.osd0_decoding)).all()
        assert (o.bp_decoding == np.array(s.bp_decoding)).all()
        assert o.converge == s.converge and o.iter == s.iter
        if kw["bp_method"] == "ms":
            assert (o.log_prob_ratios == np.array(s.llr)).all()
        else:
            assert np.allclose(o.log_prob_ratios, np.array(s.llr), rtol=1e-12, atol=0)


def test_nonuniform_probs_second_restatement(oracle_mod, cfg_codes):
    from oracle.slow_ref import SlowDecoder
    H = cfg_codes(1).hz
    n = H.shape[1]
    rng = np.random.default_rng(11)
    probs = rng.uniform(0.02, 0.2, size=n)
    kw = dict(max_iter=3, bp_method="ms", ms_scaling_factor=0.8, osd_method="osd_cs", osd_order=6)
    o = oracle_mod.OracleDecoder(H, channel_probs=probs, **kw)
    s = SlowDecoder(H, probs, 3, "ms", 0.8, "osd_cs", 6)
    _, syn = random_syndromes(H, 0.12, 80, seed=5)
    for b in range(80):
        assert (o.decode(syn[b]) == np.array(s.decode(syn[b]))).all()


def test_osd_e_full_order_is_exhaustive_minimum(oracle_mod, cfg_codes):
    """OSD-E with osd_order = k enumerates every solution of H x = s: osdw must have minimum weight."""
    H = hgp(codes.rep_code(3)).hz  # 6 x 13, k = 7
    Hd = H.toarray()
    n = Hd.shape[1]
    o = oracle_mod.OracleDecoder(H, error_rate=0.1, max_iter=1, bp_method="ms", ms_scaling_factor=0.5,
                                 osd_method="osd_e", osd_order=7)
    assert o.k == 7
    allx = np.array(list(itertools.product([0, 1], repeat=n)), dtype=np.uint8)
    allsyn = allx @ Hd.T % 2
    _, syn = random_syndromes(H, 0.2, 40, seed=9)
    for b in range(40):
        x = o.decode(syn[b])
        assert (Hd @ x % 2 == syn[b]).all()
        if not o.converge:
            sols = allx[(allsyn == syn[b]).all(1)]
            assert x.sum() == sols.sum(1).min()


@pytest.mark.parametrize("cfg,p", [(1, 0.1), (2, 0.07)])
def test_invariants(oracle_mod, cfg_codes, cfg, p):
    H = cfg_codes(cfg).hz
    o = oracle_mod.OracleDecoder(H, error_rate=p, max_iter=10, bp_method="ms", ms_scaling_factor=0,
                                 osd_method="osd_cs", osd_order=5)
    _, syn = random_syndromes(H, p, 50, seed=2)
    out = o.decode_batch(syn)
    Hd = H.toarray()
    assert ((out["osdw"] @ Hd.T % 2) == syn).all()
    assert ((out["osd0"] @ Hd.T % 2) == syn).all()
    assert (out["osdw"].sum(1) <= out["osd0"].sum(1)).all()
    c = out["converge"].astype(bool)
    assert (out["osd0"][c] == out["bp"][c]).all() and (out["osdw"][c] == out["bp"][c]).all()
    assert (out["iter"][~c] == 10).all()


def test_zero_probability_entries_do_not_nan(oracle_mod, cfg_codes):
    # the reference's example builds an X decoder with all-zero probabilities
    # (examples/qldpc_decode_example.py:11 with css_decode_sim.py:457)
    H = cfg_codes(1).hz
    probs = np.full(H.shape[1], 0.05)
    probs[::3] = 0.0
    with np.errstate(divide="ignore"):
        o = oracle_mod.OracleDecoder(H, channel_probs=probs, max_iter=0, bp_method="ms", ms_scaling_factor=0,
                                     osd_method="osd_cs", osd_order=3)
        _, syn = random_syndromes(H, 0.05, 30, seed=4)
        out = o.decode_batch(syn)
    assert not np.isnan(out["llr"]).any()


def _ulps(got, want):
    """|got - want| in units of the last place of `want` (finite, non-zero values)."""
    import math
    return abs(got - want) / math.ulp(want)


def test_shared_math_is_pinned_to_glibc(oracle_mod):
    """include/bposd_math.h (compiled into the oracle AND the CUDA kernels) against the host libm: the portable tanh / log
    stay within 4 / 1 ulp of glibc's on the ranges product-sum BP visits (classical algorithms with a
    < 2.5 ulp / < 0.85 ulp error measured against long double; glibc: 2.15 / 0.52 ulp), and agree on
    every special value."""
    import math
    L = oracle_mod.lib()
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-25, 25, 40000), rng.uniform(-1, 1, 40000), rng.uniform(-1e-3, 1e-3, 5000),
                         np.ldexp(rng.uniform(-1, 1, 5000), rng.integers(-70, 5, 5000))])
    worst_t = worst_l = 0.0
    for x in xs:
        x = float(x)
        t, tw = L.oracle_math_tanh(x), math.tanh(x)
        if tw != 0.0:
            worst_t = max(worst_t, _ulps(t, tw))
        else:
            assert t == tw
        if abs(tw) < 1.0:
            y = (1 + tw) / (1 - tw)        # the argument product-sum feeds to log (row a5)
            worst_l = max(worst_l, _ulps(L.oracle_math_log(y), math.log(y)) if math.log(y) != 0 else 0.0)
    for y in np.concatenate([rng.uniform(1e-6, 1e6, 20000), 1 + rng.uniform(-1e-3, 1e-3, 5000),
                             np.ldexp(rng.uniform(0.5, 1, 5000), rng.integers(-1000, 1000, 5000))]):
        y = float(y)
        lw = math.log(y)
        if lw != 0.0:
            worst_l = max(worst_l, _ulps(L.oracle_math_log(y), lw))
    assert worst_t <= 4.0, worst_t
    assert worst_l <= 1.0, worst_l
    inf, nan = float("inf"), float("nan")
    assert L.oracle_math_tanh(inf) == 1.0 and L.oracle_math_tanh(-inf) == -1.0 and math.isnan(L.oracle_math_tanh(nan))
    assert L.oracle_math_tanh(40.0) == 1.0 and L.oracle_math_tanh(-40.0) == -1.0 and L.oracle_math_tanh(1e-300) == 1e-300
    assert math.copysign(1.0, L.oracle_math_tanh(-0.0)) == -1.0 and L.oracle_math_tanh(0.0) == 0.0
    assert L.oracle_math_log(inf) == inf and L.oracle_math_log(0.0) == -inf and L.oracle_math_log(1.0) == 0.0
    assert math.isnan(L.oracle_math_log(-1.0)) and math.isnan(L.oracle_math_log(nan))
    assert abs(L.oracle_math_log(5e-324) - math.log(5e-324)) < 1e-12


def test_product_sum_shared_vs_libm_modes(oracle_mod, cfg_codes):
    """The two arithmetic modes of the oracle on config 4 (product-sum + OSD-E 10): decodings agree on every shot that
    stops early and on nearly all others (a last-bit difference needs hundreds of iterations to reach a hard decision)."""
    H = cfg_codes(4).hz
    kw = dict(max_iter=60, bp_method="ps", ms_scaling_factor=0, osd_method="osd_e", osd_order=10)
    _, syn = random_syndromes(H, 0.05, 150, seed=21)
    a = oracle_mod.OracleDecoder(H, error_rate=0.05, math="shared", **kw).decode_batch(syn)
    b = oracle_mod.OracleDecoder(H, error_rate=0.05, math="libm", **kw).decode_batch(syn)
    quick = b["iter"] <= 20
    assert quick.sum() > 50
    assert (a["iter"][quick] == b["iter"][quick]).all() and (a["bp"][quick] == b["bp"][quick]).all()
    assert np.allclose(a["llr"][quick], b["llr"][quick], rtol=1e-9, atol=0)
    assert (a["osdw"] == b["osdw"]).all(1).mean() >= 0.95


@pytest.mark.parametrize("cfg,p,method,alpha,max_iter", [(1, 0.1, "ms", 0.0, 10), (1, 0.1, "ps", 0.0, 10), (2, 0.06, "ms", 0.8, 12)])
def test_serial_schedule_two_restatements_agree(oracle_mod, cfg_codes, cfg, p, method, alpha, max_iter):
    """SURVEY row f4 (ldpc's `schedule="serial"`, never passed by the reference): the C oracle and the independently
    written Python restatement agree bit for bit, in the natural order and in a given `serial_schedule_order`; on an
    acyclic graph (repetition code) the serial schedule converges to the same decoding as the parallel one."""
    from oracle.slow_ref import SlowDecoder
    H = cfg_codes(cfg).hz
    n = H.shape[1]
    rng = np.random.default_rng(cfg)
    tanh, log = oracle_mod.lib().oracle_math_tanh, oracle_mod.lib().oracle_math_log
    for order in (None, rng.permutation(n)):
        _, syn = random_syndromes(H, p, 8, seed=3)
        kw = dict(max_iter=max_iter, bp_method=method, ms_scaling_factor=alpha, osd_method="osd_cs", osd_order=4)
        ref = oracle_mod.OracleDecoder(H, error_rate=p, schedule="serial", serial_schedule_order=order, **kw).decode_batch(syn)
        slow = SlowDecoder(H, [p] * n, max_iter, method, alpha, "osd_cs", 4, tanh=tanh, log=log, schedule="serial",
                           serial_schedule_order=order)
        for b in range(syn.shape[0]):
            x = slow.decode(syn[b])
            assert (np.array(x) == ref["osdw"][b]).all() and (np.array(slow.bp_decoding) == ref["bp"][b]).all()
            assert np.array_equal(np.array(slow.llr), ref["llr"][b], equal_nan=True)
            assert slow.iter == ref["iter"][b] and slow.converge == bool(ref["converge"][b])
    from bp_osd_b200 import codes
    R = codes.rep_code(9)
    _, syn = random_syndromes(R, 0.1, 40, seed=5)
    a = oracle_mod.OracleDecoder(R, error_rate=0.1, max_iter=20, bp_method="ms", ms_scaling_factor=1.0, schedule="serial").decode_batch(syn)
    b = oracle_mod.OracleDecoder(R, error_rate=0.1, max_iter=20, bp_method="ms", ms_scaling_factor=1.0).decode_batch(syn)
    assert a["converge"].all() and b["converge"].all() and (a["bp"] == b["bp"]).all()
    assert (a["iter"] <= b["iter"]).all()   # information travels the whole chain within one serial pass
