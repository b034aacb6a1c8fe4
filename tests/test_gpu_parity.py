"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

fp64 mode: bit-exact osdw/osd0/bp decodings, converge flags, iteration counts and LLRs for min-sum;
product-sum LLRs within the stated tolerance (CUDA libm != glibc libm bit for bit).
fp32 mode: statistical (logical error rate inside the 95% CI of the oracle's).
Nothing here reads /root/reference.
"""
import numpy as np
import pytest

from bp_osd_b200 import codes
from bp_osd_b200.hgp import hgp
from tests._util import golden_names, load_golden, random_syndromes

pytestmark = pytest.mark.gpu

MS_CS7 = dict(max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=7)


@pytest.fixture(scope="module")
def torch_cuda(cuda_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def gpu_decode(torch, H, syn, kernel=None, threads=0, precision=64, probs=None, error_rate=None, **kw):
    from bp_osd_b200 import BpOsdDecoder
    d = BpOsdDecoder(H, error_rate=error_rate, channel_probs=probs, precision=precision, **kw)
    if kernel is not None or threads:
        d.set_tuning(bp_kernel=kernel, bp_threads=threads)
    r = d.decode_batch(torch.tensor(syn, device="cuda"))
    torch.cuda.synchronize()
    out = dict(osdw=r.osdw_decoding.cpu().numpy(), osd0=r.osd0_decoding.cpu().numpy(), bp=r.bp_decoding.cpu().numpy(),
               llr=r.log_prob_ratios.cpu().numpy(), converge=r.converge.cpu().numpy(), iter=r.iter.cpu().numpy())
    return d, out


def assert_exact(out, ref, llr_exact=True):
    for k in ("osdw", "osd0", "bp"):
        bad = np.flatnonzero((out[k] != ref[k]).any(1))
        assert bad.size == 0, f"{k} differs for shots {bad[:10]}"
    assert (out["converge"] == ref["converge"].astype(bool)).all()
    assert (out["iter"] == ref["iter"]).all()
    if llr_exact:
        # equal_nan: product-sum shots whose sums overflowed hold NaN in the same places (x86 and sm_100a generate different
        # quiet-NaN payloads, the one thing that legitimately differs); every other value is compared bit for bit
        assert np.array_equal(out["llr"], ref["llr"], equal_nan=True), "log_prob_ratios not bit-exact"


def test_g1_readme_known_answer_on_gpu(torch_cuda):
    # /root/reference/README.md:145-216 through the drop-in class, legacy signature
    from bp_osd_b200 import bposd_decoder
    sc = hgp(codes.rep_code(3))
    bpd = bposd_decoder(sc.hz, error_rate=0.05, channel_probs=[None], max_iter=sc.N, bp_method="ms",
                        ms_scaling_factor=0, osd_method="osd_cs", osd_order=7)
    error = np.zeros(sc.N).astype(int)
    error[[5, 12]] = 1
    syndrome = sc.hz @ error % 2
    bpd.decode(syndrome)
    assert (bpd.osdw_decoding == [0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0]).all()
    residual = (bpd.osdw_decoding + error) % 2
    assert not (sc.lz @ residual % 2).any()
    assert bpd.converge and bpd.iter == 2
    assert (bpd.osd0_decoding == bpd.bp_decoding).all()


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("kernel", [0, 1, 2, 3])
def test_golden_fixtures(torch_cuda, cfg_codes, name, kernel):
    g = load_golden(name)
    if g["kw"]["bp_method"] == "ps" and kernel == 3:
        pytest.skip("the cluster kernel is a min-sum kernel; product-sum runs on the in-place and the two-array kernels")
    H = cfg_codes(g["cfg"]).hz
    d, out = gpu_decode(torch_cuda, H, g["syndromes"], kernel=kernel, error_rate=g["p"], **g["kw"])
    assert_exact(out, g)


@pytest.mark.parametrize("cfg,p,B,kw", [
    (1, 0.05, 3000, MS_CS7),
    (1, 0.10, 1500, dict(max_iter=7, bp_method="ms", ms_scaling_factor=0.625, osd_method="osd_e", osd_order=9)),
    (2, 0.05, 1500, MS_CS7),
    (2, 0.08, 400, dict(MS_CS7, osd_method="osd0", osd_order=0)),
    (3, 0.05, 600, MS_CS7),
    (3, 0.06, 200, dict(MS_CS7, max_iter=30)),
    (4, 0.05, 300, dict(max_iter=20, bp_method="ms", ms_scaling_factor=0.9, osd_method="osd_e", osd_order=10)),
], ids=["d5", "d5-osd_e9", "hgp400", "hgp400-osd0", "hgp1922", "hgp1922-it30", "lp882-ms-e10"])
@pytest.mark.parametrize("kernel", [0, 1, 2, 3])
def test_min_sum_bit_exact(torch_cuda, oracle_mod, cfg_codes, cfg, p, B, kw, kernel):
    H = cfg_codes(cfg).hz
    _, syn = random_syndromes(H, p, B, seed=12345)
    ref = oracle_mod.OracleDecoder(H, error_rate=p, **kw).decode_batch(syn)
    d, out = gpu_decode(torch_cuda, H, syn, kernel=kernel, error_rate=p, **kw)
    assert d.info()["bp_kernel"] == kernel
    assert_exact(out, ref)
    st = d.stats()
    assert st["bp_converged"] == int(ref["converge"].sum())
    assert st["bp_iterations"] == int(ref["iter"].sum())
    assert st["osd_invocations"] == B - st["bp_converged"]


@pytest.mark.parametrize("threads", [32, 64, 256, 1024])
def test_thread_counts_do_not_change_results(torch_cuda, oracle_mod, cfg_codes, threads):
    H = cfg_codes(2).hz
    _, syn = random_syndromes(H, 0.06, 300, seed=8)
    ref = oracle_mod.OracleDecoder(H, error_rate=0.06, **MS_CS7).decode_batch(syn)
    for kernel in (1, 2):
        _, out = gpu_decode(torch_cuda, H, syn, kernel=kernel, threads=threads, error_rate=0.06, **MS_CS7)
        assert_exact(out, ref)


def test_product_sum_lifted_product(torch_cuda, oracle_mod, cfg_codes):
    """Config 4: product-sum BP + OSD-E 10 at max_iter = n.  tanh / log are the portable FMA-free functions of
    include/bposd_math.h on BOTH sides (the oracle's default `math="shared"`), so everything is compared exactly:
    decodings, converge flags, iteration counts and every bit of the LLRs (NaN where the reference arithmetic itself
    produces inf - inf)."""
    H = cfg_codes(4).hz
    kw = dict(max_iter=0, bp_method="ps", ms_scaling_factor=0, osd_method="osd_e", osd_order=10)
    _, syn = random_syndromes(H, 0.05, 300, seed=12345)
    ref = oracle_mod.OracleDecoder(H, error_rate=0.05, **kw).decode_batch(syn)
    assert (ref["iter"] > 100).sum() >= 20  # long-running shots are part of the comparison
    for kernel in (None, 0, 1, 2):   # None: the automatic choice, the in-place kernel (2)
        d, out = gpu_decode(torch_cuda, H, syn, kernel=kernel, error_rate=0.05, **kw)
        assert d.info()["bp_kernel"] == (2 if kernel is None else kernel)
        assert_exact(out, ref, llr_exact=False)
        # NaN payloads are the one thing that legitimately differs (x86 and sm_100a generate different quiet NaNs)
        assert np.array_equal(out["llr"], ref["llr"], equal_nan=True), "log_prob_ratios not bit-exact"


def test_device_math_is_the_host_math(torch_cuda, oracle_mod, cfg_codes):
    """include/bposd_math.h, device side against host side, element by element and bit for bit: the in-range division
    sequence the kernels use instead of the compiler's out-of-line fp64 division (against the host's IEEE `/`), tanh, log
    and the (1 + x) / (1 - x) of the product-sum update, on 2^22 arguments each drawn from the ranges the check update
    produces plus the special values."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    d = BpOsdDecoder(cfg_codes(1).hz, error_rate=0.05, bp_method="ps", osd_method="osd0")
    rng = np.random.default_rng(2024)
    N = 1 << 22
    sign = lambda k: np.where(rng.random(k) < 0.5, -1.0, 1.0)
    logu = lambda k, lo, hi: np.exp2(rng.uniform(lo, hi, k)) * sign(k)
    dev = lambda a: torch.tensor(a, device="cuda", dtype=torch.float64)

    def same(fn, a, b=None):
        got = d.math_probe(fn, dev(a), None if b is None else dev(b)).cpu().numpy()
        ref = oracle_mod.math_map(fn, a, b)
        bad = np.flatnonzero((got.view(np.uint64) != ref.view(np.uint64)) & ~(np.isnan(got) & np.isnan(ref)))
        assert bad.size == 0, f"{fn}: {bad.size} of {a.size} differ, first a={a[bad[0]]!r} device={got[bad[0]]!r} host={ref[bad[0]]!r}"

    # division: generic operands over the whole range the header allows, then the three shapes that occur
    same("div", logu(N, -60, 70), logu(N, -60, 70))
    # hard cases: divisors whose mantissa is all ones or nearly (where an unbiased reciprocal seed ties), and quotients
    # within a few ulps of a representable number or of a midpoint between two
    k = rng.integers(1, 64, N).astype(np.float64)
    same("div", logu(N, -40, 40), (2.0 - k * 2.0 ** -52) * np.exp2(rng.integers(-30, 30, N)) * sign(N))
    bq, tq = logu(N, -20, 20), logu(N, -20, 20)
    near = np.nextafter(bq * tq, np.where(rng.random(N) < 0.5, np.inf, -np.inf))
    same("div", np.where(rng.random(N) < 0.3, bq * tq, near), bq)
    half = tq * (1 + 2.0 ** -53)                                         # t + half an ulp (rounded): products near midpoints
    same("div", bq * half, bq)
    u = np.expm1(rng.uniform(-2, 44, N))
    same("div", np.where(u > 1, 2.0, -u), u + 2.0)                      # inside tanh
    f = rng.uniform(np.sqrt(0.5) - 1, np.sqrt(2) - 1, N)
    f[:1000] = 0.0
    same("div", f, 2.0 + f)                                             # inside log
    x = np.tanh(logu(N, -30, 5))
    x[:8] = [1.0, -1.0, 0.0, -0.0, np.nan, 1 - 2.0 ** -53, -1 + 2.0 ** -53, 0.5]
    inside = x[8:][np.abs(x[8:]) < 1]                                   # bpm_div's contract excludes b = 0 ...
    same("div", 1 + inside, 1 - inside)
    same("ratio", x)                                                    # ... which ps_ratio selects around: x = 1 -> +inf
    # tanh over the message range, tiny and huge arguments, specials
    t = np.concatenate([rng.uniform(-25, 25, N // 2), logu(N // 2 - 8, -40, 8),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-300, 22.0, 1.7976931348623157e308]])
    same("tanh", t)
    # log over the quotient range [2^-54, 2^54], near 1, specials (0 -> -inf, inf, NaN, negative, subnormal)
    q = np.concatenate([np.abs(logu(N // 2, -54, 54)), 1 + rng.uniform(-0.3, 0.45, N // 2 - 8),
                        [0.0, np.inf, np.nan, -1.0, 5e-324, 1.0, 2.0 ** -54, 2.0 ** 54]])
    same("log", q)


@pytest.mark.parametrize("cfg,p,B", [(1, 0.08, 800), (2, 0.06, 500)])
def test_product_sum_in_place_kernel_other_degree_classes(torch_cuda, oracle_mod, cfg_codes, cfg, p, B):
    """Product-sum on the in-place kernel for the (4, 2) class and for an IRREGULAR code of the (8, 4) class (rows of 7 in
    slots of 8: the pads hold +max, whose tanh is exactly 1; bits of 3 and 4 edges), with non-uniform priors on the second:
    bit for bit against the oracle and against the two-array kernel."""
    H = cfg_codes(cfg).hz
    n = H.shape[1]
    kw = dict(max_iter=0, bp_method="ps", ms_scaling_factor=0, osd_method="osd_cs", osd_order=5)
    rng = np.random.default_rng(3 + cfg)
    probs = np.full(n, p) if cfg == 1 else rng.uniform(0.5 * p, 1.5 * p, size=n)
    e = (rng.random((B, n)) < probs).astype(np.uint8)
    syn = np.asarray((H @ e.T) % 2, dtype=np.uint8).T.copy()
    ref = oracle_mod.OracleDecoder(H, channel_probs=probs, **kw).decode_batch(syn)
    assert (ref["converge"] == 0).sum() >= 3
    for kernel in (2, 1):
        d, out = gpu_decode(torch_cuda, H, syn, kernel=kernel, probs=probs, **kw)
        assert d.info()["bp_kernel"] == kernel
        assert_exact(out, ref)


def test_received_vector_input_and_method_aliases(torch_cuda, oracle_mod, cfg_codes):
    """SURVEY row f4 (ldpc options the reference never reaches): input_vector_type='received_vector' decodes the syndrome of
    the received word and returns the corrected word r + decoding; 'auto' picks by length; the pre-v2 spellings 'ms_log' /
    'ps_log' name the same two updates.  The serial schedule stays rejected."""
    from bp_osd_b200 import BpOsdDecoder
    H = cfg_codes(2).hz
    m, n = H.shape
    kw = dict(max_iter=20, ms_scaling_factor=0, osd_method="osd_cs", osd_order=4)
    rng = np.random.default_rng(5)
    r = (rng.random(n) < 0.05).astype(np.uint8)
    syn = np.asarray(H @ r) % 2
    ref = oracle_mod.OracleDecoder(H, error_rate=0.05, bp_method="ms", **kw)
    ref.decode(syn)
    for ivt, vec in (("received_vector", r), ("auto", r), ("auto", syn), ("syndrome", syn)):
        d = BpOsdDecoder(H, error_rate=0.05, bp_method="ms_log", input_vector_type=ivt, **kw)
        out = d.decode(vec)
        want = ref.osdw_decoding ^ r if len(vec) == n else ref.osdw_decoding
        assert (out == want).all(), ivt
        if len(vec) == n:
            assert not (np.asarray(H @ out) % 2).any()   # the corrected word has a zero syndrome
            assert (d.osd0_decoding == (ref.osd0_decoding ^ r)).all() and (d.bp_decoding == (ref.bp_decoding ^ r)).all()
    ps = oracle_mod.OracleDecoder(H, error_rate=0.05, bp_method="ps", **kw)
    ps.decode(syn)
    assert (BpOsdDecoder(H, error_rate=0.05, bp_method="ps_log", **kw).decode(syn) == ps.osdw_decoding).all()
    with pytest.raises(ValueError):   # the randomised serial schedule is refused (test_serial_schedule covers the serial one)
        BpOsdDecoder(H, error_rate=0.05, bp_method="ms", schedule="serial", random_serial_schedule=True, **kw)
    with pytest.raises(ValueError):
        BpOsdDecoder(H, error_rate=0.05, bp_method="ms", input_vector_type="codeword", **kw)


@pytest.mark.parametrize("cfg,p,B,method,alpha,max_iter,osd_method,osd_order,custom_order,nonuniform", [
    (1, 0.10, 700, "ms", 0.0, 8, "osd_cs", 7, False, False),
    (1, 0.10, 300, "ps", 0.0, 8, "osd_e", 6, True, False),
    (2, 0.06, 300, "ms", 0.75, 20, "osd_cs", 7, True, True),
    (2, 0.06, 100, "ps", 0.0, 12, "osd0", 0, False, False),
    (3, 0.05, 70, "ms", 0.0, 30, "osd_cs", 7, False, False),   # 70 shots: three warps, the last one partly idle
])
def test_serial_schedule(torch_cuda, oracle_mod, cfg_codes, cfg, p, B, method, alpha, max_iter, osd_method, osd_order, custom_order,
                         nonuniform):
    """ldpc's serial BP schedule (an option the reference never passes) -- one thread per shot on the GPU --
    bit for bit against the oracle: min-sum and product-sum, natural and given bit order, non-uniform channel, followed
    by OSD on the shots that did not converge; device and host input, and the single-shot decode()."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = cfg_codes(cfg).hz
    n = H.shape[1]
    rng = np.random.default_rng(100 + cfg)
    order = rng.permutation(n) if custom_order else None
    probs = rng.uniform(0.5 * p, 1.5 * p, size=n) if nonuniform else np.full(n, p)
    kw = dict(max_iter=max_iter, bp_method=method, ms_scaling_factor=alpha, osd_method=osd_method, osd_order=osd_order)
    e = (rng.random((B, n)) < probs).astype(np.uint8)
    syn = np.asarray((H @ e.T) % 2, dtype=np.uint8).T.copy()
    ref = oracle_mod.OracleDecoder(H, channel_probs=probs, schedule="serial", serial_schedule_order=order, **kw).decode_batch(syn)
    assert 0 < (ref["converge"] == 0).sum() < B
    d = BpOsdDecoder(H, channel_probs=probs, schedule="serial", serial_schedule_order=order, **kw)
    r = d.decode_batch(torch.tensor(syn, device="cuda"))
    out = dict(osdw=r.osdw_decoding.cpu().numpy(), osd0=r.osd0_decoding.cpu().numpy(), bp=r.bp_decoding.cpu().numpy(),
               llr=r.log_prob_ratios.cpu().numpy(), converge=r.converge.cpu().numpy(), iter=r.iter.cpu().numpy())
    assert_exact(out, ref)
    rh = d.decode_batch(syn[:40], return_llr=False)          # host input, LLRs through the failed-shot workspace
    assert (np.asarray(rh.osdw_decoding) == ref["osdw"][:40]).all() and (np.asarray(rh.iter) == ref["iter"][:40]).all()
    x = d.decode(syn[0])
    assert (np.asarray(x) == ref["osdw"][0]).all() and d.iter == ref["iter"][0] and bool(d.converge) == bool(ref["converge"][0])
    # the parallel schedule on the same decoder arguments is a different algorithm: it must not be what ran
    par = oracle_mod.OracleDecoder(H, channel_probs=probs, **kw).decode_batch(syn)
    assert (par["iter"] != ref["iter"]).any()


def test_nonuniform_and_zero_probabilities(torch_cuda, oracle_mod, cfg_codes):
    """Soft OSD weights in ascending-index fp64 order, and p = 0 entries (prior +inf, weight inf) without NaN."""
    H = cfg_codes(2).hz
    n = H.shape[1]
    rng = np.random.default_rng(21)
    probs = rng.uniform(0.01, 0.15, size=n)
    probs[::17] = 0.0
    kw = dict(max_iter=12, bp_method="ms", ms_scaling_factor=0.8, osd_method="osd_cs", osd_order=6)
    e = (rng.random((400, n)) < probs).astype(np.uint8)
    syn = np.asarray((H @ e.T) % 2, dtype=np.uint8).T.copy()
    with np.errstate(divide="ignore"):
        ref = oracle_mod.OracleDecoder(H, channel_probs=probs, **kw).decode_batch(syn)
    for kernel in (1, 2):
        d, out = gpu_decode(torch_cuda, H, syn, kernel=kernel, probs=probs, **kw)
        assert_exact(out, ref)
        assert not np.isnan(out["llr"]).any()
    assert (ref["converge"] == 0).sum() > 50  # the soft-weight OSD path was exercised


def test_update_channel_probs(torch_cuda, oracle_mod, cfg_codes):
    # css_decode_sim.py:229,248: new probabilities change priors and OSD weights of the next decode
    from bp_osd_b200 import BpOsdDecoder
    H = cfg_codes(1).hz
    n = H.shape[1]
    kw = dict(max_iter=3, bp_method="ms", ms_scaling_factor=0.7, osd_method="osd_cs", osd_order=5)
    d = BpOsdDecoder(H, error_rate=0.1, **kw)
    o = oracle_mod.OracleDecoder(H, error_rate=0.1, **kw)
    _, syn = random_syndromes(H, 0.12, 200, seed=4)
    newp = np.random.default_rng(2).uniform(0.02, 0.3, size=n)
    newp.setflags(write=False)  # the harness hands over read-only arrays (css_decode_sim.py:432-434)
    d.update_channel_probs(newp)
    o.update_channel_probs(newp)
    for b in range(40):
        assert (d.decode(syn[b]) == o.decode(syn[b])).all()
        assert (d.osd0_decoding == o.osd0_decoding).all() and (d.bp_decoding == o.bp_decoding).all()
        assert d.converge == o.converge and d.iter == o.iter
        assert (d.log_prob_ratios == o.log_prob_ratios).all()


def test_per_shot_priors(torch_cuda, oracle_mod, cfg_codes):
    """decode_batch(priors=[B,n]) == update_channel_probs before every shot (the x->z channel update)."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = cfg_codes(1).hz
    n = H.shape[1]
    kw = dict(max_iter=4, bp_method="ms", ms_scaling_factor=0, osd_method="osd0", osd_order=0)
    rng = np.random.default_rng(3)
    B = 64
    probs = rng.uniform(0.02, 0.3, size=(B, n))
    _, syn = random_syndromes(H, 0.1, B, seed=6)
    d = BpOsdDecoder(H, error_rate=0.1, **kw)
    pri = torch.tensor(np.log((1 - probs) / probs), device="cuda", dtype=torch.float64)
    r = d.decode_batch(torch.tensor(syn, device="cuda"), priors=pri)
    o = oracle_mod.OracleDecoder(H, error_rate=0.1, **kw)
    for b in range(B):
        o.update_channel_probs(probs[b])
        o.decode(syn[b])
        assert (r.bp_decoding[b].cpu().numpy() == o.bp_decoding).all()
        assert (r.log_prob_ratios[b].cpu().numpy() == o.log_prob_ratios).all()
        assert bool(r.converge[b]) == o.converge and int(r.iter[b]) == o.iter


def test_edge_cases(torch_cuda, oracle_mod, cfg_codes):
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = cfg_codes(1).hz
    m, n = H.shape
    d = BpOsdDecoder(H, error_rate=0.05, **MS_CS7)
    # empty batch
    r = d.decode_batch(torch.zeros((0, m), dtype=torch.uint8, device="cuda"))
    assert r.osdw_decoding.shape == (0, n)
    # zero syndrome: BP converges in one pass to the zero vector
    r = d.decode_batch(torch.zeros((3, m), dtype=torch.uint8, device="cuda"))
    assert not r.osdw_decoding.any() and r.converge.all() and (r.iter == 1).all()
    # single shot through decode(), int and bool syndromes, wrong length
    _, syn = random_syndromes(H, 0.1, 5, seed=1)
    a = d.decode(syn[0].astype(int))
    b = d.decode(syn[0].astype(bool))
    assert (a == b).all()
    with pytest.raises(ValueError):
        d.decode(np.zeros(m + 1, dtype=int))
    with pytest.raises(ValueError):
        d.decode_batch(torch.zeros((2, m + 1), dtype=torch.uint8, device="cuda"))
    # numpy input goes through the host-buffer entry point and returns numpy
    rn = d.decode_batch(syn)
    rg = d.decode_batch(torch.tensor(syn, device="cuda"))
    assert isinstance(rn.osdw_decoding, np.ndarray)
    assert (rn.osdw_decoding == rg.osdw_decoding.cpu().numpy()).all()
    assert (rn.log_prob_ratios == rg.log_prob_ratios.cpu().numpy()).all()
    # osd_order larger than n - rank is rejected (k = 21 for this code)
    with pytest.raises(ValueError):
        BpOsdDecoder(H, error_rate=0.05, osd_method="osd_cs", osd_order=22)
    # repetition code: rows of weight 2, columns of weight 1-2, rank-deficient-free; ring code is rank deficient
    for Hs in (codes.rep_code(9), codes.ring_code(8), codes.hamming_code(4)):
        kw = dict(max_iter=3, bp_method="ms", ms_scaling_factor=0.5, osd_method="osd_e", osd_order=1)
        _, s2 = random_syndromes(Hs, 0.2, 200, seed=5)
        ref = oracle_mod.OracleDecoder(Hs, error_rate=0.2, **kw).decode_batch(s2)
        for kernel in (0, 1, 2):
            dd, out = gpu_decode(torch, Hs, s2, kernel=kernel, error_rate=0.2, **kw)
            assert_exact(out, ref)


def test_small_workspace_chunks(torch_cuda, oracle_mod, cfg_codes):
    """When LLRs are not requested the failed-shot LLR workspace bounds the chunk size; results must not change."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = cfg_codes(2).hz
    _, syn = random_syndromes(H, 0.07, 3000, seed=10)
    ref = oracle_mod.OracleDecoder(H, error_rate=0.07, **MS_CS7).decode_batch(syn, want_llr=False)
    d = BpOsdDecoder(H, error_rate=0.07, **MS_CS7)
    d.set_tuning(workspace_bytes=1 << 20)  # forces fail_cap = 1024 shots per chunk
    r = d.decode_batch(torch.tensor(syn, device="cuda"), return_llr=False)
    assert d.stats()["chunks"] == 3
    assert (r.osdw_decoding.cpu().numpy() == ref["osdw"]).all()
    assert (r.osd0_decoding.cpu().numpy() == ref["osd0"]).all()
    assert (r.converge.cpu().numpy() == ref["converge"].astype(bool)).all()


def test_sampler_syndrome_and_logical_kernels(torch_cuda, oracle_mod, cfg_codes):
    from bp_osd_b200 import BpOsdDecoder
    code = cfg_codes(2)
    n = code.N
    d = BpOsdDecoder(code.hz, error_rate=0.05, **MS_CS7)
    rng = np.random.default_rng(0)
    pz, px, py = rng.uniform(0, 0.05, n), rng.uniform(0, 0.05, n), rng.uniform(0, 0.03, n)
    d.set_error_channel(pz=pz, px=px, py=py)
    d.set_logicals(code.lz)
    ex_ref, ez_ref = oracle_mod.sample_errors(77, 5000, 500, pz, px, py)
    ex, sx = d.sample_syndromes(77, 5000, 500, sector=0)
    ez, _ = d.sample_syndromes(77, 5000, 500, sector=1)
    assert (ex.cpu().numpy() == ex_ref).all() and (ez.cpu().numpy() == ez_ref).all()
    assert (sx.cpu().numpy() == np.asarray((code.hz @ ex_ref.T) % 2).T).all()
    dec = d.decode_batch(sx).osdw_decoding
    fail = d.logical_check(ex, dec).cpu().numpy()
    want = oracle_mod.logical_fail(code.lz, ex_ref, dec.cpu().numpy()).astype(bool)
    assert (fail == want).all()


def test_sample_and_decode_counters(torch_cuda, oracle_mod, cfg_codes):
    """The device Monte-Carlo step against the same step done with the oracle (css_decode_sim.py:163-205,250-349)."""
    from bp_osd_b200 import BpOsdDecoder
    code = cfg_codes(1)
    n, p, B = code.N, 0.09, 2000
    d = BpOsdDecoder(code.hz, error_rate=p, **MS_CS7)
    d.set_error_channel(px=p)
    d.set_logicals(code.lz)
    c = d.sample_and_decode(seed=99, shot0=0, B=B // 2)
    c = d.sample_and_decode(seed=99, shot0=B // 2, B=B // 2, counters=c)
    z = np.zeros(n)
    ex, _ = oracle_mod.sample_errors(99, 0, B, z, np.full(n, p), z)
    o = oracle_mod.OracleDecoder(code.hz, error_rate=p, **MS_CS7)
    out = o.decode_batch(o.syndrome(ex), want_llr=False)
    fw = oracle_mod.logical_fail(code.lz, ex, out["osdw"]).astype(bool)
    f0 = oracle_mod.logical_fail(code.lz, ex, out["osd0"]).astype(bool)
    fb = oracle_mod.logical_fail(code.lz, ex, out["bp"]).astype(bool)
    cv = out["converge"].astype(bool)
    assert c[0] == B and c[1] == cv.sum() and c[2] == (cv & ~fb).sum()
    assert c[3] == (~f0).sum() and c[4] == (~fw).sum()
    assert c[5] == (~cv).sum() and c[6] == out["iter"].sum()
    wts = np.concatenate([(ex ^ out["osdw"])[fw].sum(1), (ex ^ out["osd0"])[f0].sum(1)])
    assert c[7] == (wts.min() if wts.size else 0)


def test_full_size_round_trip_property(torch_cuda, cfg_codes):
    """At BASELINE size the oracle is too slow; check size-independent properties instead:
    H * osdw == syndrome for every shot, converged shots have osd0 == osdw == bp, weights ordered."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    code = cfg_codes(3)
    d = BpOsdDecoder(code.hz, error_rate=0.05, **MS_CS7)
    d.set_error_channel(px=0.05)
    d.set_logicals(code.lz)
    B = 100_000
    err, syn = d.sample_syndromes(1, 0, B, sector=0)
    r = d.decode_batch(syn, return_llr=False)
    Hd = torch.tensor(code.hz.toarray(), dtype=torch.float16, device="cuda")
    for name in ("osdw_decoding", "osd0_decoding"):
        x = getattr(r, name)
        for s in range(0, B, 20000):
            chk = (x[s:s + 20000].to(torch.float16) @ Hd.T) % 2
            assert (chk.to(torch.uint8) == syn[s:s + 20000]).all(), name
    c = r.converge
    assert (r.osd0_decoding[c] == r.bp_decoding[c]).all() and (r.osdw_decoding[c] == r.bp_decoding[c]).all()
    assert (r.osdw_decoding.sum(1) <= r.osd0_decoding.sum(1)).all()
    st = d.stats()
    assert st["bp_converged"] == int(c.sum()) and st["osd_invocations"] == B - st["bp_converged"]
    fail = d.logical_check(err, r.osdw_decoding)
    assert fail.float().mean().item() < 0.01


def test_fp32_fast_mode_is_statistically_equivalent(torch_cuda, oracle_mod, cfg_codes):
    """fp32 mode: decoding-mismatch rate is reported, the logical error rate must sit inside the 95% CI
    of the fp64 oracle's on the same shots."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    code = cfg_codes(2)
    p, B = 0.06, 6000
    n = code.N
    z = np.zeros(n)
    ex, _ = oracle_mod.sample_errors(4242, 0, B, z, np.full(n, p), z)
    o = oracle_mod.OracleDecoder(code.hz, error_rate=p, **MS_CS7)
    syn = o.syndrome(ex)
    ref = o.decode_batch(syn, want_llr=False)
    ler_ref = oracle_mod.logical_fail(code.lz, ex, ref["osdw"]).mean()
    d = BpOsdDecoder(code.hz, error_rate=p, precision=32, **MS_CS7)
    r = d.decode_batch(torch.tensor(syn, device="cuda"), return_llr=False)
    out = r.osdw_decoding.cpu().numpy()
    assert ((out @ code.hz.toarray().T % 2) == syn).all()
    ler = oracle_mod.logical_fail(code.lz, ex, out).mean()
    ci = 1.96 * np.sqrt(max(ler_ref * (1 - ler_ref), 1e-6) / B)
    mismatch = (out != ref["osdw"]).any(1).mean()
    print(f"fp32: LER {ler:.4f} vs fp64 {ler_ref:.4f} +- {ci:.4f}; decoding mismatch rate {mismatch:.3f}")
    assert abs(ler - ler_ref) <= ci + 2.0 / B


# ------------------------------------------------------------------------------------------------
# Large-H path (BASELINE config 5): HBM-resident OSD-0 kernel, BP with messages outside shared memory
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [2, 4])
@pytest.mark.parametrize("cfg,p,max_iter", [(1, 0.12, 2), (2, 0.08, 4), (3, 0.06, 6)])
def test_hbm_osd0_kernel_matches_shared_memory_kernel_and_oracle(torch_cuda, oracle_mod, cfg_codes, cfg, p, max_iter, variant):
    """Variants 2 and 4 (left-looking panel elimination with the row operations in HBM: one CTA per shot / one cluster per
    shot with TMA-streamed masks) forced on codes that also fit the shared-memory and register kernels."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = cfg_codes(cfg).hz
    kw = dict(max_iter=max_iter, bp_method="ms", ms_scaling_factor=0, osd_method="osd0", osd_order=0)
    _, syn = random_syndromes(H, p, 400, seed=21)
    ref = oracle_mod.OracleDecoder(H, error_rate=p, **kw).decode_batch(syn)
    assert (~ref["converge"].astype(bool)).sum() > 50
    d = BpOsdDecoder(H, error_rate=p, **kw)
    d.set_osd_variant(variant)
    assert d.info()["osd_variant"] == variant
    r = d.decode_batch(torch.tensor(syn, device="cuda"))
    assert (r.osd0_decoding.cpu().numpy() == ref["osd0"]).all()
    assert (r.osdw_decoding.cpu().numpy() == ref["osdw"]).all()
    assert (r.bp_decoding.cpu().numpy() == ref["bp"]).all()
    # the HBM kernel does OSD-0 only
    d2 = BpOsdDecoder(H, error_rate=p, **dict(kw, osd_method="osd_cs", osd_order=3))
    with pytest.raises(NotImplementedError):
        d2.set_osd_variant(variant)
    assert d2.info()["osd_variant"] == 3  # the failed request left the automatic choice (register kernel) in place


def test_large_h_standin_selects_hbm_osd(torch_cuda, oracle_mod):
    """[[3600,144]] HGP of a (3,4)-regular 36x48 seed: T = 1728 x 1728 bits (373 KB) exceeds shared memory,
    so the HBM-resident cluster kernel (variant 4) is chosen automatically and the single-CTA one (variant 2) stays
    selectable; rank-deficient H (redundant checks) included."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    code = hgp(codes.regular_ldpc(36, 48, 3, 4, seed=11), compute_logicals=False)
    H = code.hz
    m, n = H.shape
    assert (m, n) == (1728, 3600)
    kw = dict(max_iter=6, bp_method="ms", ms_scaling_factor=0, osd_method="osd0", osd_order=0)
    _, syn = random_syndromes(H, 0.06, 64, seed=4)
    ref = oracle_mod.OracleDecoder(H, error_rate=0.06, **kw).decode_batch(syn)
    nfail = int((~ref["converge"].astype(bool)).sum())
    assert nfail >= 8
    for kernel, variant in ((None, 4), (0, 4), (3, 4), (None, 2)):   # auto (shared-memory BP), the HBM/L2-scratch BP and the cluster (DSMEM) BP of config 5
        d = BpOsdDecoder(H, error_rate=0.06, **kw)
        if kernel is not None:
            d.set_tuning(bp_kernel=kernel)
        if variant != 4:
            d.set_osd_variant(variant)
        assert d.info()["osd_variant"] == variant
        r = d.decode_batch(torch.tensor(syn, device="cuda"))
        out = dict(osdw=r.osdw_decoding.cpu().numpy(), osd0=r.osd0_decoding.cpu().numpy(), bp=r.bp_decoding.cpu().numpy(),
                   llr=r.log_prob_ratios.cpu().numpy(), converge=r.converge.cpu().numpy(), iter=r.iter.cpu().numpy())
        assert_exact(out, ref)
        assert d.stats()["osd_invocations"] == nfail
    with pytest.raises(NotImplementedError):
        BpOsdDecoder(H, error_rate=0.06, **dict(kw, osd_method="osd_cs", osd_order=2)).decode_batch(
            torch.tensor(syn, device="cuda"))


def test_config5_full_size(torch_cuda, oracle_mod, cfg_codes):
    """BASELINE config 5 at full size (m=19200, n=40000): BP from the HBM/L2 scratch + HBM-resident OSD-0.
    Oracle parity on a handful of shots (its dense elimination needs ~10 s per failed shot) and the
    size-independent property H x = s for every decoded shot."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = cfg_codes(5).hz
    m, n = H.shape
    kw = dict(max_iter=12, bp_method="ms", ms_scaling_factor=0, osd_method="osd0", osd_order=0)
    p = 0.03
    _, syn = random_syndromes(H, p, 24, seed=8)
    d = BpOsdDecoder(H, error_rate=p, **kw)
    info = d.info()
    # messages (1.07 MB in fp64) exceed one SM: the cluster kernel splits them over distributed shared memory
    assert info["bp_kernel"] == 3 and info["bp_cluster_size"] in (8, 16) and info["osd_variant"] == 4
    r = d.decode_batch(torch.tensor(syn, device="cuda"))
    osd0 = r.osd0_decoding.cpu().numpy()
    conv = r.converge.cpu().numpy()
    assert (~conv).sum() >= 2, "the case is meant to exercise OSD"
    resid = (np.asarray(H @ osd0.T) % 2).T
    assert (resid == syn).all(), "H x != s"
    o = oracle_mod.OracleDecoder(H, error_rate=p, **kw)
    pick = list(np.flatnonzero(~conv)[:2]) + list(np.flatnonzero(conv)[:2])
    for b in pick:
        o.decode(syn[b])
        assert bool(conv[b]) == bool(o.converge) and int(r.iter[b]) == o.iter
        assert (r.bp_decoding[b].cpu().numpy() == o.bp_decoding).all()
        assert (r.log_prob_ratios[b].cpu().numpy() == o.log_prob_ratios).all()
        assert (osd0[b] == o.osd0_decoding).all()


def test_per_shot_priors_and_osd_weights(torch_cuda, oracle_mod, cfg_codes):
    """priors + weights per shot == update_channel_probs before every shot, including the OSD-CS weighting (row a16)."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = cfg_codes(2).hz
    n = H.shape[1]
    kw = dict(max_iter=5, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=5)
    rng = np.random.default_rng(5)
    B = 48
    probs = rng.uniform(0.01, 0.2, size=(B, n))
    probs[:, ::7] = 0.0          # exact zeros occur in the x->z update (css_decode_sim.py:218-219)
    _, syn = random_syndromes(H, 0.08, B, seed=9)
    d = BpOsdDecoder(H, error_rate=0.08, **kw)
    with np.errstate(divide="ignore"):
        pri = torch.tensor(np.log((1 - probs) / probs), device="cuda", dtype=torch.float64)
        wts = torch.tensor(np.log(1 / probs), device="cuda", dtype=torch.float64)
    r = d.decode_batch(torch.tensor(syn, device="cuda"), priors=pri, weights=wts)
    o = oracle_mod.OracleDecoder(H, error_rate=0.08, **kw)
    nosd = 0
    for b in range(B):
        o.update_channel_probs(probs[b])
        o.decode(syn[b])
        nosd += 0 if o.converge else 1
        assert (r.bp_decoding[b].cpu().numpy() == o.bp_decoding).all()
        assert (r.osd0_decoding[b].cpu().numpy() == o.osd0_decoding).all()
        assert (r.osdw_decoding[b].cpu().numpy() == o.osdw_decoding).all(), f"shot {b}"
        assert bool(r.converge[b]) == o.converge and int(r.iter[b]) == o.iter
    assert nosd > 10


def test_host_buffer_pipeline(torch_cuda, oracle_mod, cfg_codes):
    """decode_batch(numpy) cuts large batches into double-buffered chunks on two streams; results and
    statistics must equal the single-launch device path."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = cfg_codes(2).hz
    B = 40000  # >= 2 x 16384 shots -> pipelined in 3 chunks
    _, syn = random_syndromes(H, 0.06, B, seed=12)
    d = BpOsdDecoder(H, error_rate=0.06, **MS_CS7)
    rd = d.decode_batch(torch.tensor(syn, device="cuda"))
    sd = d.stats()
    for want_llr in (True, False):
        rh = d.decode_batch(syn, return_llr=want_llr)
        sh = d.stats()
        assert sh["chunks"] == 3 and sh["shots"] == B
        for k in ("bp_converged", "bp_iterations", "osd_invocations"):
            assert sh[k] == sd[k]
        assert (rh.osdw_decoding == rd.osdw_decoding.cpu().numpy()).all()
        assert (rh.osd0_decoding == rd.osd0_decoding.cpu().numpy()).all()
        assert (rh.bp_decoding == rd.bp_decoding.cpu().numpy()).all()
        assert (rh.converge == rd.converge.cpu().numpy()).all()
        assert (rh.iter == rd.iter.cpu().numpy()).all()
        if want_llr:
            assert (rh.log_prob_ratios == rd.log_prob_ratios.cpu().numpy()).all()
    ref = oracle_mod.OracleDecoder(H, error_rate=0.06, **MS_CS7).decode_batch(syn[-2000:], want_llr=False)
    assert (rh.osdw_decoding[-2000:] == ref["osdw"]).all()


@pytest.mark.parametrize("cl", [2, 4, 8, 16])
@pytest.mark.parametrize("precision", [64, 32])
def test_cluster_kernel_sizes(torch_cuda, oracle_mod, cfg_codes, cl, precision):
    """BP kernel 3 (messages split over a thread-block cluster, bit sweep through distributed shared memory)
    forced on a code that also fits one CTA: every cluster size must reproduce the oracle bit for bit in fp64
    and equal the single-CTA kernel in fp32 (same arithmetic, same order)."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = cfg_codes(2).hz
    _, syn = random_syndromes(H, 0.06, 1500, seed=31)
    d = BpOsdDecoder(H, error_rate=0.06, precision=precision, **MS_CS7)
    d.set_tuning(bp_kernel=3)
    try:
        d.set_cluster_size(cl)
    except NotImplementedError:
        pytest.skip(f"cluster size {cl} cannot be scheduled on this device")
    info = d.info()
    assert info["bp_kernel"] == 3 and info["bp_cluster_size"] == cl
    r = d.decode_batch(torch.tensor(syn, device="cuda"))
    out = dict(osdw=r.osdw_decoding.cpu().numpy(), osd0=r.osd0_decoding.cpu().numpy(), bp=r.bp_decoding.cpu().numpy(),
               llr=r.log_prob_ratios.cpu().numpy(), converge=r.converge.cpu().numpy(), iter=r.iter.cpu().numpy())
    if precision == 64:
        ref = oracle_mod.OracleDecoder(H, error_rate=0.06, **MS_CS7).decode_batch(syn)
        assert_exact(out, ref)
    else:
        d2 = BpOsdDecoder(H, error_rate=0.06, precision=32, **MS_CS7)
        d2.set_tuning(bp_kernel=2)
        r2 = d2.decode_batch(torch.tensor(syn, device="cuda"))
        assert (out["llr"] == r2.log_prob_ratios.cpu().numpy()).all()
        assert (out["osdw"] == r2.osdw_decoding.cpu().numpy()).all()
        assert (out["iter"] == r2.iter.cpu().numpy()).all()


def test_cluster_kernel_irregular_and_per_shot_priors(torch_cuda, oracle_mod):
    """Irregular degrees (absent slots), a row count that does not divide the cluster size, non-uniform priors."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = hgp(codes.rep_code(7)).hz          # rows of weight 3-4, columns of weight 1-2; m = 42, n = 85
    m, n = H.shape
    kw = dict(max_iter=9, bp_method="ms", ms_scaling_factor=0.8, osd_method="osd_e", osd_order=4)
    rng = np.random.default_rng(17)
    probs = rng.uniform(0.02, 0.2, size=n)
    _, syn = random_syndromes(H, 0.1, 800, seed=3)
    ref = oracle_mod.OracleDecoder(H, channel_probs=probs, **kw).decode_batch(syn)
    for cl in (2, 4, 8):
        d = BpOsdDecoder(H, channel_probs=probs, **kw)
        d.set_tuning(bp_kernel=3)
        d.set_cluster_size(cl)
        r = d.decode_batch(torch.tensor(syn, device="cuda"))
        out = dict(osdw=r.osdw_decoding.cpu().numpy(), osd0=r.osd0_decoding.cpu().numpy(), bp=r.bp_decoding.cpu().numpy(),
                   llr=r.log_prob_ratios.cpu().numpy(), converge=r.converge.cpu().numpy(), iter=r.iter.cpu().numpy())
        assert_exact(out, ref)


def test_cluster_kernel_rows_of_seven_nonuniform(torch_cuda, oracle_mod, cfg_codes):
    """The cluster kernel's own degree class (rows of exactly 7 slots moved element by element, no padding: what lets a
    cluster of 8 hold config 5) with non-uniform priors -- the prior array in shared memory -- and with a uniform channel
    (no prior array), against the oracle."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = cfg_codes(2).hz
    n = H.shape[1]
    assert int(np.diff(H.tocsr().indptr).max()) == 7
    kw = dict(max_iter=30, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=7)
    rng = np.random.default_rng(5)
    _, syn = random_syndromes(H, 0.06, 600, seed=8)
    for probs in (rng.uniform(0.03, 0.09, size=n), np.full(n, 0.06)):
        ref = oracle_mod.OracleDecoder(H, channel_probs=probs, **kw).decode_batch(syn)
        for cl in (4, 16):
            d = BpOsdDecoder(H, channel_probs=probs, **kw)
            d.set_tuning(bp_kernel=3)
            d.set_cluster_size(cl)
            r = d.decode_batch(torch.tensor(syn, device="cuda"))
            out = dict(osdw=r.osdw_decoding.cpu().numpy(), osd0=r.osd0_decoding.cpu().numpy(), bp=r.bp_decoding.cpu().numpy(),
                       llr=r.log_prob_ratios.cpu().numpy(), converge=r.converge.cpu().numpy(), iter=r.iter.cpu().numpy())
            assert_exact(out, ref)


@pytest.mark.parametrize("variant", [1, 3])
@pytest.mark.parametrize("cfg,p,max_iter,osd_method,osd_order,nonuniform", [
    (1, 0.12, 2, "osd0", 0, False),
    (1, 0.12, 2, "osd_cs", 7, False),
    (1, 0.12, 2, "osd_cs", 21, True),      # order == n - rank, non-uniform channel: ordered fp64 weights
    (1, 0.12, 2, "osd_e", 9, False),
    (1, 0.12, 2, "osd_e", 6, True),
    (2, 0.08, 3, "osd_cs", 7, False),
    (2, 0.08, 3, "osd_cs", 40, False),     # pairs reach into the second kept candidate panel
    (2, 0.08, 3, "osd_cs", 36, True),
    (2, 0.08, 3, "osd_e", 10, False),
    (3, 0.06, 8, "osd_cs", 7, False),
    (3, 0.06, 8, "osd_e", 5, True),
])
def test_osd_kernel_variants(torch_cuda, oracle_mod, cfg_codes, variant, cfg, p, max_iter, osd_method, osd_order, nonuniform):
    """Both shared-memory OSD kernels (3: pivot-block panels, 1: T matrix) against the oracle: OSD-0, OSD-E, OSD-CS,
    uniform and non-uniform channels, rank-deficient H (cfg 3), search depth beyond one panel."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    H = cfg_codes(cfg).hz
    n = H.shape[1]
    kw = dict(max_iter=max_iter, bp_method="ms", ms_scaling_factor=0, osd_method=osd_method, osd_order=osd_order)
    B = 300 if cfg < 3 else 120
    _, syn = random_syndromes(H, p, B, seed=77)
    if nonuniform:
        probs = np.random.default_rng(9).uniform(0.5 * p, 1.5 * p, size=n)
        ref = oracle_mod.OracleDecoder(H, channel_probs=probs, **kw).decode_batch(syn)
        d = BpOsdDecoder(H, channel_probs=probs, **kw)
    else:
        ref = oracle_mod.OracleDecoder(H, error_rate=p, **kw).decode_batch(syn)
        d = BpOsdDecoder(H, error_rate=p, **kw)
    assert (~ref["converge"].astype(bool)).sum() > B // 4
    d.set_osd_variant(variant)
    assert d.info()["osd_variant"] == variant
    r = d.decode_batch(torch.tensor(syn, device="cuda"))
    bad0 = np.flatnonzero((r.osd0_decoding.cpu().numpy() != ref["osd0"]).any(1))
    badw = np.flatnonzero((r.osdw_decoding.cpu().numpy() != ref["osdw"]).any(1))
    assert bad0.size == 0, f"osd0 differs for shots {bad0[:10]}"
    assert badw.size == 0, f"osdw differs for shots {badw[:10]}"


@pytest.mark.parametrize("seed", range(6))
def test_random_parity_check_matrices(torch_cuda, oracle_mod, seed):
    """Random sparse H with irregular degrees (empty rows and columns, weight up to 12 / 8), random channel,
    random decoder settings: every BP kernel and both shared-memory OSD kernels against the oracle."""
    from bp_osd_b200 import BpOsdDecoder
    import scipy.sparse as sp
    torch = torch_cuda
    rng = np.random.default_rng(1000 + seed)
    m, n = int(rng.integers(5, 60)), int(rng.integers(20, 140))
    H = np.zeros((m, n), dtype=np.uint8)
    for j in range(n):                                   # column weights 0..min(m, 6)
        w = int(rng.integers(0, min(m, 6) + 1))
        H[rng.choice(m, size=w, replace=False), j] = 1
    heavy = np.flatnonzero(H.sum(1) > 12)                # keep rows within the specialised kernels' degree class
    for i in heavy:
        ones = np.flatnonzero(H[i])
        H[i, rng.choice(ones, size=ones.size - 12, replace=False)] = 0
    if seed % 2 == 0:
        H[int(rng.integers(0, m))] = 0                   # an empty check
    H = sp.csr_matrix(H)
    probs = rng.uniform(0.01, 0.3, size=n)
    if seed % 3 == 0:
        probs[:] = 0.07
    osd_method, osd_order = [("osd_cs", 5), ("osd_e", 4), ("osd0", 0)][seed % 3]
    kw = dict(max_iter=int(rng.integers(1, 12)), bp_method="ms", ms_scaling_factor=[0, 0.625, 1.0][seed % 3],
              osd_method=osd_method, osd_order=osd_order)
    o = oracle_mod.OracleDecoder(H, channel_probs=probs, **kw)
    if kw["osd_order"] > o.k:
        kw["osd_order"] = o.k
        o = oracle_mod.OracleDecoder(H, channel_probs=probs, **kw)
    _, syn = random_syndromes(H, 0.1, 400, seed=seed)
    ref = o.decode_batch(syn)
    for kernel in (0, 1, 2, 3):
        for variant in (1, 3):
            d = BpOsdDecoder(H, channel_probs=probs, **kw)
            d.set_tuning(bp_kernel=kernel)
            d.set_osd_variant(variant)
            r = d.decode_batch(torch.tensor(syn, device="cuda"))
            out = dict(osdw=r.osdw_decoding.cpu().numpy(), osd0=r.osd0_decoding.cpu().numpy(), bp=r.bp_decoding.cpu().numpy(),
                       llr=r.log_prob_ratios.cpu().numpy(), converge=r.converge.cpu().numpy(), iter=r.iter.cpu().numpy())
            try:
                assert_exact(out, ref)
            except AssertionError as ex:
                raise AssertionError(f"kernel {kernel} osd variant {variant} m={m} n={n}: {ex}")


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("cfg,p,kw", [
    (1, 0.10, dict(max_iter=3, bp_method="ms", ms_scaling_factor=0.625, osd_method="osd_e", osd_order=6)),
    (2, 0.06, dict(MS_CS7, max_iter=20)),
    (3, 0.05, MS_CS7),
    (3, 0.06, dict(MS_CS7, max_iter=30)),
])
def test_latency_path_small_host_batches(torch_cuda, oracle_mod, cfg_codes, cfg, p, kw, precision):
    """Host batches of at most one shot per SM take the latency path of bposd_decode_host (one launch, results written
    straight to pinned host memory, latency geometry): same bits as the oracle, shot by shot
    through decode() and as small batches, including shots that need the OSD stage."""
    from bp_osd_b200 import BpOsdDecoder
    H = cfg_codes(cfg).hz
    _, syn = random_syndromes(H, p, 48, seed=300 + cfg)
    ref = oracle_mod.OracleDecoder(H, error_rate=p, **kw).decode_batch(syn)
    d = BpOsdDecoder(H, error_rate=p, precision=precision, **kw)
    if precision == 64:
        assert (ref["converge"] == 0).any() or cfg == 3, "the case should exercise the OSD stage"
    for B in (1, 5, 48):
        r = d.decode_batch(syn[:B])
        st = d.stats()
        assert st["shots"] == B and st["chunks"] == 1
        assert H.shape[1] > 2048 or d.info()["bp_kernel"] != 2 or st["launches"] <= 2   # BP (+ OSD), no copies, no sampler
        if precision == 64:
            out = dict(osdw=r.osdw_decoding, osd0=r.osd0_decoding, bp=r.bp_decoding, llr=r.log_prob_ratios,
                       converge=r.converge, iter=r.iter)
            assert_exact(out, {k: v[:B] for k, v in ref.items()})
            assert st["bp_converged"] == int(ref["converge"][:B].sum())
            assert st["bp_iterations"] == int(ref["iter"][:B].sum())
            assert st["osd_invocations"] == int((ref["converge"][:B] == 0).sum())
        else:
            assert not ((H @ r.osdw_decoding.T) % 2 != syn[:B].T).any(), "fp32 osdw must still satisfy the syndrome"
    for i in range(12):
        s = syn[i].astype(np.int64)
        out = d.decode(s)
        assert out.dtype == np.int64 and out is d.osdw_decoding
        if precision == 64:
            assert (out == ref["osdw"][i]).all() and (d.osd0_decoding == ref["osd0"][i]).all()
            assert (d.bp_decoding == ref["bp"][i]).all() and d.bp_decoding.dtype == np.int64
            assert (d.log_prob_ratios == ref["llr"][i]).all()
            assert d.converge == bool(ref["converge"][i]) and d.iter == int(ref["iter"][i])
        else:
            assert not ((H @ out) % 2 != syn[i]).any()


@pytest.mark.parametrize("precision", [64, 32])
def test_bit_packed_io_matches_byte_io(torch_cuda, oracle_mod, cfg_codes, precision):
    """`packed=True` (bit-packed syndromes in, bit-packed decodings out: bposd_decode_batch_packed /
    bposd_decode_host_packed / bposd_sample_syndromes_packed) gives exactly the bits of the byte-per-bit interface, on the
    device path, the chunked host pipeline, the small-batch staging path and the single-shot latency path."""
    from bp_osd_b200 import BpOsdDecoder, pack_bits, unpack_bits
    torch = torch_cuda
    code = cfg_codes(2)
    H = code.hz
    m, n = H.shape
    kw = dict(max_iter=6, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=5)
    d = BpOsdDecoder(H, error_rate=0.07, precision=precision, **kw)
    d.set_error_channel(px=0.07)
    # sampler: packed syndromes are the packed form of the byte syndromes of the same shots
    e1, s_bytes = d.sample_syndromes(11, 1000, 40000)
    e2, s_bits = d.sample_syndromes(11, 1000, 40000, packed=True)
    assert s_bits.shape == (40000, (m + 7) // 8)
    assert (e1 == e2).all() and (unpack_bits(s_bits, m) == s_bytes).all() and (pack_bits(s_bytes) == s_bits).all()
    # device path
    ref = d.decode_batch(s_bytes)
    st_ref = d.stats()
    got = d.decode_batch(s_bits, packed=True)
    assert got.packed and got.osdw_decoding.shape == (40000, (n + 7) // 8)
    for k in ("osdw_decoding", "osd0_decoding", "bp_decoding"):
        assert (unpack_bits(getattr(got, k), n) == getattr(ref, k)).all(), k
    assert (got.converge == ref.converge).all() and (got.iter == ref.iter).all()
    assert torch.equal(got.log_prob_ratios, ref.log_prob_ratios)
    assert d.stats()["osd_invocations"] == st_ref["osd_invocations"] > 1000
    if precision == 64:
        o = oracle_mod.OracleDecoder(H, error_rate=0.07, **kw).decode_batch(s_bytes[:3000].cpu().numpy())
        assert (unpack_bits(got.osdw_decoding[:3000], n).cpu().numpy() == o["osdw"]).all()
    # host paths: 40 000 shots take the double-buffered pipeline, 2 000 the staging block, 5 the latency kernel
    hb, hp = s_bytes.cpu().numpy(), s_bits.cpu().numpy()
    for B in (40000, 2000, 5, 1):
        r = d.decode_batch(hp[:B], packed=True)
        for k in ("osdw_decoding", "osd0_decoding", "bp_decoding"):
            assert (unpack_bits(getattr(r, k), n) == getattr(ref, k)[:B].cpu().numpy()).all(), (B, k)
        assert (r.converge == ref.converge[:B].cpu().numpy()).all() and (r.iter == ref.iter[:B].cpu().numpy()).all()
        assert np.array_equal(r.log_prob_ratios, ref.log_prob_ratios[:B].cpu().numpy())
        r2 = d.decode_batch(hp[:B], packed=True, return_llr=False, return_all=False)
        assert r2.osd0_decoding is None and (r2.osdw_decoding == r.osdw_decoding).all()
    with pytest.raises(ValueError):
        d.decode_batch(hb[:10], packed=True)       # byte syndromes are not [B, ceil(m/8)]
    with pytest.raises(ValueError):
        d.decode_batch(hb[:10], priors=ref.log_prob_ratios[:10])   # per-shot priors need CUDA syndromes


def test_tensor_arguments_are_validated(torch_cuda, cfg_codes):
    """logical_check / channel_update hand raw pointers to kernels that read B*n bytes: wrong dtype, shape or residency is
    a ValueError, never an out-of-bounds read."""
    from bp_osd_b200 import BpOsdDecoder
    torch = torch_cuda
    code = cfg_codes(1)
    d = BpOsdDecoder(code.hz, error_rate=0.05, osd_method="osd0")
    d.set_logicals(code.lz)
    n = code.hz.shape[1]
    good = torch.zeros((4, n), dtype=torch.uint8, device="cuda")
    assert not d.logical_check(good, good).any()
    for bad in (good.to(torch.int64), good.cpu(), good[:, :-1], good[0]):
        with pytest.raises(ValueError):
            d.logical_check(bad, good)
        with pytest.raises(ValueError):
            d.channel_update(bad, np.full(n, 0.1), np.full(n, 0.2))
    with pytest.raises(ValueError):
        d.logical_check(good, good[:2])
