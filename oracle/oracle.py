"""ctypes front-end of the C oracle (oracle/bposd_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  It exposes the same surface the reference uses on
``bposd_decoder`` (/root/reference/README.md:178-204; src/bposd/css_decode_sim.py:444-463,
174-202, 217-258): constructor keywords, ``decode``, ``osdw_decoding``,
``osd0_decoding``, ``bp_decoding``, ``log_prob_ratios``, ``converge``, ``iter``,
``update_channel_probs``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbposd_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "bposd_oracle.c"), os.path.join(_HERE, "..", "include", "bposd_math.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libbposd_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        P = C.c_void_p
        L.oracle_create.restype = P
        L.oracle_create.argtypes = [P, P, C.c_int, C.c_int, P, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
        L.oracle_destroy.argtypes = [P]
        L.oracle_decode.argtypes = [P, P]
        L.oracle_decode_batch.argtypes = [P, P, C.c_long, P, P, P, P, P, P]
        L.oracle_update_channel_probs.argtypes = [P, P]
        for name in ("rank", "k", "converge", "iter", "osd_ran"):
            f = getattr(L, "oracle_" + name)
            f.restype, f.argtypes = C.c_int, [P]
        for name in ("stat_elim_wordxors", "total_elim_wordxors", "total_osd"):
            f = getattr(L, "oracle_" + name)
            f.restype, f.argtypes = C.c_long, [P]
        for name in ("llr", "bp_decoding", "osd0_decoding", "osdw_decoding"):
            f = getattr(L, "oracle_" + name)
            f.restype, f.argtypes = P, [P]
        L.oracle_set_math.argtypes = [P, C.c_int]
        L.oracle_set_schedule.restype, L.oracle_set_schedule.argtypes = None, [P, C.c_int, P]
        for name in ("tanh", "log", "expm1"):
            f = getattr(L, "oracle_math_" + name)
            f.restype, f.argtypes = C.c_double, [C.c_double]
        L.oracle_math_map.restype, L.oracle_math_map.argtypes = None, [C.c_int, P, P, P, C.c_longlong]
        L.oracle_philox.argtypes = [P, P, P]
        L.oracle_sample_errors.argtypes = [C.c_uint64, C.c_uint64, C.c_long, C.c_int, P, P, P, P, P]
        L.oracle_syndrome.argtypes = [P, P, C.c_long, P]
        L.oracle_logical_fail.argtypes = [P, P, C.c_int, C.c_int, P, P, C.c_long, P]
        _lib = L
    return _lib


def math_map(fn, a, b=None):
    """Host side of include/bposd_math.h element-wise: fn "div" (a / b), "tanh", "log", "ratio" ((1 + a) / (1 - a))."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    out = np.empty_like(a)
    lib().oracle_math_map({"div": 0, "tanh": 1, "log": 2, "ratio": 3}[fn], _ptr(a), _ptr(b), _ptr(out), a.size)
    return out


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_BP = {"ps": 0, "product_sum": 0, "prod_sum": 0, "0": 0, 0: 0,
       "ms": 1, "minimum_sum": 1, "min_sum": 1, "1": 1, 1: 1}
_OSD = {"osd0": 0, "osd_0": 0, "0": 0, "osd_e": 1, "osde": 1, "exhaustive": 1,
        "osd_cs": 2, "osdcs": 2, "combination_sweep": 2}


def csr_of(h):
    h = sp.csr_matrix(h).astype(np.uint8)
    h.data %= 2
    h.eliminate_zeros()
    h.sort_indices()
    return h


class OracleDecoder:
    def __init__(self, parity_check_matrix, error_rate=None, channel_probs=None, max_iter=0,
                 bp_method="ms", ms_scaling_factor=1.0, osd_method="osd0", osd_order=0, math="shared",
                 schedule="parallel", serial_schedule_order=None):
        """math: "shared" = the portable tanh/log of include/bposd_math.h (the functions the CUDA kernels use, so
        product-sum compares bit for bit), "libm" = the host libm, as ldpc itself calls (last bits machine dependent)."""
        h = csr_of(parity_check_matrix)
        self.m, self.n = h.shape
        cp = None
        if channel_probs is not None and len(channel_probs) and channel_probs[0] is not None:
            cp = np.ascontiguousarray(channel_probs, dtype=np.float64)
            if cp.shape != (self.n,):
                raise ValueError("channel_probs length must equal the number of columns")
        elif error_rate is not None:
            cp = np.full(self.n, float(error_rate))
        else:
            raise ValueError("either error_rate or channel_probs is required")
        self._ip = np.ascontiguousarray(h.indptr, dtype=np.int32)
        self._ix = np.ascontiguousarray(h.indices, dtype=np.int32)
        bm = _BP[bp_method.lower() if isinstance(bp_method, str) else bp_method]
        om = _OSD[str(osd_method).lower()]
        self._h = lib().oracle_create(_ptr(self._ip), _ptr(self._ix), self.m, self.n, _ptr(cp),
                                      int(max_iter), bm, float(ms_scaling_factor), om, int(osd_order))
        if math not in ("shared", "libm"):
            raise ValueError("math must be 'shared' or 'libm'")
        lib().oracle_set_math(self._h, 1 if math == "libm" else 0)
        if schedule not in ("parallel", "serial"):
            raise ValueError("schedule must be 'parallel' or 'serial'")
        if schedule == "serial":   # row f4: an ldpc option the reference never passes
            o = None if serial_schedule_order is None else np.ascontiguousarray(serial_schedule_order, dtype=np.int32)
            assert o is None or sorted(o.tolist()) == list(range(self.n))
            lib().oracle_set_schedule(self._h, 1, _ptr(o))
        self.rank = lib().oracle_rank(self._h)
        self.k = lib().oracle_k(self._h)
        if om != 0 and int(osd_order) > self.k:
            raise ValueError("osd_order must not exceed n - rank")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def update_channel_probs(self, probs):
        p = np.ascontiguousarray(probs, dtype=np.float64)
        assert p.shape == (self.n,)
        lib().oracle_update_channel_probs(self._h, _ptr(p))

    def decode(self, syndrome):
        s = np.ascontiguousarray(np.asarray(syndrome).astype(np.int64) & 1, dtype=np.uint8)
        if s.shape != (self.m,):
            raise ValueError("syndrome length must equal the number of rows")
        lib().oracle_decode(self._h, _ptr(s))
        return self.osdw_decoding

    def decode_batch(self, syndromes, want_llr=True):
        s = np.ascontiguousarray(np.asarray(syndromes).astype(np.uint8) & 1)
        B = s.shape[0]
        out = {
            "osdw": np.zeros((B, self.n), np.uint8), "osd0": np.zeros((B, self.n), np.uint8),
            "bp": np.zeros((B, self.n), np.uint8),
            "llr": np.zeros((B, self.n), np.float64) if want_llr else None,
            "converge": np.zeros(B, np.uint8), "iter": np.zeros(B, np.int32),
        }
        lib().oracle_decode_batch(self._h, _ptr(s), B, _ptr(out["osdw"]), _ptr(out["osd0"]), _ptr(out["bp"]),
                                  _ptr(out["llr"]), _ptr(out["converge"]), _ptr(out["iter"]))
        return out

    def _vec(self, fn, dtype):
        p = fn(self._h)
        ct = C.c_double if dtype == np.float64 else C.c_uint8
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(self.n,)).copy()

    @property
    def osdw_decoding(self):
        return self._vec(lib().oracle_osdw_decoding, np.uint8).astype(int)

    @property
    def osd0_decoding(self):
        return self._vec(lib().oracle_osd0_decoding, np.uint8).astype(int)

    @property
    def bp_decoding(self):
        return self._vec(lib().oracle_bp_decoding, np.uint8).astype(int)

    @property
    def log_prob_ratios(self):
        return self._vec(lib().oracle_llr, np.float64)

    @property
    def converge(self):
        return bool(lib().oracle_converge(self._h))

    @property
    def iter(self):
        return int(lib().oracle_iter(self._h))

    @property
    def osd_ran(self):
        return bool(lib().oracle_osd_ran(self._h))

    @property
    def elim_wordxors(self):
        return int(lib().oracle_stat_elim_wordxors(self._h))

    @property
    def totals(self):
        """(32-bit word XORs spent in GF(2) elimination, OSD invocations) accumulated over all decodes."""
        return int(lib().oracle_total_elim_wordxors(self._h)), int(lib().oracle_total_osd(self._h))

    def syndrome(self, errors):
        e = np.ascontiguousarray(errors, dtype=np.uint8).reshape(-1, self.n)
        out = np.zeros((e.shape[0], self.m), np.uint8)
        lib().oracle_syndrome(self._h, _ptr(e), e.shape[0], _ptr(out))
        return out


def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, np.uint32)
    lib().oracle_philox(_ptr(c), _ptr(k), _ptr(out))
    return out


def prob_threshold(p):
    """u32 threshold T with P(r < T) = floor(p * 2^32) / 2^32 (clipped)."""
    t = np.floor(np.asarray(p, dtype=np.float64) * 4294967296.0)
    return np.clip(t, 0, 4294967295).astype(np.uint32)


def sample_errors(seed, shot0, B, pz, px, py):
    pz, px, py = (np.asarray(a, dtype=np.float64) for a in (pz, px, py))
    n = pz.size
    t1, t2, t3 = prob_threshold(pz), prob_threshold(pz + px), prob_threshold(pz + px + py)
    ex = np.zeros((B, n), np.uint8)
    ez = np.zeros((B, n), np.uint8)
    lib().oracle_sample_errors(int(seed), int(shot0), B, n, _ptr(t1), _ptr(t2), _ptr(t3), _ptr(ex), _ptr(ez))
    return ex, ez


def logical_fail(logicals, errors, decodings):
    l = csr_of(logicals)
    ip = np.ascontiguousarray(l.indptr, dtype=np.int32)
    ix = np.ascontiguousarray(l.indices, dtype=np.int32)
    e = np.ascontiguousarray(errors, dtype=np.uint8)
    d = np.ascontiguousarray(decodings, dtype=np.uint8)
    B, n = e.shape
    out = np.zeros(B, np.uint8)
    lib().oracle_logical_fail(_ptr(ip), _ptr(ix), l.shape[0], n, _ptr(e), _ptr(d), B, _ptr(out))
    return out
