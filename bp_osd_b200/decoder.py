"""``bposd_decoder`` / ``BpOsdDecoder``: the reference's decoder class, backed by the sm_100a kernels.

Drop-in surface.  Constructor keywords and their meaning follow the
reference's two call styles -- README.md:178-187 (``error_rate`` + ``channel_probs=[None]``) and
src/bposd/css_decode_sim.py:444-463 (``channel_probs`` only) of /root/reference -- and the result
attributes are the ones the reference reads after ``decode``: ``osdw_decoding``,
``osd0_decoding``, ``bp_decoding`` (css_decode_sim.py:257,294,338; README.md:202), ``converge``
(css_decode_sim.py:331-336), plus ``log_prob_ratios`` and ``iter``.  ``update_channel_probs``
is css_decode_sim.py:229,248.  New: ``decode_batch(syndromes[B, m])`` for CUDA tensors (or numpy
arrays), and the device-side Monte-Carlo step ``sample_and_decode``.

All arithmetic runs in libbposd_b200.so through the C ABI (include/bposd_b200.h); this class
only validates arguments, owns buffers and converts types.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import scipy.sparse as sp

from . import _capi

__all__ = ["BpOsdDecoder", "bposd_decoder", "BatchResult", "prob_thresholds", "pack_bits", "unpack_bits"]

_BP_NAMES = {
    "ps": _capi.BP_PRODUCT_SUM, "product_sum": _capi.BP_PRODUCT_SUM, "prod_sum": _capi.BP_PRODUCT_SUM,
    "0": _capi.BP_PRODUCT_SUM, 0: _capi.BP_PRODUCT_SUM,
    # the pre-v2 "log-domain" spellings: ldpc v2 runs every BP variant in the log domain, so they name the same two updates
    "ps_log": _capi.BP_PRODUCT_SUM, "psl": _capi.BP_PRODUCT_SUM, "product_sum_log": _capi.BP_PRODUCT_SUM,
    "ms_log": _capi.BP_MINIMUM_SUM, "msl": _capi.BP_MINIMUM_SUM, "minimum_sum_log": _capi.BP_MINIMUM_SUM,
    "ms": _capi.BP_MINIMUM_SUM, "minimum_sum": _capi.BP_MINIMUM_SUM, "min_sum": _capi.BP_MINIMUM_SUM,
    "1": _capi.BP_MINIMUM_SUM, 1: _capi.BP_MINIMUM_SUM,
}
_OSD_NAMES = {
    "osd0": _capi.OSD_0, "osd_0": _capi.OSD_0, "0": _capi.OSD_0, "zero": _capi.OSD_0,
    "osd_e": _capi.OSD_E, "osde": _capi.OSD_E, "e": _capi.OSD_E, "exhaustive": _capi.OSD_E,
    "osd_cs": _capi.OSD_CS, "osdcs": _capi.OSD_CS, "cs": _capi.OSD_CS, "combination_sweep": _capi.OSD_CS,
    "off": _capi.OSD_OFF, "osd_off": _capi.OSD_OFF, "none": _capi.OSD_OFF,
}


def prob_thresholds(p) -> np.ndarray:
    """uint32 threshold T such that a 32-bit uniform r satisfies r < T with probability floor(p 2^32)/2^32."""
    t = np.floor(np.asarray(p, dtype=np.float64) * 4294967296.0)
    return np.clip(t, 0, 4294967295).astype(np.uint32)


def pack_bits(a):
    """[B, k] 0/1 array (numpy or torch) -> [B, ceil(k/8)] uint8, bit i%8 of byte i/8 = entry i: the layout of the
    ``packed=True`` interfaces (``numpy.packbits(..., bitorder="little")``)."""
    try:
        import torch
        if isinstance(a, torch.Tensor):
            B, k = a.shape
            pad = (-k) % 8
            x = (a & 1).to(torch.uint8) if a.dtype != torch.bool else a.to(torch.uint8)
            if pad:
                x = torch.nn.functional.pad(x, (0, pad))
            w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int32, device=a.device)
            return (x.view(B, -1, 8).to(torch.int32) * w).sum(-1).to(torch.uint8)
    except ImportError:  # pragma: no cover
        pass
    return np.packbits(np.asarray(a).astype(np.uint8) & 1, axis=1, bitorder="little")


def unpack_bits(a, k):
    """Inverse of :func:`pack_bits`: [B, ceil(k/8)] uint8 -> [B, k] uint8 of 0/1."""
    try:
        import torch
        if isinstance(a, torch.Tensor):
            sh = torch.arange(8, dtype=torch.uint8, device=a.device)
            return ((a.unsqueeze(-1) >> sh) & 1).reshape(a.shape[0], -1)[:, :k].contiguous()
    except ImportError:  # pragma: no cover
        pass
    return np.unpackbits(np.asarray(a, dtype=np.uint8), axis=1, bitorder="little")[:, :k]


@dataclass
class BatchResult:
    """Outputs of ``decode_batch``; tensors live where the syndromes lived (CUDA or host).  With ``packed`` the three
    decodings are bit-packed rows of ceil(n/8) bytes (see :func:`unpack_bits`)."""
    osdw_decoding: object
    osd0_decoding: object
    bp_decoding: object
    log_prob_ratios: object
    converge: object
    iter: object
    packed: bool = False


def _ptr(arr) -> Optional[int]:
    return None if arr is None else arr.ctypes.data


class BpOsdDecoder:
    def __init__(self, parity_check_matrix, error_rate=None, channel_probs=None, max_iter=0,
                 bp_method="minimum_sum", ms_scaling_factor=1.0, osd_method="osd_0", osd_order=0,
                 error_channel=None, schedule="parallel", input_vector_type="syndrome",
                 precision=64, device=None, **ignored):
        # keywords of ldpc's BpOsdDecoder that have no effect on the results are accepted and ignored; anything else is a
        # spelling mistake and must not be swallowed
        unknown = set(ignored) - {"omp_thread_count", "random_schedule_seed", "serial_schedule_order", "random_serial_schedule"}
        if unknown:
            raise TypeError(f"unexpected keyword argument(s): {sorted(unknown)}")
        # BP schedule (ldpc v2 option, not reachable from the reference): 'parallel' (flooding, the default) or 'serial'
        # (bit after bit, in `serial_schedule_order` or 0 .. n-1).  The randomised serial schedule reshuffles the order
        # with the C++ standard library's generator every iteration; that stream is not reproducible here and is refused.
        sched = {"parallel": 0, 0: 0, "0": 0, "serial": 1, 1: 1, "1": 1}.get(schedule.lower() if isinstance(schedule, str) else schedule)
        if sched is None:
            raise ValueError("schedule must be 'parallel' or 'serial'")
        if ignored.get("random_serial_schedule") not in (None, False, 0) or ignored.get("random_schedule_seed") not in (None, False, 0, -1):
            raise ValueError("the randomised serial schedule (random_serial_schedule / random_schedule_seed) is not implemented")
        order = ignored.get("serial_schedule_order")
        self.schedule = "serial" if sched else "parallel"
        ivt = {"syndrome": 0, 0: 0, "0": 0, "received_vector": 1, 1: 1, "1": 1, "auto": 2, 2: 2, "2": 2}.get(
            input_vector_type.lower() if isinstance(input_vector_type, str) else input_vector_type)
        if ivt is None:
            raise ValueError("input_vector_type must be 'syndrome', 'received_vector' or 'auto'")
        # 'received_vector' (ldpc v2 option, not reachable from the reference): decode(r) decodes the syndrome H r and returns
        # r + decoding, the corrected word; 'auto' picks by the length of the input (ambiguous when m == n)
        self._input_vector_type = ivt
        h = parity_check_matrix
        if not sp.issparse(h):
            h = np.asarray(h)
            if h.ndim != 2:
                raise TypeError("parity_check_matrix must be a 2-D array or a scipy sparse matrix")
        h = sp.csr_matrix(h)
        h.data = (np.asarray(h.data).astype(np.int64) % 2).astype(np.uint8)
        h.eliminate_zeros()
        h.sum_duplicates()
        h.data %= 2
        h.eliminate_zeros()
        h.sort_indices()
        self.m, self.n = (int(x) for x in h.shape)
        self._pcm = h.astype(np.uint8)

        probs = error_channel if error_channel is not None else channel_probs
        if probs is not None and len(probs) and probs[0] is not None:
            probs = np.array(probs, dtype=np.float64)  # copy: the caller's array may be read-only
            if probs.shape != (self.n,):
                raise ValueError(f"channel probability vector must have length {self.n}")
        elif error_rate is not None:
            probs = np.full(self.n, float(error_rate), dtype=np.float64)
        else:
            raise ValueError("specify either error_rate or channel_probs")
        if np.any(probs < 0) or np.any(probs > 1) or np.any(np.isnan(probs)):
            raise ValueError("channel probabilities must lie in [0, 1]")
        self._probs = probs

        key = bp_method.lower() if isinstance(bp_method, str) else bp_method
        if key not in _BP_NAMES:
            raise ValueError(f"bp_method '{bp_method}' is invalid; use 'ms'/'minimum_sum' or 'ps'/'product_sum'")
        okey = str(osd_method).lower()
        if okey not in _OSD_NAMES:
            raise ValueError(f"osd_method '{osd_method}' is invalid; use 'osd0', 'osd_e' or 'osd_cs'")
        if precision not in (64, 32):
            raise ValueError("precision must be 64 (bit-exact mode) or 32 (fast mode)")
        if int(max_iter) < 0:
            raise ValueError("max_iter must be non-negative")
        if int(osd_order) < 0:
            raise ValueError("osd_order must be non-negative")
        self.precision = int(precision)
        self._real = np.float64 if self.precision == 64 else np.float32

        if device is None:
            device = 0
            try:
                import torch
                if torch.cuda.is_available():
                    device = torch.cuda.current_device()
            except Exception:
                pass
        self.device = int(device)

        lib = _capi.load()
        ip = np.ascontiguousarray(h.indptr, dtype=np.int32)
        ix = np.ascontiguousarray(h.indices, dtype=np.int32)
        handle = C.c_void_p()
        rc = lib.bposd_create(_ptr(ip), _ptr(ix), self.m, self.n, _ptr(self._probs), int(max_iter),
                              _BP_NAMES[key], float(ms_scaling_factor), _OSD_NAMES[okey], int(osd_order),
                              self.precision, self.device, C.byref(handle))
        _capi.check(None, rc)
        self._h = handle
        if sched:
            o = None
            if order is not None:
                o = np.ascontiguousarray(order, dtype=np.int32)
                if o.shape != (self.n,):
                    raise ValueError(f"serial_schedule_order must have length {self.n}")
            self._check(lib.bposd_set_schedule(self._h, 1, _ptr(o)))
        info = self.info()
        self.rank, self.k = info["rank"], info["k"]
        self.max_iter = info["max_iter"]
        self.bp_method = "minimum_sum" if info["bp_method"] == 1 else "product_sum"
        self.osd_method = {0: "osd_0", 1: "osd_e", 2: "osd_cs", 3: "off"}[info["osd_method"]]
        self.osd_order = info["osd_order"]
        self.ms_scaling_factor = float(ms_scaling_factor)
        # result attributes, overwritten by every decode (reference semantics)
        self.osdw_decoding = np.zeros(self.n, dtype=int)
        self._lazy = None
        self.osd0_decoding = np.zeros(self.n, dtype=int)
        self.bp_decoding = np.zeros(self.n, dtype=int)
        self.log_prob_ratios = np.zeros(self.n, dtype=np.float64)
        self.converge = False
        self.iter = 0
        self._one = None

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _capi.load().bposd_destroy(h)
            except Exception:
                pass
            self._h = None

    def _check(self, rc):
        _capi.check(self._h, rc)

    def info(self) -> dict:
        inf = _capi.Info()
        self._check(_capi.load().bposd_get_info(self._h, C.byref(inf)))
        return {k: getattr(inf, k) for k, _ in _capi.Info._fields_}

    def stats(self) -> dict:
        st = _capi.Stats()
        self._check(_capi.load().bposd_get_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in _capi.Stats._fields_}

    def set_tuning(self, bp_kernel=None, bp_threads=0, workspace_bytes=0):
        """bp_kernel: None (auto), 0 generic/global, 1 generic/smem, 2 in-place smem, 3 cluster DSMEM."""
        sel = 0 if bp_kernel is None else int(bp_kernel) + 1
        self._check(_capi.load().bposd_set_tuning(self._h, sel, int(bp_threads), int(workspace_bytes)))

    def int32_peak(self) -> float:
        """Measured LOP3 rate of the device in ops/s (denominator of the OSD roofline)."""
        v = C.c_double()
        self._check(_capi.load().bposd_int32_peak(self._h, C.byref(v)))
        return float(v.value)

    def smem_peak(self) -> float:
        """Measured shared-memory bandwidth of the device in bytes/s (denominator of the BP roofline)."""
        v = C.c_double()
        self._check(_capi.load().bposd_smem_peak(self._h, C.byref(v)))
        return float(v.value)

    def fp64_peak(self) -> float:
        """Measured fp64 fused-multiply-add rate of the device in DFMA/s (denominator of the product-sum roofline)."""
        v = C.c_double()
        self._check(_capi.load().bposd_fp64_peak(self._h, C.byref(v)))
        return float(v.value)

    def math_probe(self, fn: str, a, b=None):
        """Test hook: the device side of include/bposd_math.h on CUDA float64 tensors.  fn: "div" (a / b by the in-range
        division sequence), "tanh", "log", "ratio" ((1 + a) / (1 - a) as the product-sum update forms it)."""
        import torch
        code = {"div": 0, "tanh": 1, "log": 2, "ratio": 3}[fn]
        for t in (a, b):
            if t is not None and not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64
                                      and t.is_contiguous() and t.device.index == self.device):
                raise ValueError("math_probe takes contiguous float64 tensors on the decoder's device")
        if code == 0 and (b is None or b.numel() != a.numel()):
            raise ValueError("div needs b with as many elements as a")
        out = torch.empty_like(a)
        self._check(_capi.load().bposd_math_probe(self._h, code, a.data_ptr(), b.data_ptr() if b is not None else None,
                                                  out.data_ptr(), a.numel()))
        return out

    def set_cluster_size(self, cluster_size=0):
        """Thread-block-cluster size of BP kernel 3 (0 = smallest that fits, else 2, 4, 8 or 16)."""
        self._check(_capi.load().bposd_set_cluster_size(self._h, int(cluster_size)))

    def set_osd_variant(self, variant=None, workspace_bytes=0):
        """variant: None (auto), 3 register kernel, 1 shared-memory T-matrix kernel, 4 / 2 HBM-resident OSD-0 kernels for
        large H (one thread-block cluster per failed shot with TMA-streamed masks / one CTA per failed shot)."""
        self._check(_capi.load().bposd_set_osd_variant(self._h, 0 if variant is None else int(variant),
                                                       int(workspace_bytes)))

    @property
    def channel_probs(self):
        return self._probs.copy()

    @property
    def decoding(self):
        return self.osdw_decoding

    def update_channel_probs(self, channel_probs):
        p = np.array(channel_probs, dtype=np.float64)
        if p.shape != (self.n,):
            raise ValueError(f"channel probability vector must have length {self.n}")
        if np.any(p < 0) or np.any(p > 1) or np.any(np.isnan(p)):
            raise ValueError("channel probabilities must lie in [0, 1]")
        self._probs = p
        self._check(_capi.load().bposd_update_channel_probs(self._h, _ptr(p)))

    # ------------------------------------------------------------------ decode
    def decode(self, syndrome):
        """Decode one syndrome; returns ``osdw_decoding`` and refreshes the result attributes.

        The call is one ``bposd_decode_host`` with B = 1 (the library's latency path: one kernel launch, results
        written straight to pinned host memory).  ``osd0_decoding``, ``bp_decoding`` and ``log_prob_ratios`` are
        converted to the caller's dtype when they are first read, not here."""
        s = syndrome if type(syndrome) is np.ndarray else np.asarray(syndrome)
        if self._input_vector_type != 0:
            if s.ndim != 1:
                raise ValueError("input vector must be one-dimensional")
            received = self._input_vector_type == 1
            if self._input_vector_type == 2:
                if self.m == self.n:
                    raise ValueError("input_vector_type='auto' is ambiguous for a square parity-check matrix")
                received = s.shape[0] == self.n
            if received:
                if s.shape[0] != self.n:
                    raise ValueError(f"received vector must have length {self.n}")
                r = (s.astype(np.int64) & 1).astype(np.uint8)
                self._input_vector_type, keep = 0, self._input_vector_type
                try:
                    self.decode(np.asarray(self._pcm @ r) % 2)
                finally:
                    self._input_vector_type = keep
                self.osdw_decoding = (self.osdw_decoding ^ r.astype(self.osdw_decoding.dtype))
                self._osd0 = self.osd0_decoding ^ r.astype(self.osdw_decoding.dtype)
                self._bp = self.bp_decoding ^ r.astype(self.osdw_decoding.dtype)
                return self.osdw_decoding
        if s.ndim != 1 or s.shape[0] != self.m:
            raise ValueError(f"syndrome must have length {self.m}")
        kind = s.dtype.kind
        dtype = s.dtype if kind in "iu" else int
        one = self._one
        if one is None:   # result buffers (and their raw addresses) of the single-shot path are set up once
            one = self._one = {"osdw": np.empty((1, self.n), np.uint8), "osd0": np.empty((1, self.n), np.uint8),
                               "bp": np.empty((1, self.n), np.uint8), "llr": np.empty((1, self.n), self._real),
                               "converge": np.empty(1, np.uint8), "iter": np.empty(1, np.int32),
                               "synd": np.empty((1, self.m), np.uint8)}
            one["args"] = tuple(one[k].ctypes.data for k in ("synd", "osdw", "osd0", "bp", "llr", "converge", "iter"))
            one["fn"] = _capi.load().bposd_decode_host
            one["rows"] = tuple(one[k][0] for k in ("synd", "osdw", "osd0", "bp", "llr"))
        synd0, osdw0 = one["rows"][0], one["rows"][1]
        if kind == "b":
            np.copyto(synd0, s, casting="unsafe")
        elif kind in "iu":
            np.bitwise_and(s, 1, out=synd0, casting="unsafe")
        else:
            synd0[:] = s.astype(np.int64) & 1
        ptr = one["args"]
        rc = one["fn"](self._h, ptr[0], 1, ptr[1], ptr[2], ptr[3], ptr[4], ptr[5], ptr[6])
        if rc:
            self._check(rc)
        self._lazy = (one["rows"], dtype)
        self._osd0 = self._bp = self._llr = None
        self.osdw_decoding = osdw0.astype(dtype)
        self.converge = bool(one["converge"][0])
        self.iter = int(one["iter"][0])
        return self.osdw_decoding

    # results of the last decode() that the caller may or may not look at: converted on first access
    @property
    def osd0_decoding(self):
        if self._osd0 is None and self._lazy is not None:
            self._osd0 = self._lazy[0][2].astype(self._lazy[1])
        return self._osd0

    @osd0_decoding.setter
    def osd0_decoding(self, value):
        self._osd0 = value

    @property
    def bp_decoding(self):
        if self._bp is None and self._lazy is not None:
            self._bp = self._lazy[0][3].astype(self._lazy[1])
        return self._bp

    @bp_decoding.setter
    def bp_decoding(self, value):
        self._bp = value

    @property
    def log_prob_ratios(self):
        if self._llr is None and self._lazy is not None:
            self._llr = self._lazy[0][4].astype(np.float64)
        return self._llr

    @log_prob_ratios.setter
    def log_prob_ratios(self, value):
        self._llr = value

    def _decode_host(self, synd_u8: np.ndarray, want_llr: bool = True, want_all: bool = True,
                     out: Optional[dict] = None, packed: bool = False) -> BatchResult:
        B = synd_u8.shape[0]
        out = out or {}
        nd = (self.n + 7) // 8 if packed else self.n

        def buf(key, shape, dtype, want=True):
            a = out.get(key)
            if a is None:
                return np.empty(shape, dtype) if want else None
            if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.shape == tuple(shape) and a.flags.c_contiguous):
                raise ValueError(f"out['{key}'] must be a C-contiguous {np.dtype(dtype)} array of shape {tuple(shape)}")
            return a

        osdw = buf("osdw", (B, nd), np.uint8)
        osd0 = buf("osd0", (B, nd), np.uint8, want_all)
        bp = buf("bp", (B, nd), np.uint8, want_all)
        llr = buf("llr", (B, self.n), self._real, want_llr)
        conv = buf("converge", (B,), np.uint8)
        it = buf("iter", (B,), np.int32)
        fn = _capi.load().bposd_decode_host_packed if packed else _capi.load().bposd_decode_host
        self._check(fn(self._h, _ptr(synd_u8), B, _ptr(osdw), _ptr(osd0), _ptr(bp), _ptr(llr), _ptr(conv), _ptr(it)))
        return BatchResult(osdw, osd0, bp, llr, None if conv is None else conv.astype(bool), it, packed)

    def decode_batch(self, syndromes, return_llr: bool = True, return_all: bool = True, priors=None, out=None,
                     weights=None, packed: bool = False):
        """Decode ``syndromes[B, m]`` (``packed=True``: bit-packed rows ``[B, ceil(m/8)]`` in, bit-packed decodings
        ``[B, ceil(n/8)]`` out -- :func:`pack_bits` / :func:`unpack_bits` give the layout; an eighth of the bytes over
        PCIe for host arrays).

        A CUDA ``torch.Tensor`` (uint8/bool/int, 0/1) is decoded in place on its device and the
        result holds CUDA tensors; a numpy array goes through pinned-staging copies
        (``bposd_decode_host``) and the result holds numpy arrays.  ``priors`` (CUDA tensor
        [B, n] of prior LLRs log((1-p)/p) in the handle's precision) selects per-shot channel priors and
        ``weights`` (CUDA float64 tensor [B, n] of log(1/p)) the matching per-shot OSD weights -- together
        they are the batched form of ``update_channel_probs`` before every shot
        (css_decode_sim.py:207-248).
        ``out`` may hold preallocated result buffers (keys ``osdw osd0 bp llr converge iter``; CUDA
        tensors for CUDA input, numpy arrays -- ideally pinned -- for host input).
        """
        out = out or {}
        try:
            import torch
        except Exception:  # pragma: no cover
            torch = None
        ms = (self.m + 7) // 8 if packed else self.m
        nd = (self.n + 7) // 8 if packed else self.n
        if torch is not None and isinstance(syndromes, torch.Tensor):
            if not syndromes.is_cuda:
                if priors is not None or weights is not None:
                    raise ValueError("per-shot priors / weights need CUDA syndromes (they are CUDA tensors)")
                return self.decode_batch(syndromes.numpy(), return_llr, return_all, out=out, packed=packed)
            if syndromes.dim() != 2 or syndromes.shape[1] != ms:
                raise ValueError(f"syndromes must have shape [B, {ms}]")
            if syndromes.device.index != self.device:
                raise ValueError(f"syndromes live on cuda:{syndromes.device.index}, decoder on cuda:{self.device}")
            s = syndromes
            if packed:
                if s.dtype != torch.uint8:
                    raise ValueError("bit-packed syndromes must be uint8")
            elif s.dtype == torch.bool:
                s = s.to(torch.uint8)
            elif s.dtype != torch.uint8:
                s = (s & 1).to(torch.uint8) if not s.dtype.is_floating_point else (s.to(torch.int64) & 1).to(torch.uint8)
            s = s.contiguous()  # (uint8 entries are reduced mod 2 by the kernels, like every other integer type)
            B = s.shape[0]
            dev = s.device
            tdt = torch.float64 if self.precision == 64 else torch.float32
            def buf(key, shape, dtype, want=True):
                t = out.get(key)
                if t is None:
                    return torch.empty(shape, dtype=dtype, device=dev) if want else None
                if not (t.is_cuda and t.device == dev and t.dtype == dtype and tuple(t.shape) == tuple(shape)
                        and t.is_contiguous()):
                    raise ValueError(f"out['{key}'] must be a contiguous {dtype} CUDA tensor of shape {tuple(shape)}")
                return t

            osdw = buf("osdw", (B, nd), torch.uint8)
            osd0 = buf("osd0", (B, nd), torch.uint8, return_all)
            bp = buf("bp", (B, nd), torch.uint8, return_all)
            llr = buf("llr", (B, self.n), tdt, return_llr)
            conv = buf("converge", (B,), torch.uint8)
            it = buf("iter", (B,), torch.int32)
            pri = None
            if priors is not None:
                if not (isinstance(priors, torch.Tensor) and priors.is_cuda and priors.shape == (B, self.n)
                        and priors.dtype == tdt):
                    raise ValueError("priors must be a CUDA tensor [B, n] in the decoder's precision")
                pri = priors.contiguous()
            wts = None
            if weights is not None:
                if not (isinstance(weights, torch.Tensor) and weights.is_cuda and weights.shape == (B, self.n)
                        and weights.dtype == torch.float64):
                    raise ValueError("weights must be a CUDA float64 tensor [B, n]")
                wts = weights.contiguous()
            o = _capi.Out(osdw.data_ptr(), osd0.data_ptr() if osd0 is not None else None,
                          bp.data_ptr() if bp is not None else None,
                          llr.data_ptr() if llr is not None else None, conv.data_ptr(), it.data_ptr())
            stream = torch.cuda.current_stream(dev).cuda_stream
            fn = _capi.load().bposd_decode_batch_packed if packed else _capi.load().bposd_decode_batch
            self._check(fn(self._h, s.data_ptr(), B, C.byref(o), pri.data_ptr() if pri is not None else None,
                           wts.data_ptr() if wts is not None else None, stream))
            return BatchResult(osdw, osd0, bp, llr, conv.bool(), it, packed)
        if priors is not None or weights is not None:
            raise ValueError("per-shot priors / weights need CUDA syndromes (they are CUDA tensors)")
        s = np.asarray(syndromes)
        if s.ndim != 2 or s.shape[1] != ms:
            raise ValueError(f"syndromes must have shape [B, {ms}]")
        if packed:
            if s.dtype != np.uint8:
                raise ValueError("bit-packed syndromes must be uint8")
            s = np.ascontiguousarray(s)
        else:
            s = np.ascontiguousarray((s.astype(np.int64) & 1).astype(np.uint8)) if s.dtype != np.uint8 else np.ascontiguousarray(s)
        return self._decode_host(s, want_llr=return_llr, want_all=return_all, out=out, packed=packed)

    # ------------------------------------------------------------------ harness step on the device
    def set_error_channel(self, pz=None, px=None, py=None):
        """Per-qubit Pauli probabilities for the device sampler (css_decode_sim.py:471-496 split)."""
        z = np.zeros(self.n) if pz is None else np.broadcast_to(np.asarray(pz, np.float64), (self.n,))
        x = np.zeros(self.n) if px is None else np.broadcast_to(np.asarray(px, np.float64), (self.n,))
        y = np.zeros(self.n) if py is None else np.broadcast_to(np.asarray(py, np.float64), (self.n,))
        t1, t2, t3 = (np.ascontiguousarray(prob_thresholds(v)) for v in (z, z + x, z + x + y))
        self._check(_capi.load().bposd_set_channel_thresholds(self._h, _ptr(t1), _ptr(t2), _ptr(t3)))

    def set_logicals(self, logicals):
        l = sp.csr_matrix(logicals).astype(np.uint8)
        l.data %= 2
        l.eliminate_zeros()
        l.sort_indices()
        if l.shape[1] != self.n:
            raise ValueError(f"logical operators must have {self.n} columns")
        ip = np.ascontiguousarray(l.indptr, dtype=np.int32)
        ix = np.ascontiguousarray(l.indices, dtype=np.int32)
        self._check(_capi.load().bposd_set_logicals(self._h, _ptr(ip), _ptr(ix), int(l.shape[0])))

    def sample_syndromes(self, seed: int, shot0: int, B: int, sector: int = 0, return_errors: bool = True,
                         packed: bool = False):
        """Device sampler + syndrome kernel; returns CUDA tensors (errors[B, n] or None, syndromes[B, m] -- or the
        bit-packed syndromes[B, ceil(m/8)] with ``packed``)."""
        import torch
        dev = torch.device("cuda", self.device)
        err = torch.empty((B, self.n), dtype=torch.uint8, device=dev) if return_errors else None
        syn = torch.empty((B, (self.m + 7) // 8 if packed else self.m), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        fn = _capi.load().bposd_sample_syndromes_packed if packed else _capi.load().bposd_sample_syndromes
        self._check(fn(self._h, int(seed), int(shot0), int(B), int(sector), err.data_ptr() if err is not None else None,
                       syn.data_ptr(), stream))
        return err, syn

    def _check_u8_cuda(self, t, name):
        """The kernels read B*n bytes through the raw pointer: anything but a uint8 [B, n] tensor on this decoder's GPU
        would be an out-of-bounds or illegal-address read, so it is a Python error here."""
        import torch
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.dim() == 2 and t.shape[1] == self.n):
            raise ValueError(f"{name} must be a CUDA uint8 tensor of shape [B, {self.n}]")
        if t.device.index != self.device:
            raise ValueError(f"{name} lives on cuda:{t.device.index}, decoder on cuda:{self.device}")

    def logical_check(self, errors, decodings, return_weight: bool = False):
        """fail[b] = ((L @ (e ^ d)) % 2).any() on the device for CUDA uint8 tensors [B, n];
        with ``return_weight`` also the Hamming weight of every residual (int32 [B])."""
        import torch
        self._check_u8_cuda(errors, "errors")
        self._check_u8_cuda(decodings, "decodings")
        if errors.shape != decodings.shape:
            raise ValueError("errors and decodings must have the same shape")
        B = errors.shape[0]
        fail = torch.empty(B, dtype=torch.uint8, device=errors.device)
        wt = torch.empty(B, dtype=torch.int32, device=errors.device) if return_weight else None
        stream = torch.cuda.current_stream(errors.device).cuda_stream
        self._check(_capi.load().bposd_logical_check(self._h, errors.contiguous().data_ptr(),
                                                     decodings.contiguous().data_ptr(), B, fail.data_ptr(),
                                                     None, None, wt.data_ptr() if wt is not None else None, stream))
        return (fail.bool(), wt) if return_weight else fail.bool()

    def channel_update(self, first_decoding, probs_if0, probs_if1):
        """Per-shot priors / OSD weights of this decoder given the other sector's decoding (CUDA uint8 [B, n]).
        Returns (priors [B, n] in the decoder's precision, weights [B, n] float64) for ``decode_batch``."""
        import torch
        self._check_u8_cuda(first_decoding, "first_decoding")
        d = first_decoding.contiguous()
        B = d.shape[0]
        p0 = np.ascontiguousarray(probs_if0, dtype=np.float64)
        p1 = np.ascontiguousarray(probs_if1, dtype=np.float64)
        if p0.shape != (self.n,) or p1.shape != (self.n,) or d.shape[1] != self.n:
            raise ValueError(f"channel update needs [B, {self.n}] decodings and two probability vectors of length {self.n}")
        tdt = torch.float64 if self.precision == 64 else torch.float32
        pri = torch.empty((B, self.n), dtype=tdt, device=d.device)
        wts = torch.empty((B, self.n), dtype=torch.float64, device=d.device)
        stream = torch.cuda.current_stream(d.device).cuda_stream
        self._check(_capi.load().bposd_channel_update(self._h, d.data_ptr(), B, _ptr(p0), _ptr(p1), pri.data_ptr(),
                                                      wts.data_ptr(), stream))
        return pri, wts

    def sample_and_decode(self, seed: int, shot0: int, B: int, sector: int = 0, counters=None) -> np.ndarray:
        """One Monte-Carlo step of B shots on the device; returns / accumulates int64 counters[8]:
        shots, bp_converged, bp_success, osd0_success, osdw_success, osd_invocations, bp_iterations,
        min_logical_weight."""
        if counters is None:
            counters = np.zeros(8, dtype=np.int64)
        stream = None
        try:
            import torch
            stream = torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream
        except Exception:
            pass
        self._check(_capi.load().bposd_sample_and_decode(self._h, int(seed), int(shot0), int(B), int(sector),
                                                         _ptr(counters), stream))
        return counters


class bposd_decoder(BpOsdDecoder):
    """Legacy-signature class the reference re-exports (src/bposd/__init__.py:1; README.md:176-187).

    Defaults follow the v1 signature as far as it can be recalled (``bp_method`` 0 = product-sum, ``osd_order`` -1);
    both reference call sites pass ``bp_method``, ``osd_method`` and ``osd_order`` explicitly (README.md:183-186;
    css_decode_sim.py:448-451), so the defaults are never on the reference's path."""

    def __init__(self, parity_check_matrix, error_rate=None, max_iter=0, bp_method="ps",
                 ms_scaling_factor=1.0, channel_probs=(None,), osd_order=-1, osd_method="osd0",
                 input_vector_type="syndrome", **kw):
        okey = str(osd_method).lower()
        if osd_order == -1:
            osd_order = 0
        if okey in ("osd0", "osd_0", "0", "zero"):
            osd_order = 0
        super().__init__(parity_check_matrix, error_rate=error_rate, channel_probs=list(channel_probs)
                         if channel_probs is not None and len(channel_probs) == 1 and channel_probs[0] is None
                         else channel_probs, max_iter=max_iter, bp_method=bp_method,
                         ms_scaling_factor=ms_scaling_factor, osd_method=osd_method, osd_order=osd_order,
                         input_vector_type=input_vector_type, **kw)
