"""``css_decode_sim``: the reference's Monte-Carlo harness on top of the batched device pipeline.

Same keyword arguments, attributes and output keys as /root/reference/src/bposd/css_decode_sim.py
(``css_decode_sim(hx, hz, **input_dict)`` :63, defaults :65-84, output values :94-115,
``run_decode_sim`` :500, ``output_dict`` :557), so a script written for the reference
(examples/qldpc_decode_example.py:8-23) runs unchanged.  What differs is the inside of the loop:
the reference decodes one shot per Python iteration (:519-520 -> ``_single_run`` :163-205); here one
iteration is a *batch* of shots that never leaves the GPU:

    Philox sampler + H e mod 2 (both sectors)          css_decode_sim.py:465-498, 173/194
    -> decode first sector                              :174-195
    -> per-shot channel update (priors + OSD weights)   :207-248
    -> decode second sector with per-shot priors        :178-202
    -> residual . logicals, `elif` rule, counters       :250-365

All arithmetic is in libbposd_b200.so (C ABI, include/bposd_b200.h); this file only sequences
calls and keeps the bookkeeping (rates, error bars, JSON dump, resume, early stop).

Deliberate differences, all stated here:
* random numbers come from the device Philox4x32-10 stream keyed by ``seed`` with the global shot
  index as counter, not from numpy's MT19937 (:135-137): results are statistically, not
  shot-for-shot, comparable with the reference -- and independent of batch size and GPU count;
* rates, error bars, the save interval and the error-bar stopping rule (:545-553) are evaluated once
  per batch instead of once per shot;
* extra keywords ``batch_size``, ``precision`` (64 bit-exact / 32 fast) and ``device``.
Quirks kept for drop-in compatibility: ``run_decode_sim`` returns the doubly JSON-encoded text of
the reference (:555,567); the word-error-rate error bar uses the reference's expression (:286-290).
"""
from __future__ import annotations

import datetime
import json
import time

import numpy as np
import scipy.sparse as sp

from .css import css_code
from .decoder import BpOsdDecoder
from . import _capi

__all__ = ["css_decode_sim"]

_DEFAULTS = {
    "error_rate": None, "xyz_error_bias": [1, 1, 1], "target_runs": 100, "seed": 0,
    "bp_method": "minimum_sum", "ms_scaling_factor": 0.625, "max_iter": 0,
    "osd_method": "osd_cs", "osd_order": 2, "save_interval": 2, "output_file": None,
    "check_code": 1, "tqdm_disable": 0, "run_sim": 1, "channel_update": "x->z",
    "hadamard_rotate": 0, "hadamard_rotate_sector1_length": 0, "error_bar_precision_cutoff": 1e-3,
}
_EXTRA_DEFAULTS = {"batch_size": 65536, "precision": 64, "device": None}
_OUTPUTS = {
    "K": None, "N": None, "start_date": None, "runtime": 0.0, "runtime_readable": None, "run_count": 0,
    "bp_converge_count_x": 0, "bp_converge_count_z": 0, "bp_success_count": 0,
    "bp_logical_error_rate": 0, "bp_logical_error_rate_eb": 0,
    "osd0_success_count": 0, "osd0_logical_error_rate": 0.0, "osd0_logical_error_rate_eb": 0.0,
    "osdw_success_count": 0, "osdw_logical_error_rate": 0.0, "osdw_logical_error_rate_eb": 0.0,
    "osdw_word_error_rate": 0.0, "osdw_word_error_rate_eb": 0.0, "min_logical_weight": 1e9,
}
_NOT_SAVED = ("channel_probs_x", "channel_probs_z", "channel_probs_y", "hx", "hz")


class css_decode_sim:
    def __init__(self, hx=None, hz=None, **input_dict):
        for src in (input_dict, {k: v for k, v in _DEFAULTS.items() if k not in input_dict},
                    {k: v for k, v in _EXTRA_DEFAULTS.items() if k not in input_dict}):
            for key, value in src.items():
                setattr(self, key, value)
        for key, value in _OUTPUTS.items():
            if key not in self.__dict__:
                setattr(self, key, value)
        # everything known at this point is written to the output file, matrices excepted (:121-132)
        self.output_keys = [k for k in self.__dict__ if k not in _NOT_SAVED]

        if self.seed == 0 or self.run_count != 0:   # a resumed run never replays its old stream (:135-136)
            self.seed = int(np.random.randint(low=1, high=2**32 - 1))
            # multi-rank run: every rank must key Philox with the SAME seed (the stream is indexed by global shot), so
            # rank 0's draw is the one that counts -- and the one its output file records
            from .sharding import broadcast_from_rank0
            self.seed = broadcast_from_rank0(self.seed)
        print(f"RNG Seed: {self.seed}")

        self.hx = sp.csr_matrix(hx).astype(np.uint8)
        self.hz = sp.csr_matrix(hz).astype(np.uint8)
        self.N = self.hz.shape[1]
        if self.min_logical_weight == 1e9:
            self.min_logical_weight = self.N

        self._construct_code()
        self._error_channel_setup()
        self._decoder_setup()
        if self.run_sim:
            self.run_decode_sim()

    # ------------------------------------------------------------------ setup (host, once per run)
    def _construct_code(self):
        print("Constructing CSS code from hx and hz matrices...")
        qcode = css_code(self.hx, self.hz)
        self.lx, self.lz, self.K, self.N = qcode.lx, qcode.lz, qcode.K, qcode.N
        print("Checking the CSS code is valid...")
        if self.check_code and not qcode.test():
            raise Exception("Error: invalid CSS code. Check the form of your hx and hz matrices!")

    def _error_channel_setup(self):
        bias = np.array(self.xyz_error_bias, dtype=np.float64)
        if np.isinf(bias).any():                      # an infinite entry selects that Pauli alone (:395-406)
            which = int(np.flatnonzero(np.isinf(bias))[0])
            rates = [0, 0, 0]
            rates[which] = self.error_rate
            self.px, self.py, self.pz = rates
        else:
            self.px, self.py, self.pz = self.error_rate * bias / np.sum(bias)
        ones = np.ones(self.N)
        if self.hadamard_rotate == 0:
            self.channel_probs_x, self.channel_probs_z = ones * self.px, ones * self.pz
        elif self.hadamard_rotate == 1:               # X and Z swapped beyond sector 1 (:418-427)
            n1 = self.hadamard_rotate_sector1_length
            first = np.arange(self.N) < n1
            self.channel_probs_x = np.where(first, self.px, self.pz).astype(np.float64)
            self.channel_probs_z = np.where(first, self.pz, self.px).astype(np.float64)
        else:
            raise ValueError(f"The hadamard rotate attribute should be set to 0 or 1. Not '{self.hadamard_rotate}")
        self.channel_probs_y = ones * self.py
        for a in (self.channel_probs_x, self.channel_probs_y, self.channel_probs_z):
            a.setflags(write=False)

    def _decoder_setup(self):
        self.ms_scaling_factor = float(self.ms_scaling_factor)
        kw = dict(max_iter=self.max_iter, bp_method=self.bp_method, ms_scaling_factor=self.ms_scaling_factor,
                  osd_method=self.osd_method, osd_order=self.osd_order, precision=self.precision, device=self.device)
        # Z errors are seen by the X stabilisers and vice versa (:444-463)
        self.bpd_z = BpOsdDecoder(self.hx, channel_probs=self.channel_probs_z + self.channel_probs_y, **kw)
        self.bpd_x = BpOsdDecoder(self.hz, channel_probs=self.channel_probs_x + self.channel_probs_y, **kw)
        for d in (self.bpd_x, self.bpd_z):
            d.set_error_channel(pz=self.channel_probs_z, px=self.channel_probs_x, py=self.channel_probs_y)
        self.bpd_x.set_logicals(self.lz)   # residual X errors anticommute with Z logicals (:261)
        self.bpd_z.set_logicals(self.lx)
        px, py, pz = self.channel_probs_x, self.channel_probs_y, self.channel_probs_z
        with np.errstate(divide="ignore", invalid="ignore"):
            if self.channel_update == "x->z":      # Bayes update of the Z channel given the X decoding (:212-229)
                self._upd = (np.where(px + py == 0, 0.0, py / (px + py)), pz / (1 - px - py))
            elif self.channel_update == "z->x":    # and the mirror image (:232-248)
                self._upd = (np.where(pz + py == 0, 0.0, py / (pz + py)), px / (1 - pz - py))
            elif self.channel_update is None:
                self._upd = None
            else:
                raise ValueError(f"channel_update must be None, 'x->z' or 'z->x', not '{self.channel_update}'")

    # ------------------------------------------------------------------ one batch on the device
    def _batch(self, shot0: int, B: int, counters: np.ndarray):
        import ctypes as C
        import torch
        bx, bz = self.bpd_x, self.bpd_z
        ex, sx = bx.sample_syndromes(self.seed, shot0, B, sector=0)
        ez, sz = bz.sample_syndromes(self.seed, shot0, B, sector=1)
        if self.channel_update == "x->z":
            rx = bx.decode_batch(sx, return_llr=False)
            pri, wts = bz.channel_update(rx.osdw_decoding, self._upd[1], self._upd[0])
            rz = bz.decode_batch(sz, return_llr=False, priors=pri, weights=wts)
        elif self.channel_update == "z->x":
            rz = bz.decode_batch(sz, return_llr=False)
            pri, wts = bx.channel_update(rz.osdw_decoding, self._upd[1], self._upd[0])
            rx = bx.decode_batch(sx, return_llr=False, priors=pri, weights=wts)
        else:
            rz = bz.decode_batch(sz, return_llr=False)
            rx = bx.decode_batch(sx, return_llr=False)
        keep = []   # the tensors must outlive the asynchronous kernels that read them

        def sector(dx, dz, weights=True):
            fx = bx.logical_check(ex, dx, return_weight=weights)
            fz = bz.logical_check(ez, dz, return_weight=weights)
            fxf, wx = fx if weights else (fx, None)
            fzf, wz = fz if weights else (fz, None)
            fxu, fzu = fxf.to(torch.uint8), fzf.to(torch.uint8)
            keep.extend([fxu, fzu, wx, wz])
            return _capi.CssSector(fxu.data_ptr(), fzu.data_ptr(), wx.data_ptr() if wx is not None else None,
                                   wz.data_ptr() if wz is not None else None)

        s_w = sector(rx.osdw_decoding, rz.osdw_decoding)
        s_0 = sector(rx.osd0_decoding, rz.osd0_decoding)
        s_b = sector(rx.bp_decoding, rz.bp_decoding, weights=False)
        cx, cz = rx.converge.to(torch.uint8), rz.converge.to(torch.uint8)
        stream = torch.cuda.current_stream(torch.device("cuda", bx.device)).cuda_stream
        bx._check(_capi.load().bposd_css_counters(bx._h, B, C.byref(s_w), C.byref(s_0), C.byref(s_b), cx.data_ptr(),
                                                  cz.data_ptr(), counters.ctypes.data, stream))
        del keep

    # ------------------------------------------------------------------ bookkeeping (:250-365 per batch)
    def _rates(self):
        n = self.run_count
        for name in ("osdw", "osd0", "bp"):
            ler = 1 - getattr(self, f"{name}_success_count") / n
            eb = np.sqrt((1 - ler) * ler / n)
            setattr(self, f"{name}_logical_error_rate", float(ler))
            setattr(self, f"{name}_logical_error_rate_eb", float(eb))
            setattr(self, f"{name}_word_error_rate", float(1.0 - (1 - ler) ** (1 / self.K)))
            setattr(self, f"{name}_word_error_rate_eb", float(eb * ((1 - eb) ** (1 / self.K - 1)) / self.K))

    def run_decode_sim(self):
        self.start_date = datetime.datetime.fromtimestamp(time.time()).strftime("%A, %B %d, %Y %H:%M:%S")
        start = save_time = time.time()
        rank, world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
        except Exception:
            pass
        from .sharding import shard_range, all_reduce_vector
        try:
            from tqdm import tqdm
            bar = tqdm(total=int(self.target_runs), initial=int(self.run_count), disable=bool(self.tqdm_disable), ncols=0)
        except Exception:
            bar = None
        while self.run_count < self.target_runs:
            B = int(min(int(self.batch_size) * world, self.target_runs - self.run_count))
            lo, cnt = shard_range(B, rank, world)      # global shot indices: the stream does not depend on `world`
            c = np.zeros(8, dtype=np.int64)
            if cnt:
                self._batch(self.run_count + lo, cnt, c)
            # slot 7: "rank 0's save interval has elapsed".  The save / early-stop block below must be entered by every
            # rank or by none (a rank that stops while another loops on would leave that one in a collective for ever),
            # and wall clocks differ between ranks, so rank 0's clock decides and rides along with the counters.
            c[7] = 1 if (rank == 0 and int(time.time() - save_time) > self.save_interval) else 0
            if world > 1:
                c = all_reduce_vector(c, min_slots=(6,), device=self.bpd_x.device)
            gate_open = bool(c[7])
            self.run_count += int(c[0])
            self.bp_converge_count_x += int(c[1]); self.bp_converge_count_z += int(c[2])
            self.bp_success_count += int(c[3]); self.osd0_success_count += int(c[4]); self.osdw_success_count += int(c[5])
            if c[6] > 0 and c[6] < self.min_logical_weight:
                self.min_logical_weight = int(c[6])
            self._rates()
            if bar is not None:
                bar.update(int(c[0]))
                bar.set_description(
                    f"d_max: {self.min_logical_weight}; OSDW_WER: {self.osdw_word_error_rate*100:.3g}±"
                    f"{self.osdw_word_error_rate_eb*100:.2g}%; OSDW: {self.osdw_logical_error_rate*100:.3g}±"
                    f"{self.osdw_logical_error_rate_eb*100:.2g}%; OSD0: {self.osd0_logical_error_rate*100:.3g}±"
                    f"{self.osd0_logical_error_rate_eb*100:.2g}%;")
            now = time.time()
            if gate_open or self.run_count >= self.target_runs:
                self.runtime = (now - save_time) + self.runtime
                save_time = now
                self.runtime_readable = time.strftime("%H:%M:%S", time.gmtime(self.runtime))
                if self.output_file is not None and rank == 0:
                    with open(self.output_file, "w+") as f:
                        print(self.output_dict(), file=f)
                if (self.osdw_logical_error_rate_eb > 0 and
                        self.osdw_logical_error_rate_eb / self.osdw_logical_error_rate < self.error_bar_precision_cutoff):
                    print("\nTarget error bar precision reached. Stopping simulation...")
                    break
        if bar is not None:
            bar.close()
        self.wall_time = time.time() - start
        return json.dumps(self.output_dict(), sort_keys=True, indent=4)

    def output_dict(self):
        out = {k: v for k, v in self.__dict__.items() if k in self.output_keys}
        return json.dumps(out, sort_keys=True, indent=4, default=lambda o: o.item() if hasattr(o, "item") else str(o))
