"""Per-shot Monte-Carlo loop in the shape of the reference's harness, with the CPU oracle as the decoder.

TEST INFRASTRUCTURE.  Mirrors the control flow of /root/reference/src/bposd/css_decode_sim.py
(`_single_run` :163-205, `_channel_update` :207-248, `_encoded_error_rates` :250-365) one shot at a
time, but draws its errors from the same Philox stream as the device sampler so that the batched GPU
harness can be compared counter for counter.
"""
import numpy as np

from oracle.oracle import OracleDecoder, sample_errors


def run(hx, hz, lx, lz, px, py, pz, seed, shots, channel_update, **dec_kw):
    n = hz.shape[1]
    bpd_z = OracleDecoder(hx, channel_probs=pz + py, **dec_kw)
    bpd_x = OracleDecoder(hz, channel_probs=px + py, **dec_kw)
    ex_all, ez_all = sample_errors(seed, 0, shots, pz, px, py)
    lx_d, lz_d = np.asarray(lx.todense()), np.asarray(lz.todense())
    hx_d, hz_d = np.asarray(hx.todense()), np.asarray(hz.todense())
    out = dict(run_count=0, bp_converge_count_x=0, bp_converge_count_z=0, bp_success_count=0,
               osd0_success_count=0, osdw_success_count=0, min_logical_weight=n)

    def update(first, a, b):  # probabilities of the second sector given the first sector's decoding
        p = np.zeros(n)
        for i in range(n):
            if first[i] == 1:
                p[i] = 0 if (a[i] + py[i]) == 0 else py[i] / (a[i] + py[i])
            else:
                p[i] = b[i] / (1 - a[i] - py[i])
        return p

    for s in range(shots):
        ex, ez = ex_all[s], ez_all[s]
        synd_x, synd_z = hz_d @ ex % 2, hx_d @ ez % 2
        if channel_update == "x->z":
            bpd_x.decode(synd_x)
            rx = {k: getattr(bpd_x, k).copy() if hasattr(getattr(bpd_x, k), "copy") else getattr(bpd_x, k)
                  for k in ("osdw_decoding", "osd0_decoding", "bp_decoding", "converge")}
            bpd_z.update_channel_probs(update(rx["osdw_decoding"], px, pz))
            bpd_z.decode(synd_z)
        elif channel_update == "z->x":
            bpd_z.decode(synd_z)
            bpd_x.update_channel_probs(update(bpd_z.osdw_decoding, pz, px))
            bpd_x.decode(synd_x)
        else:
            bpd_z.decode(synd_z)
            bpd_x.decode(synd_x)
        out["run_count"] += 1
        for name in ("osdw", "osd0"):
            rxv = (ex + getattr(bpd_x, name + "_decoding")) % 2
            rzv = (ez + getattr(bpd_z, name + "_decoding")) % 2
            if (lz_d @ rxv % 2).any():
                out["min_logical_weight"] = min(out["min_logical_weight"], int(rxv.sum()))
            elif (lx_d @ rzv % 2).any():
                out["min_logical_weight"] = min(out["min_logical_weight"], int(rzv.sum()))
            else:
                out[name + "_success_count"] += 1
        out["bp_converge_count_z"] += int(bool(bpd_z.converge))
        out["bp_converge_count_x"] += int(bool(bpd_x.converge))
        if bpd_z.converge and bpd_x.converge:
            rxv, rzv = (ex + bpd_x.bp_decoding) % 2, (ez + bpd_z.bp_decoding) % 2
            if not (lz_d @ rxv % 2).any() and not (lx_d @ rzv % 2).any():
                out["bp_success_count"] += 1
    return out
