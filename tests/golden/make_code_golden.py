"""Regenerates tests/golden/ref_hgp_logicals.npz from data the REFERENCE ships.

Unlike the decoder goldens (oracle outputs), these come from the reference itself: the logical
operators of its three example hypergraph-product codes,
/root/reference/examples/codes/hgp_codes/hgp_(4,7)-[[N,K,D]]_{lx,lz}.txt, written by the reference's
generate_codes.py with the real `ldpc.mod2` behind `bposd.hgp`, plus the three classical seeds
(classical_seed_codes/mkmn_*.txt) they were built from.  Stored bit-packed (np.packbits, axis 1).
tests/test_host_codes.py::test_reference_shipped_logicals checks the host GF(2) toolkit against them.

Run in the build container (needs /root/reference):  python tests/golden/make_code_golden.py
"""
import os

import numpy as np

REF = "/root/reference/examples/codes"
CODES = {"400_16_6": "mkmn_16_4_6", "625_25_8": "mkmn_20_5_8", "900_36_10": "mkmn_24_6_10"}


def main():
    out = {}
    for tag, seed in CODES.items():
        N, K, D = tag.split("_")
        stem = f"{REF}/hgp_codes/hgp_(4,7)-[[{N},{K},{D}]]"
        h = np.loadtxt(f"{REF}/classical_seed_codes/{seed}.txt").astype(np.uint8)
        out[f"seed_{tag}"] = np.packbits(h, axis=1)
        out[f"seed_shape_{tag}"] = np.array(h.shape)
        for which in ("lx", "lz"):
            a = np.loadtxt(f"{stem}_{which}.txt").astype(np.uint8)
            assert a.shape == (int(K), int(N))
            out[f"{which}_{tag}"] = np.packbits(a, axis=1)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_hgp_logicals.npz"), **out)
    # sha256 of every code file the reference ships (dense np.savetxt text, generate_codes.py:17-20): the on-disk format
    # test (tests/test_host_codes.py::test_regenerated_code_files_match_the_shipped_ones) regenerates them byte for byte
    import glob
    import hashlib
    import json
    dig = {os.path.basename(f): hashlib.sha256(open(f, "rb").read()).hexdigest() for f in sorted(glob.glob(f"{REF}/hgp_codes/*.txt"))}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_code_file_sha256.json"), "w") as fh:
        json.dump(dig, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
