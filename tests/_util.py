import ast
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    n, m = int(z["n"]), int(z["m"])
    unpack = lambda a, w: np.unpackbits(a, axis=1)[:, :w]
    return dict(cfg=int(z["cfg"]), p=float(z["p"]), kw=ast.literal_eval(str(z["kw"])),
                syndromes=unpack(z["syndromes"], m), osdw=unpack(z["osdw"], n), osd0=unpack(z["osd0"], n),
                bp=unpack(z["bp"], n), llr=z["llr"], converge=z["converge"], iter=z["iter"])


def golden_names():
    """Decoder goldens only (ref_*.npz are code-construction data from the reference, see make_code_golden.py)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith("ref_"))


def random_syndromes(H, p, B, seed):
    rng = np.random.default_rng(seed)
    e = (rng.random((B, H.shape[1])) < p).astype(np.uint8)
    s = np.asarray((H @ e.T) % 2, dtype=np.uint8).T.copy()
    return e, s
