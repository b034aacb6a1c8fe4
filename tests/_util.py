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


# ---- the CPU oracle on all host cores (large parity samples) -----------------------------------
_PW = {}


def _par_init(H, p, kw):
    from oracle.oracle import OracleDecoder
    _PW["dec"] = OracleDecoder(H, error_rate=p, **kw)


def _par_work(syn):
    return _PW["dec"].decode_batch(syn, want_llr=_PW.get("llr", True))


def oracle_decode_parallel(H, syn, p, kw, procs=None, chunk=None):
    """OracleDecoder(H, error_rate=p, **kw).decode_batch(syn), the shots spread over `procs` forked workers."""
    import multiprocessing as mp
    from functools import partial
    procs = procs or len(os.sched_getaffinity(0))
    B = syn.shape[0]
    chunk = chunk or max(1, min(2000, (B + 4 * procs - 1) // (4 * procs)))
    parts = [syn[i:i + chunk] for i in range(0, B, chunk)]
    with mp.get_context("fork").Pool(procs, initializer=partial(_par_init, H, p, kw)) as pool:
        res = pool.map(_par_work, parts)
    return {k: np.concatenate([r[k] for r in res]) for k in res[0]}
