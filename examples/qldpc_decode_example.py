"""The reference's example (examples/qldpc_decode_example.py of quantumgizmos/bp_osd) on the B200 decoder.

Same flow: take the 16-bit (3,4)-regular seed code `mkmn_16_4_6`, build its hypergraph product
[[400,16,6]], and run the Monte-Carlo BP+OSD simulation with the reference's option dictionary.  The only
differences are the import lines and that 1000 shots are one batch on the GPU instead of 1000 Python
iterations.  Pass a path to a dense 0/1 text matrix (np.savetxt format) to use another seed code.

    python examples/qldpc_decode_example.py [seed_matrix.txt]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bp_osd_b200 import codes  # noqa: E402
from bp_osd_b200.hgp import hgp  # noqa: E402
from bp_osd_b200.css_decode_sim import css_decode_sim  # noqa: E402

h = codes.load_alist_txt(sys.argv[1]) if len(sys.argv) > 1 else codes.mkmn_16_4_6()
qcode = hgp(h)  # construct quantum LDPC code using the symmetric hypergraph product
qcode.test()

osd_options = {
    "error_rate": 0.05,
    "target_runs": 1000,
    "xyz_error_bias": [0, 0, 1],
    "output_file": "test.json",
    "bp_method": "ms",
    "ms_scaling_factor": 0,
    "osd_method": "osd_cs",
    "osd_order": 42,
    "channel_update": None,
    "seed": 42,
    "max_iter": 0,
    "output_file": "test.json",
}

lk = css_decode_sim(hx=qcode.hx, hz=qcode.hz, **osd_options)
print(f"[[{lk.N},{lk.K}]] runs={lk.run_count} OSDW logical error rate {lk.osdw_logical_error_rate:.4f} "
      f"+- {lk.osdw_logical_error_rate_eb:.4f}, OSD0 {lk.osd0_logical_error_rate:.4f}, "
      f"BP converged x/z {lk.bp_converge_count_x}/{lk.bp_converge_count_z}")
