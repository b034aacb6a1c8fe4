"""Import shim: the reference's package name over the B200 implementation.

A script written against quantumgizmos/bp_osd (`from bposd import bposd_decoder`, `from bposd.hgp import hgp`,
`from bposd.css import css_code`, `from bposd.css_decode_sim import css_decode_sim` -- /root/reference/src/bposd/__init__.py:1,
README.md:150-216, examples/qldpc_decode_example.py) runs unmodified with this directory on the path: every name resolves to
bp_osd_b200.  `bposd.stab` (non-CSS stabiliser codes) is not on the decode path and is not provided.
"""
from bp_osd_b200 import __version__  # noqa: F401
from bp_osd_b200.decoder import BpOsdDecoder, bposd_decoder  # noqa: F401
from bp_osd_b200.css import css_code  # noqa: F401
from bp_osd_b200.hgp import hgp, hgp_single  # noqa: F401


def get_include():
    """Directory of the C-ABI header (the reference returns its package directory, src/bposd/__init__.py:6-8)."""
    import os
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
