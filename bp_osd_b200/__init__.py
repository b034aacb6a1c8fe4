"""B200-native BP+OSD decoder behind the ``bposd_decoder`` API.

``bposd_decoder`` / ``BpOsdDecoder`` call hand-written sm_100a CUDA kernels through the C ABI in
include/bposd_b200.h; ``css_code`` / ``hgp`` code construction stays on the host.
"""
__version__ = "0.1.0"

from .css import css_code  # noqa: F401
from .hgp import hgp, hgp_single  # noqa: F401
from .decoder import BpOsdDecoder, bposd_decoder, BatchResult, pack_bits, unpack_bits  # noqa: F401
