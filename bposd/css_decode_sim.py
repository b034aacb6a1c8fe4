"""bposd.css_decode_sim of the reference -> bp_osd_b200.css_decode_sim (see bposd/__init__.py)."""
from bp_osd_b200.css_decode_sim import *  # noqa: F401,F403
from bp_osd_b200 import css_decode_sim as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
