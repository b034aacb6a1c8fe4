"""bposd.hgp of the reference -> bp_osd_b200.hgp (see bposd/__init__.py)."""
from bp_osd_b200.hgp import *  # noqa: F401,F403
from bp_osd_b200 import hgp as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
