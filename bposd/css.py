"""bposd.css of the reference -> bp_osd_b200.css (see bposd/__init__.py)."""
from bp_osd_b200.css import *  # noqa: F401,F403
from bp_osd_b200 import css as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
