"""B200-native BP+OSD decoder behind the ``bposd_decoder`` API."""
__version__ = "0.1.0"
