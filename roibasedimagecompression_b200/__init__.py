"""B200-native hot path of the RHCCQ encoder (palette DBSCAN + region quantiser).

See DESIGN.md.  The CUDA library is loaded on first use (``_lib.lib()``) and
its absence is an error: there is no CPU fallback in this package.
"""
__version__ = "0.1.0"
