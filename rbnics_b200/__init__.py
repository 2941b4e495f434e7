"""rbnics_b200 -- B200-native engine for the RBniCS greedy-sweep hot path.

Host-side mirror of the reference's online backend for the path (``rbnics_b200.online``), the batched greedy
sweep (``rbnics_b200.sweep``), the offline basis linear algebra (``rbnics_b200.offline``) and the thin ctypes
binding to the sm_100a CUDA library (``rbnics_b200._lib``).  There is no CPU fallback: every compute entry
raises if ``librbnics_b200.so`` is missing or no B200 is visible.
"""
__version__ = "0.1.0"
