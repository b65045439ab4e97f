"""arbitrarystyletransfer_b200 -- B200-native (sm_100a) AdaIN style-transfer hot path.

Drop-in for the hot-path surface of rwickman/ArbitraryStyleTransfer's ``models.py``, ``losses.py``
and ``model_util.py``; all device work is hand-written CUDA behind the C ABI of
``libast_b200.so`` (include/ast_b200.h).  Importing the package does not need a GPU; calling
any op does, and there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (ctypes binding; the library itself is loaded on first use)
