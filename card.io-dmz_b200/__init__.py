"""card.io-dmz_b200 -- B200-native implementation of card.io-dmz's detect -> warp -> OCR hot path.

The product is the CUDA shared library ``libb200dmz.so`` built from ``csrc/`` (C ABI in
``include/b200_dmz.h``).  This Python module is a thin ctypes mirror of that ABI for tests, the
benchmark and Python callers; it contains no arithmetic.  The directory name is not a valid Python
identifier, so load it with::

    import importlib.util, sys
    spec = importlib.util.spec_from_file_location("cardio_dmz_b200", "card.io-dmz_b200/__init__.py")
    mod = importlib.util.module_from_spec(spec); sys.modules[spec.name] = mod; spec.loader.exec_module(mod)
"""
from .binding import *  # noqa: F401,F403
from .binding import __all__  # noqa: F401
