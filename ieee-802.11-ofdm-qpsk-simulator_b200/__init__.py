"""B200-native 802.11a OFDM QPSK TX -> AWGN -> RX stage chain.

The product is libofdm_b200.so (hand-written sm_100a kernels behind the C-ABI declared in
include/ofdm_b200.h) plus the C host driver under host/.  This Python package is only the ctypes
binding used by tests/, bench.py and __graft_entry__.py; the directory name is not a Python
identifier, so load it with ``__graft_entry__.load_pkg()`` (importlib, module name ``ofdm_b200``).
"""
from .binding import *          # noqa: F401,F403
from . import binding           # noqa: F401
from . import sweep             # noqa: F401
