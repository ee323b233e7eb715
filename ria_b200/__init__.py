"""ria_b200 -- B200-native batched receive chain for the RIA HF modem (hot path only).

Python here is plumbing (torch for device memory / streams / torch.distributed); all compute
is hand-written sm_100a CUDA in libria_b200.so behind the C ABI of include/ria_b200.h.
There is no CPU fallback.
"""
from ._lib import Context, RiaError, LIB_PATH, DECODE_RETRY_LADDER, DECODE_FP_REPAIR, DECODE_FULL, exported_symbols, lib  # noqa: F401
from . import fec  # noqa: F401
from . import ofdm  # noqa: F401
from . import mcdpsk  # noqa: F401
from . import sim  # noqa: F401
from . import sync  # noqa: F401
from . import dist  # noqa: F401
from . import selection  # noqa: F401
from . import txsynth  # noqa: F401
from . import stream  # noqa: F401

__all__ = ["Context", "RiaError", "LIB_PATH", "exported_symbols", "lib", "fec", "ofdm", "mcdpsk", "sim", "sync", "dist", "selection", "txsynth", "stream"]
