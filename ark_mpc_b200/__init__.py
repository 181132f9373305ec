"""ark_mpc_b200 — Blackwell-native online-phase gate engine for ark-mpc.

The product is the C-ABI library `lib/libarkmpc_b200.so` (include/arkmpc_b200.h); this package holds
its sources (`csrc/`), the ctypes binding (`_native`), a tensor-shaped engine (`engine`) and the
host-side mirror of the reference's operator surface (`fabric`)."""
from . import _native  # noqa: F401

__all__ = ["_native"]
