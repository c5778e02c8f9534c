"""ministark_b200 -- B200 (sm_100a) implementation of mini-stark's data-parallel prover core.

The compute path is libministark.so (hand-written CUDA behind the C ABI of include/ministark.h);
this package is the thin host side: the ctypes binding, a device handle that borrows torch for
memory/streams, and the mirror of the reference's public API (air / starks modules)."""
from .api import BABYBEAR, GOLDILOCKS, Context, MiniStarkError, StarkParams  # noqa: F401

__all__ = ["Context", "MiniStarkError", "StarkParams", "GOLDILOCKS", "BABYBEAR"]
