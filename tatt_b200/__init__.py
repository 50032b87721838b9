"""tatt_b200 -- B200-native (sm_100a) implementation of the TATT/TSRN forward+backward hot path.

Public surface mirrors the reference's `model.tsrn`: `TSRN`, `TSRN_TL_TRANS` (see tatt_b200/tsrn.py).
Compute runs exclusively in the C-ABI CUDA library `tatt_b200/lib/libtatt_b200.so`
(include/tatt_b200.h); build it with `python -m tatt_b200.build`.
"""
from .tsrn import TSRN, TSRN_TL_TRANS  # noqa: F401

__version__ = "0.1.0"


def manual_seed(seed: int) -> None:
    """Seed the device-resident dropout RNG streams (graph-capture safe Philox state)."""
    from .ops import DeviceRNG
    DeviceRNG.manual_seed(seed)


def set_precision(mode: str) -> None:
    """'fp32' (default, bf16 hi/lo split x3 MMAs = fp32-grade parity), 'bf16' (single-plane bf16 MMAs), 'ffma'."""
    from . import ops
    ops.set_precision(mode)
