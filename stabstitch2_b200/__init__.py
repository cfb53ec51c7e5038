"""stabstitch2_b200 - B200-native (sm_100a) StabStitch++ inference hot path.

Host side in Python mirroring the reference's module names (spatial_network, temporal_network,
smooth_network, utils.*, grid_res); all device work is hand-written CUDA in libss2.so behind
the C ABI of include/ss2.h.  No CPU fallback."""
from . import grid_res  # noqa: F401

__all__ = ["grid_res"]
