"""B200-native VirtualRadar (drop-in for layers/virtual_radar.py of
itskalvik/skeleton-action-recognition): hand-written sm_100a kernels behind a C ABI."""
from .layers.virtual_radar import VirtualRadar, edges  # noqa: F401
from .sharding import shard_bounds, sharded_forward  # noqa: F401
from .upsample import pad_frames, pad_frames_notebook  # noqa: F401

__all__ = ["VirtualRadar", "edges", "shard_bounds", "sharded_forward", "pad_frames", "pad_frames_notebook"]
