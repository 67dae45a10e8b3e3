"""
nimrud_b200 -- B200-native (sm_100a) implementation of nimrud's multiscale neighborhood
eigenfeature path.  Drop-in for `nimrud.minimal.multiscale` / `nimrud.utils.geometry.VoxelFilter`.

The compute path is hand-written CUDA behind a C ABI (include/nimrud_b200.h,
nimrud_b200/lib/libnimrud_b200.so).  There is no CPU fallback: importing the submodules works
anywhere, calling them without the library or without a GPU raises.
"""
from . import _lib            # noqa: F401
from . import geometry        # noqa: F401
from . import multiscale      # noqa: F401
from .multiscale import process_single_core, one_scale_single_core   # noqa: F401



def trim_memory():
    """release the scratch memory the library keeps cached on the current CUDA device (its private stream-ordered
    pool is invisible to torch's caching allocator); synchronises the device."""
    _lib.check(_lib.lib().nbr_trim_memory())
    from . import _results
    _results.trim()


__all__ = ["geometry", "multiscale", "process_single_core", "one_scale_single_core", "trim_memory"]
