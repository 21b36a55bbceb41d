"""hare_b200 -- B200 (sm_100a) implementation of Hare's batched closest-hit
Spatial_Partition.Shoot, behind Hare's own API names.

The product is hare_b200/libhare_b200.so (C ABI: include/hare_b200.h).  This package is
the host-side mirror used where no .NET toolchain exists; hare_b200/csharp/ holds the
P/Invoke binding for the real Hare_NC build.
"""
from ._lib import HareError, build, lib, SO_PATH  # noqa: F401
from .host import (KDTree, Octree, Point, Ray, Spatial_Partition, Topology,  # noqa: F401
                   Voxel_Grid, X_Event)


def init(device_ids=None, host_only=False):
    """hare_init: choose the CUDA devices of this process (host_only: build-time tooling, no GPU)."""
    import ctypes as C
    import numpy as np
    from ._lib import check
    if host_only:
        return check(lib().hare_init(None, -1), "hare_init")
    if device_ids is None:
        return check(lib().hare_init(None, 0), "hare_init")
    ids = np.ascontiguousarray(device_ids, np.int32)
    check(lib().hare_init(ids.ctypes.data_as(C.c_void_p), int(ids.shape[0])), "hare_init")


def launch_count():
    return int(lib().hare_launch_count())
