"""Ray batches for the BASELINE configs."""
import numpy as np

from .rng import unit_directions


def rays_from_sources(n, srcs, stream=1, first=0):
    """n rays; ray i starts at srcs[i % len(srcs)] with an isotropic random direction.

    Returns (o, d): contiguous float64 (n, 3) arrays.  Ray_ID convention: i + 1.
    """
    srcs = np.asarray(srcs, dtype=np.float64).reshape(-1, 3)
    d = unit_directions(n, stream, first)
    o = np.ascontiguousarray(srcs[(np.arange(n) + first) % srcs.shape[0]])
    return o, np.ascontiguousarray(d)
