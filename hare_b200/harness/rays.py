"""Ray batches for the BASELINE configs."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .rng import unit_directions


def rays_from_sources(n, srcs, stream=1, first=0, threads=None, out=None):
    """n rays; ray i starts at srcs[(first + i) % len(srcs)] with an isotropic random direction.

    Returns (o, d): contiguous float64 (n, 3) arrays.  Ray_ID convention: i + 1.
    The generator is counter-based, so blocks are independent: large batches are produced in 1 M-ray blocks on `threads`
    host threads (numpy releases the GIL), identical to the single-threaded result.  `out` = (o, d) preallocated arrays
    (e.g. page-locked) to fill in place.
    """
    srcs = np.asarray(srcs, dtype=np.float64).reshape(-1, 3)
    if out is None:
        o = np.empty((n, 3), dtype=np.float64); d = np.empty((n, 3), dtype=np.float64)
    else:
        o, d = out
    block = 1 << 20

    def fill(b0):
        m = min(block, n - b0)
        d[b0:b0 + m] = unit_directions(m, stream, first + b0)
        o[b0:b0 + m] = srcs[(np.arange(b0, b0 + m) + first) % srcs.shape[0]]

    starts = list(range(0, n, block))
    nt = threads if threads is not None else min(len(starts), os.cpu_count() or 1)
    if nt <= 1 or len(starts) <= 1:
        for b0 in starts:
            fill(b0)
    else:
        with ThreadPoolExecutor(nt) as ex:
            list(ex.map(fill, starts))
    return o, d
