"""Ray batches for the BASELINE configs."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .rng import unit_directions


def source_index(i, n_src, order="interleaved", total=None):
    """Source of global ray number(s) i: round-robin ("interleaved") or one source after the other ("source-major": source s emits rays
    [ceil(s * total / n_src), ceil((s + 1) * total / n_src)) of a `total`-ray workload -- the order a caller that loops over its sources
    produces)."""
    i = np.asarray(i, dtype=np.int64)
    if order == "interleaved":
        return i % n_src
    if order != "source-major" or total is None:
        raise ValueError("order must be 'interleaved' or 'source-major' (with total)")
    return np.minimum((i * n_src) // max(1, int(total)), n_src - 1)


def rays_from_sources(n, srcs, stream=1, first=0, threads=None, out=None, order="interleaved", total=None):
    """n rays, global ray numbers first .. first + n - 1; ray i starts at srcs[source_index(i)] with an isotropic random direction
    (the direction depends on the ray number only, not on `order`).

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
        o[b0:b0 + m] = srcs[source_index(np.arange(b0, b0 + m) + first, srcs.shape[0], order, total)]

    starts = list(range(0, n, block))
    nt = threads if threads is not None else min(len(starts), os.cpu_count() or 1)
    if nt <= 1 or len(starts) <= 1:
        for b0 in starts:
            fill(b0)
    else:
        with ThreadPoolExecutor(nt) as ex:
            list(ex.map(fill, starts))
    return o, d


def sample_blocks(length, n, blocks=64):
    """Ascending indices of (about) n of `length` rays: `blocks` evenly spaced runs of consecutive rays -- a bounded sample that
    covers every source whatever the order of the batch.  The first n rays when the batch is too short for that."""
    n = int(min(n, length))
    per = n // blocks
    if per < 1 or length // blocks < per:
        return np.arange(n, dtype=np.int64)
    return (np.arange(blocks, dtype=np.int64)[:, None] * (length // blocks) + np.arange(per, dtype=np.int64)[None, :]).ravel()


def sample_rays(total, n, srcs, stream=1, order="interleaved", blocks=64):
    """(idx, o, d): the rays sample_blocks(total, n) picks from a `total`-ray workload, generated without generating the rest."""
    idx = sample_blocks(total, n, blocks)
    o = np.empty((len(idx), 3), dtype=np.float64); d = np.empty((len(idx), 3), dtype=np.float64)
    if len(idx) == 0:
        return idx, o, d
    cuts = np.flatnonzero(np.diff(idx) != 1) + 1
    pos = 0
    for run in np.split(idx, cuts):
        m = len(run)
        rays_from_sources(m, srcs, stream=stream, first=int(run[0]), threads=1, out=(o[pos:pos + m], d[pos:pos + m]), order=order, total=total)
        pos += m
    return idx, o, d
