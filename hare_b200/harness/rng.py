"""Counter-based SplitMix64 (SURVEY.md 8(d) "RNG").

Only IEEE-exact operations (+ - * / sqrt) are applied to the doubles, so a C#,
C++ or CUDA port of this file produces identical bits.
"""
import numpy as np

SEED = np.uint64(0x48415245)  # "HARE"
_G = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(counter, stream=0):
    """counter: uint64 array. Returns uint64 array; state = SEED + stream*2^40 + (counter+1)*G."""
    with np.errstate(over="ignore"):
        c = np.asarray(counter, dtype=np.uint64)
        z = (SEED + np.uint64(stream) * np.uint64(1 << 40)) + (c + np.uint64(1)) * _G
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def uniform01(counter, stream=0):
    """Double in [0,1) from the top 53 bits."""
    return (splitmix64(counter, stream) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def unit_directions(n, stream, first=0, max_tries=64):
    """n unit vectors by rejection from [-1,1)^3: accept 1e-6 < r^2 <= 1, divide by sqrt(r^2).

    Ray i, attempt k uses counters ((first+i)*max_tries + k)*3 + {0,1,2}.
    """
    out = np.empty((n, 3), dtype=np.float64)
    todo = np.arange(n, dtype=np.uint64)
    for k in range(max_tries):
        if todo.size == 0:
            break
        base = ((todo + np.uint64(first)) * np.uint64(max_tries) + np.uint64(k)) * np.uint64(3)
        x = uniform01(base, stream) * 2.0 - 1.0
        y = uniform01(base + np.uint64(1), stream) * 2.0 - 1.0
        z = uniform01(base + np.uint64(2), stream) * 2.0 - 1.0
        r2 = (x * x + y * y) + z * z
        ok = (r2 > 1e-6) & (r2 <= 1.0)
        r = np.sqrt(r2[ok])
        idx = todo[ok].astype(np.int64)
        out[idx, 0] = x[ok] / r
        out[idx, 1] = y[ok] / r
        out[idx, 2] = z[ok] / r
        todo = todo[~ok]
    if todo.size:
        raise RuntimeError("rejection sampling did not converge")
    return out
